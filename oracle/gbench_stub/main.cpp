// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE ONLY: the main() Google Benchmark's benchmark_main library would provide.
#include <benchmark/benchmark.h>
BENCHMARK_MAIN()

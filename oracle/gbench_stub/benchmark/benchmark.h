// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE ONLY.  The sliver of the Google Benchmark interface the reference's src/bench_gpu.cu, src/bench_cpu.cu and third_party/fss/bench.cu use (the library
// itself is fetched from the network by the reference's CMake and is not available offline): a State that runs a fixed
// number of iterations (timed by SetIterationTime when the benchmark calls it, else by the wall clock of the loop body), BENCHMARK(fn)->Name(..)->UseManualTime() registration, and a main() that runs every
// registered benchmark and prints one line each.  Own code; it only has to be enough to build that file UNMODIFIED against
// include/ of this repository (oracle/Makefile: refbench) so that the reference's own benchmark harness runs on this library.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace benchmark {

class State {
 public:
  explicit State(int iterations) : left_(iterations) {}
  struct Sentinel {};
  struct Iterator {
    State *s;
    bool operator!=(Sentinel) const { return s->left_ > 0; }
    void operator++() {  // end of one iteration: without a manual time, the wall clock of the loop body counts
      if (!s->manual_) s->Record(std::chrono::duration<double>(std::chrono::steady_clock::now() - s->started_).count());
      s->manual_ = false;
      --s->left_;
      s->started_ = std::chrono::steady_clock::now();
    }
    int operator*() const { return 0; }
  };
  Iterator begin() {
    started_ = std::chrono::steady_clock::now();
    return Iterator{this};
  }
  Sentinel end() { return Sentinel{}; }
  void SetIterationTime(double seconds) {
    manual_ = true;
    Record(seconds);
  }
  long long iterations() const { return timed_; }
  void SetItemsProcessed(long long items) { items_ = items; }
  long long items_processed() const { return items_; }
  double mean_seconds() const { return timed_ ? total_ / timed_ : 0.0; }
  double best_seconds() const { return best_ < 0 ? 0.0 : best_; }
  int timed() const { return timed_; }

 private:
  void Record(double seconds) {
    total_ += seconds;
    ++timed_;
    if (best_ < 0 || seconds < best_) best_ = seconds;
  }
  std::chrono::steady_clock::time_point started_{};
  bool manual_ = false;
  int left_;
  int timed_ = 0;
  long long items_ = 0;
  double total_ = 0, best_ = -1;
};

// keeps `value` alive in the eyes of the optimiser (what the real library's function of this name is for)
template <typename T>
inline void DoNotOptimize(T const &value) {
  asm volatile("" : : "r,m"(value) : "memory");
}
template <typename T>
inline void DoNotOptimize(T &value) {
  asm volatile("" : "+r,m"(value) : : "memory");
}

namespace internal {
struct Benchmark {
  void (*fn)(State &);
  std::string name;
  Benchmark *Name(const char *n) {
    name = n;
    return this;
  }
  Benchmark *UseManualTime() { return this; }
};
inline std::vector<Benchmark *> &Registry() {
  static std::vector<Benchmark *> r;
  return r;
}
inline Benchmark *Register(void (*fn)(State &), const char *name) {
  Benchmark *b = new Benchmark{fn, name};
  Registry().push_back(b);
  return b;
}
}  // namespace internal

inline internal::Benchmark *RegisterBenchmark(const char *name, void (*fn)(State &)) { return internal::Register(fn, name); }

// Runs every registered benchmark whose name contains `filter` (argv[1], optional): 2 warm-up + FSS_BENCH_ITERS (10) iterations.
inline int RunAll(int argc, char **argv) {
  const char *filter = argc > 1 ? argv[1] : "";
  const char *e = std::getenv("FSS_BENCH_ITERS");
  const int iters = e && *e ? std::atoi(e) : 10;
  for (internal::Benchmark *b : internal::Registry()) {
    if (b->name.find(filter) == std::string::npos) continue;
    State warm(2);
    b->fn(warm);
    State st(iters);
    b->fn(st);
    const double per_iter = st.timed() ? double(st.items_processed()) / st.timed() : 0.0;
    std::printf("%-44s mean %12.3f us   best %12.3f us   %14.0f items/s   (%d iterations)\n", b->name.c_str(),
        st.mean_seconds() * 1e6, st.best_seconds() * 1e6, st.mean_seconds() > 0 ? per_iter / st.mean_seconds() : 0.0, st.timed());
    std::fflush(stdout);
  }
  return 0;
}

}  // namespace benchmark

#define FSS_GBENCH_CAT2(a, b) a##b
#define FSS_GBENCH_CAT(a, b) FSS_GBENCH_CAT2(a, b)
#define BENCHMARK(...) \
  static ::benchmark::internal::Benchmark *FSS_GBENCH_CAT(fss_gbench_reg_, __COUNTER__) = ::benchmark::internal::Register(__VA_ARGS__, #__VA_ARGS__)
#define BENCHMARK_MAIN() \
  int main(int argc, char **argv) { return ::benchmark::RunAll(argc, argv); }

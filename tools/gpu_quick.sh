#!/bin/bash
# quick A/B session: parity tests, then bench (kernel numbers only) per point mode
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
for m in ${MODES:-default 0 1}; do
  if [ "$m" = default ]; then unset FSSB200_POINT_MODE; else export FSSB200_POINT_MODE=$m; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_mode_$m.json 2> gpurun_out/bench_mode_$m.err
done
tail -3 gpurun_out/pytest_gpu.log
python tools/summarize.py gpurun_out/bench_mode_*.json

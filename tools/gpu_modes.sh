#!/bin/bash
# parity tests under each point mode in $MODES, then kernel-only bench per mode
set -u
mkdir -p gpurun_out
for m in ${MODES:-default}; do
  if [ "$m" = default ]; then unset FSSB200_POINT_MODE; else export FSSB200_POINT_MODE=$m; fi
  ( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_mode_$m.log
  timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_mode_$m.json 2> gpurun_out/bench_mode_$m.err
  echo "mode $m: $(tail -1 gpurun_out/pytest_gpu_mode_$m.log)"; tail -2 gpurun_out/bench_mode_$m.err
done
python tools/summarize.py gpurun_out/bench_mode_*.json

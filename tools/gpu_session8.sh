#!/bin/bash
# Session 8: parity after the session-7 gen / relayout / Grotto changes, smoke, full bench line, section-8(f) numbers.
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_s8.log
tail -3 gpurun_out/pytest_gpu_s8.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_s8.log 2>&1; tail -2 gpurun_out/smoke_s8.log
timeout 900 python bench.py > gpurun_out/bench_s8.json 2> gpurun_out/bench_s8.err; tail -3 gpurun_out/bench_s8.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_s8_ref.json 2> gpurun_out/bench_s8_ref.err
timeout 600 python tools/bench_next.py --out gpurun_out/bench_next_s8.json > gpurun_out/bench_next_s8.log 2>&1; tail -5 gpurun_out/bench_next_s8.log
python tools/summarize.py gpurun_out/bench_s8.json
echo done

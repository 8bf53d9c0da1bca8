#!/bin/bash
# Session 7: parity tests after the gen (TMA-written Cw tiles, DCF gen without spills), relayout (tile transpose)
# and Grotto (bit-packed leaf rows, multi-level parity kernel) changes; section-8(f) numbers before/after.
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_s7.log
tail -3 gpurun_out/pytest_gpu_s7.log
timeout 600 python tools/bench_next.py --out gpurun_out/bench_next_s7.json > gpurun_out/bench_next_s7.log 2>&1
FSSB200_GEN_MODE=0 timeout 300 python tools/bench_next.py --sections f1 --out gpurun_out/bench_next_s7_genmode0.json > gpurun_out/bench_next_s7_genmode0.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_s7.log 2>&1; tail -2 gpurun_out/smoke_s7.log
echo done

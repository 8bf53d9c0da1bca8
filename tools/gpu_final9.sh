#!/bin/bash
# Session 9 final: ncu launch list of the bench command, full captures of the final EvalAll kernels, full bench line (own + reference arm).
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches9.csv \
   python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches9.log 2>&1
WHICH="evalall_dpf grotto" bash tools/gpu_ncu9.sh > gpurun_out/ncu9.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_s9.json 2> gpurun_out/bench_s9.err; tail -3 gpurun_out/bench_s9.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_s9_ref.json 2> gpurun_out/bench_s9_ref.err
timeout 600 python tools/bench_next.py --out gpurun_out/bench_next_s9.json > gpurun_out/bench_next_s9.log 2>&1; tail -2 gpurun_out/bench_next_s9.log | cut -c1-300
python tools/summarize.py gpurun_out/bench_s9.json
echo done

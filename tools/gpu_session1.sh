#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (own + reference arm), ncu launch list, ncu full captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
kill $SMI
# launch list of the bench command (short run, no e2e/cpu legs)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches.log 2>&1
# full captures: C2 point kernel, C3 DCF point kernel, EvalAll kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:point_kernel -s 3 -c 1 -o gpurun_out/prof_dpf_point \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/prof_dpf_point.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:evalall_kernel -s 1 -c 1 -o gpurun_out/prof_evalall \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --keys 65536 --evalall-bits 26 --evalall-keys 2 > gpurun_out/prof_evalall.log 2>&1
echo done
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_ref.json; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err

#!/bin/bash
# ncu --set full of the C2 point kernel (default mode) and the C3 DCF point kernel
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:point_kernel -s 3 -c 1 -o gpurun_out/prof_dpf_point_v3 -f \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/prof_dpf_point_v3.log 2>&1
# C3: point_kernel launches: C2 warmup 3 + 1 timed + 1 recon = 5, then C3 launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:point_kernel -s 8 -c 1 -o gpurun_out/prof_dcf_point_v3 -f \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --evalall-bits 20 --evalall-keys 1 > gpurun_out/prof_dcf_point_v3.log 2>&1
ls -la gpurun_out/*.ncu-rep

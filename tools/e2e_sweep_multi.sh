#!/bin/bash
# N concurrent instances of tools/e2e_sweep.py, one per GPU, with the environment torchrun would give them
# (LOCAL_WORLD_SIZE): how the host-buffer pipeline behaves when several ranks share the host.  usage: e2e_sweep_multi.sh N
# THREADS (default 0 = the library's per-rank budget) and TUNE (default multi; none = the three host modes only) pass through.
N=${1:-2}
mkdir -p gpurun_out
for r in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$r LOCAL_WORLD_SIZE=$N LOCAL_RANK=$r timeout 900 python tools/e2e_sweep.py --threads ${THREADS:-0} --tune ${TUNE:-multi} \
    --out gpurun_out/e2e_sweep_n${N}_r$r.jsonl > gpurun_out/e2e_sweep_n${N}_r$r.log 2>&1 &
done
wait
tail -n 12 gpurun_out/e2e_sweep_n${N}_r0.log

#!/bin/bash
# Session 9 ncu --set full captures: EvalAll kernels after the cooperative bottom stage, DCF gen after the looped level.
set -u
mkdir -p gpurun_out
for w in ${WHICH:-evalall_dpf evalall_dcf}; do
  case $w in
    evalall_dcf) k=dcf_evalall_kernel;; evalall_*) k='^evalall_kernel';; grotto) k='^evalall_kernel';; gen_*) k=gen_kernel;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof9_$w -f \
     python tools/prof_one.py $w > gpurun_out/prof9_$w.log 2>&1
  tail -2 gpurun_out/prof9_$w.log
done
ls -la gpurun_out/*.ncu-rep

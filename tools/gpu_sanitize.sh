#!/bin/bash
# compute-sanitizer over a small invocation of every kernel family (tools/sanitize_cases.py).
# Writes gpurun_out/sanitize_<tool>.log; summary lines at the end.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  timeout ${LIMIT:-600} $CS --tool $tool --print-limit 20 --error-exitcode 9 \
     python tools/sanitize_cases.py ${CASES:-} > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ALL OK' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done

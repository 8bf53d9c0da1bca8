#!/bin/bash
# Session 9: parity after the EvalAll cooperative bottom stage + looped node expansions, section-8(f) numbers, bench line.
set -u
TAG=${1:-s9}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python tools/bench_next.py --sections ${SECTIONS:-f1,f3} --out gpurun_out/bench_next_$TAG.json > gpurun_out/bench_next_$TAG.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/bench_next_$TAG.json"))
for k,v in d["rows"].items(): print(k, round(v.get("ms",0),3), v.get("leaves_per_s") or v.get("keys_per_s") or v.get("evals_per_s"), v.get("lsu_roofline_frac"))
PY
if [ "${BENCH:-1}" = 1 ]; then
  timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
  python tools/summarize.py gpurun_out/bench_$TAG.json
fi
echo done

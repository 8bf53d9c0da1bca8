#!/bin/bash
# Session 9 verification: GPU parity suite, smoke, full bench line (own + reference arm), section-8(f) numbers.
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_gpu_v9.log; tail -2 gpurun_out/pytest_gpu_v9.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_v9.log 2>&1; tail -1 gpurun_out/smoke_v9.log
timeout 900 python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; tail -2 gpurun_out/bench_v9.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_v9_ref.json 2> gpurun_out/bench_v9_ref.err
timeout 600 python tools/bench_next.py --out gpurun_out/bench_next_v9.json > gpurun_out/bench_next_v9.log 2>&1
python tools/summarize.py gpurun_out/bench_v9.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_v9.json").read().strip().splitlines()[-1])
e=d["e2e"]; print("e2e", e["value"], e["ms_per_step"], e.get("host_pack_threads"), "direct", e.get("direct_copy",{}).get("value"))
r=json.loads(open("gpurun_out/bench_v9_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
echo done

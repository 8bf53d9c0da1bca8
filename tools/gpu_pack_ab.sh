#!/bin/bash
# A/B of the packing host path: FSSB200_PACK_DIRECT_EVERY = 0 (all packed), 2, 3, 4, 6 and FSSB200_PACK_THREADS variants
set -u
mkdir -p gpurun_out
for de in 0 2 3 4 6; do
  FSSB200_PACK_DIRECT_EVERY=$de timeout 300 python bench.py --no-cpu --no-extra --steps 5 --warmup 3 > gpurun_out/pack_ab_$de.json 2> gpurun_out/pack_ab_$de.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/pack_ab_$de.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("direct_every=$de", "e2e %.1f M evals/s (%.1f ms)"%(e["value"]/1e6, e["ms_per_step"]), "direct %.1f ms"%e.get("direct_copy",{}).get("ms_per_step",0), "threads", e.get("host_pack_threads"))
PY
done
for th in 8 12; do
  FSSB200_PACK_THREADS=$th timeout 300 python bench.py --no-cpu --no-extra --steps 5 --warmup 3 > gpurun_out/pack_ab_t$th.json 2> gpurun_out/pack_ab_t$th.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/pack_ab_t$th.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("threads=$th direct_every=4", "e2e %.1f M evals/s (%.1f ms)"%(e["value"]/1e6, e["ms_per_step"]))
PY
done

#!/bin/bash
# Synchronised (torchrun, barrier per step) e2e legs of bench.py for several FSSB200_PACK_THREADS values per rank:
# how many packing threads per rank the host-buffer pipeline should use when N ranks share the host.
# usage: e2e_threads_torchrun.sh N "0 4 6 8"      (0 = the library's default budget)
N=$1; shift
mkdir -p gpurun_out
: > gpurun_out/e2e_threads_n$N.jsonl
for T in ${1:-0 4 6 8 10}; do
  if [ "$T" = "0" ]; then unset FSSB200_PACK_THREADS; else export FSSB200_PACK_THREADS=$T; fi
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-check > gpurun_out/e2e_threads_tmp.json 2> gpurun_out/e2e_threads_tmp.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-extra --no-cpu --no-check > gpurun_out/e2e_threads_tmp.json 2> gpurun_out/e2e_threads_tmp.err
  fi
  python - "$T" <<'PY' | tee -a gpurun_out/e2e_threads_n$N.jsonl
import json, sys
try:
    d = json.loads([l for l in open('gpurun_out/e2e_threads_tmp.json') if l.startswith('{')][-1])
    e = d['e2e']
    print(json.dumps({"threads": sys.argv[1], "n_gpus": d['n_gpus'], "e2e_ms": e.get('ms_per_step'), "e2e_value": e['value'],
                      "packed_keys": e.get('packed_keys'), "direct_keys": e.get('direct_keys'), "host_threads": e.get("host_threads"),
                      "direct_ms": (e.get('direct_copy') or {}).get('ms_per_step'), "staged_ms": (e.get('staged_only') or {}).get('ms_per_step')}))
except Exception as ex:
    print(json.dumps({"threads": sys.argv[1], "error": str(ex)}))
PY
done

#!/bin/bash
# Synchronised (torchrun, barrier per step) e2e legs of bench.py under several environments: how the host-buffer
# pipeline should be configured when N ranks share the host.  Each configuration is a thread count per rank
# (0 = the library's default budget) optionally followed by ",KEY=VALUE,..." pipeline knobs.
# usage: e2e_threads_torchrun.sh N "0 4 6 8"        or  "0 0,FSSB200_PIPE_PIECE_BITS=14 0,FSSB200_PIPE_PIECE_BITS=12,FSSB200_PIPE_SLOTS=8"
N=$1; shift
mkdir -p gpurun_out
: > gpurun_out/e2e_threads_n$N.jsonl
for CFG in ${1:-0 4 6 8 10}; do
  T=${CFG%%,*}
  EXTRA=""
  if [ "$CFG" != "$T" ]; then EXTRA=$(echo "${CFG#*,}" | tr ',' ' '); fi
  if [ "$T" != "0" ]; then EXTRA="$EXTRA FSSB200_PACK_THREADS=$T"; fi
  if [ "$N" = "1" ]; then
    env $EXTRA timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu --no-check > gpurun_out/e2e_threads_tmp.json 2> gpurun_out/e2e_threads_tmp.err
  else
    env $EXTRA timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-extra --no-cpu --no-check > gpurun_out/e2e_threads_tmp.json 2> gpurun_out/e2e_threads_tmp.err
  fi
  python - "$CFG" <<'PY' | tee -a gpurun_out/e2e_threads_n$N.jsonl
import json, sys
try:
    d = json.loads([l for l in open('gpurun_out/e2e_threads_tmp.json') if l.startswith('{')][-1])
    e = d['e2e']
    print(json.dumps({"config": sys.argv[1], "n_gpus": d['n_gpus'], "e2e_ms": e.get('ms_per_step'), "e2e_value": e['value'],
                      "packed_keys": e.get('packed_keys'), "direct_keys": e.get('direct_keys'), "host_threads": e.get("host_threads"),
                      "direct_ms": (e.get('direct_copy') or {}).get('ms_per_step'), "staged_ms": (e.get('staged_only') or {}).get('ms_per_step')}))
except Exception as ex:
    print(json.dumps({"config": sys.argv[1], "error": str(ex)}))
PY
done

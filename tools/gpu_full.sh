#!/bin/bash
# full session: parity tests + full bench line (with e2e and cpu baseline) [+ N=2 run if GPUS2=1]
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "${GPUS2:-0}" = 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
fi
tail -3 gpurun_out/pytest_gpu.log; tail -5 gpurun_out/bench.err
python tools/summarize.py gpurun_out/bench.json gpurun_out/bench_n2.json

#!/bin/bash
# scaling session: bench.py under torchrun at N GPUs (N = $NGPUS, default 8), own arm + reference arm
set -u
N=${NGPUS:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/scale_n${N}_gpus.csv 2>&1
nvidia-smi topo -m > gpurun_out/scale_n${N}_topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_ref.json 2> gpurun_out/bench_n${N}_ref.err
tail -5 gpurun_out/bench_n${N}.err
python tools/summarize.py gpurun_out/bench_n${N}.json

#!/bin/bash
# N-GPU session (gpurun --gpus N): host topology, torchrun bench as the driver launches it, the single-process
# (fssb200_*_multi) bench, and the concurrent host-pipeline sweep.  usage: gpu_multi.sh N [steps...]
N=$1; shift
mkdir -p gpurun_out
for step in "${@:-host bench single sweep}"; do
  case "$step" in
    host) bash tools/gpu_session.sh host; cp gpurun_out/host.txt gpurun_out/host_n$N.txt ;;
    bench)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
      echo "bench N=$N rc=$?"; tail -c 400 gpurun_out/bench_n$N.err ;;
    single)
      timeout 600 python bench.py --gpus $N --single-process --steps 10 --warmup 3 > gpurun_out/bench_single_n$N.json 2> gpurun_out/bench_single_n$N.err
      echo "single-process N=$N rc=$?"; tail -c 400 gpurun_out/bench_single_n$N.err ;;
    sweep) bash tools/e2e_sweep_multi.sh $N ;;
    tests) ( timeout 900 python -m pytest tests/test_multi.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_multi_n$N.log; tail -3 gpurun_out/pytest_multi_n$N.log ;;
  esac
done

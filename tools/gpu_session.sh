#!/bin/bash
# One parameterised GPU session (replaces the per-session scripts of round 1).  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh tests bench sweep'
# Steps (any subset, in the order given):
#   newtests  the round-2 test files only (fast feedback)        -> gpurun_out/pytest_new.log
#   tests     the whole -m gpu suite                             -> gpurun_out/pytest_gpu.log
#   bench     bench.py own arm, N = 1 ($BENCH_ARGS)               -> gpurun_out/bench.json
#   refarm    bench.py --impl reference                          -> gpurun_out/bench_ref.json
#   sweep     tools/e2e_sweep.py ($SWEEP_ARGS)                    -> gpurun_out/e2e_sweep.jsonl
#   launches  ncu launch list of the bench command               -> gpurun_out/launches.csv
#   ncu       ncu --set full of the kernels in $NCU_KERNELS      -> gpurun_out/prof_<name>.ncu-rep
#   host      lscpu / numactl / nvidia-smi topo of the box       -> gpurun_out/host.txt
#   fuzz      tests/test_gpu_fuzz.py with $FUZZ_CASES cases (default 1500) and seed $FUZZ_SEED  -> gpurun_out/fuzz.log
#   sanitize  compute-sanitizer ($SAN_TOOLS) over tools/sanitize_cases.py ($SAN_CASES) -> gpurun_out/sanitize_<tool>.log
#   refbench  the reference's own benchmark programs, unmodified, on this library (oracle/_ref/reftests/bench_{gpu,xlib,cpu};
#             $REFBENCH_FILTER = substring of the benchmark names, FSS_BENCH_ITERS iterations) -> gpurun_out/refbench_<name>.txt
set -u
mkdir -p gpurun_out
for step in "$@"; do
  case "$step" in
    host)
      { lscpu; echo; numactl -H 2>/dev/null || cat /sys/devices/system/node/node*/cpulist; echo; nvidia-smi topo -m; echo; free -g; nproc; } > gpurun_out/host.txt 2>&1 ;;
    newtests)
      ( timeout 1200 python -m pytest tests/test_host_pipeline.py tests/test_multi.py tests/test_plugins.py tests/test_ref_gtests.py -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_new.log
      tail -5 gpurun_out/pytest_new.log ;;
    tests)
      ( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
      tail -5 gpurun_out/pytest_gpu.log ;;
    bench)
      timeout 900 python bench.py ${BENCH_ARGS:---steps 20 --warmup 5} > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "bench rc=$?"; tail -c 600 gpurun_out/bench.err ;;
    refarm)
      timeout 900 python bench.py --impl reference ${REF_ARGS:---steps 5 --warmup 3} > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
      echo "refarm rc=$?" ;;
    sweep)
      timeout 900 python tools/e2e_sweep.py ${SWEEP_ARGS:-} > gpurun_out/e2e_sweep.log 2>&1
      echo "sweep rc=$?"; tail -40 gpurun_out/e2e_sweep.log ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu --no-check > gpurun_out/launches_bench.log 2>&1
      echo "launches rc=$?" ;;
    ncu)
      for k in ${NCU_KERNELS:-c2}; do
        case $k in c2|c3|ht|packed|vdpf|walk|lm) rx=point_kernel ;; gen_*) rx=gen_kernel ;; relayout) rx=relayout_kernel ;; evalall_dcf) rx=dcf_evalall ;; *) rx=evalall_kernel ;; esac
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:${NCU_REGEX:-$rx} -s ${NCU_SKIP:-2} -c 1 \
          -f -o gpurun_out/prof_$k python tools/prof_one.py $k > gpurun_out/prof_$k.log 2>&1
        echo "ncu $k rc=$?"
      done ;;
    fuzz)
      ( FSSB200_FUZZ_CASES=${FUZZ_CASES:-1500} FSSB200_FUZZ_SEED=${FUZZ_SEED:-7} timeout ${FUZZ_TIMEOUT:-1200} python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/fuzz.log
      tail -6 gpurun_out/fuzz.log ;;
    sanitize)
      for tool in ${SAN_TOOLS:-memcheck racecheck}; do
        timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py ${SAN_CASES:-walk pipeline packed} \
          > gpurun_out/sanitize_$tool.log 2>&1
        echo "sanitize $tool rc=$?"; tail -4 gpurun_out/sanitize_$tool.log
      done ;;
    refbench)
      for b in bench_gpu bench_xlib bench_cpu; do
        [ -x oracle/_ref/reftests/$b ] || { echo "$b not built (make -C oracle reftests, needs the reference checkout)"; continue; }
        FSS_BENCH_ITERS=${FSS_BENCH_ITERS:-10} timeout 600 oracle/_ref/reftests/$b ${REFBENCH_FILTER:-} > gpurun_out/refbench_$b.txt 2>&1
        echo "refbench $b rc=$?"; tail -5 gpurun_out/refbench_$b.txt
      done ;;
    *) echo "unknown step $step" ;;
  esac
done

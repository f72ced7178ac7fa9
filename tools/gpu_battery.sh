#!/usr/bin/env bash
# One-GPU validation + measurement battery (run from the repo root on a B200 box, e.g.
#   gpurun --timeout 1500 -- 'bash tools/gpu_battery.sh' ).  Every step has its own timeout and log under
# gpurun_out/; a failing step does not stop the later ones.  Order = cheapest and most informative first.
set -u
mkdir -p gpurun_out
run() {  # run <seconds> <log> <command...>
  local t=$1 log=$2; shift 2
  echo "=== $* (timeout ${t}s) -> gpurun_out/$log"
  timeout "$t" "$@" > "gpurun_out/$log" 2>&1
  echo "    rc=$?"
}
# 1. torch-free: every kernel variant against the others bit for bit, with timings.  First the part that has already
#    run green on a B200 (round 1), then each arm written without a GPU at hand in its OWN process, so that a fault
#    in one cannot hide the others (SELFTEST_ONLY / SELFTEST_SKIP: see tools/cabi_selftest.cu).
SELFTEST_SKIP=deep,prefetch,lean,streamscreen,host,exchange run 180 selftest_core.log tools/cabi_selftest 1000000 1250000 10000000
grep -E "MISMATCH|selftest|search time" gpurun_out/selftest_core.log | tail -12
for arm in deep prefetch lean streamscreen host exchange; do
  only=$arm
  case $arm in lean|streamscreen) only="$arm,sweep";; esac      # these two also have arms in the shape sweep
  SELFTEST_ONLY=$only run 120 "selftest_$arm.log" tools/cabi_selftest
  grep -E "MISMATCH|selftest|search time|batch-|end to end|CUDA error|mmf error" "gpurun_out/selftest_$arm.log" | tail -8
done
# 2. does 1 KB of every 2 KB cost HBM efficiency?
run 60 hbm_stride.log tools/hbm_stride_micro
cat gpurun_out/hbm_stride.log
# 3. the parity suites
run 900 pytest_gpu.log python -m pytest tests -m gpu -x -q
tail -3 gpurun_out/pytest_gpu.log
MMF_EXPERIMENTAL=1 run 600 pytest_experimental.log python -m pytest tests/test_gpu_experimental.py -q
tail -3 gpurun_out/pytest_experimental.log
# 4. bench lines: default workload with both e2e paths, batch-1 latency mode, 10M-row bf16 on one GPU
run 300 bench_c2.json python bench.py
run 200 bench_c2_e2e_host.json python bench.py --e2e-api host --graph --no-cpu-baseline
run 200 bench_c3.json python bench.py --workload c3 --no-cpu-baseline
run 400 bench_c4_n1.json python bench.py --workload c4 --no-cpu-baseline --steps 10
run 120 bench_reference.json python bench.py --impl reference --steps 5 --warmup 1
run 400 bench_c5.json python tools/bench_c5.py --autocast
for f in bench_c2 bench_c2_e2e_host bench_c3 bench_c4_n1 bench_c5; do tail -c 1200 "gpurun_out/$f.json"; echo; done
# 5. A/B of the experiments through the Python API
run 300 ab_experimental.log python tools/ab_experimental.py
cat gpurun_out/ab_experimental.log
# 6. ncu (B200_PROFILING.md: only after the SAME command line has exited 0 without ncu, `&&` directly before it):
#    launch list of the default bench command, full capture of the dominant kernels
echo "=== ncu launch list + full capture"
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain_bench.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/c2_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "    rc=$?"
timeout 60 tools/cabi_selftest profile > gpurun_out/ncu_plain_selftest.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:vault_mma_topk -c 3 -f \
  -o gpurun_out/c2_c4shard_full tools/cabi_selftest profile > gpurun_out/ncu_full.log 2>&1
echo "    rc=$?"

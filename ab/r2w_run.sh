#!/bin/bash
# round 2, call w: extract_aligned_kernel for single sequences with a partial last group (C4); compute-sanitizer on the new kernels
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest.log); tail -3 gpurun_out/r2w_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2 3; do
  python tools/bench_configs.py --cases c4 --steps 10 2>/dev/null | line aligned
  KMC_ALIGNED_KERNEL=0 python tools/bench_configs.py --cases c4 --steps 10 2>/dev/null | line generic
done
python bench.py --no-cpu --no-e2e --no-check --no-sustained --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 leg', d['c4']['ms_per_step'], d['c4']['achieved_GBps_per_gpu'])"
ncu --set full --clock-control none --import-source on -k regex:extract_ -c 2 -o gpurun_out/r2w_c4 -f python tools/bench_configs.py --cases c4 --steps 1 --warmup 0 > gpurun_out/r2w_ncu.log 2>&1
(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aligned or fused_bins or test_bucket_count or digest_fused or uniform_read_set" > gpurun_out/r2w_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r2w_memcheck.log); tail -4 gpurun_out/r2w_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_bins or test_bucket_count or digest_fused" > gpurun_out/r2w_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/r2w_racecheck.log); tail -4 gpurun_out/r2w_racecheck.log

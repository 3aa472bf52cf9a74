#!/bin/bash
# round 2, call p: bucket_bin_kernel (ids + 64-way binning in one kernel, atomic cursors, fixed bin capacity with exact fallback);
# resident-block cap of extract_aligned_kernel (KMC_ALIGNED_SMEM) on the HBM-bound modes
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_sketch.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest.log); tail -3 gpurun_out/r2p_pytest.log
(KMC_FUSED_BIN=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k bucket > gpurun_out/r2p_pytest_exact.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest_exact.log); tail -3 gpurun_out/r2p_pytest_exact.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line fused
  KMC_FUSED_BIN=0 python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line exact
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_c5_launches.csv python tools/bench_configs.py --cases c5 --steps 1 --warmup 1 > /dev/null 2>&1
for smem in 0 40000 56000 75000 110000; do
  KMC_ALIGNED_SMEM=$smem python tools/bench_configs.py --cases modes --steps 10 2>/dev/null | line smem$smem
  KMC_ALIGNED_SMEM=$smem python bench.py --no-legs --no-cpu --no-e2e --no-check --steps 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('smem$smem bench', d['ms_per_step'], d['value']/1e9, r['frac'], r.get('frac_sustained'), r.get('frac_burst'), d['clocks'])"
done
ncu --set full --clock-control none --import-source on -k regex:bucket_bin -s 1 -c 1 -o gpurun_out/r2p_c5_binkernel -f python tools/bench_configs.py --cases c5 --steps 1 --warmup 1 > gpurun_out/r2p_ncu.log 2>&1
ls -la gpurun_out/r2p_*.ncu-rep

#!/bin/bash
# round 2, call q: bucket_bin_kernel v2 (full iterations branch-free, dump area instead of a per-key test), exact-path scatter likewise;
# extract_aligned_kernel capped at 3 blocks per SM by default; launch list of the ragged case
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_sketch.py tests/test_gpu_kmer4.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2q_pytest.log); tail -3 gpurun_out/r2q_pytest.log
(KMC_FUSED_BIN=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k bucket > gpurun_out/r2q_pytest_exact.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2q_pytest_exact.log); tail -3 gpurun_out/r2q_pytest_exact.log
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
python tools/bench_configs.py --cases count --steps 5 2>/dev/null | line count
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_c5_launches.csv python tools/bench_configs.py --cases c5 --steps 1 --warmup 1 > /dev/null 2>&1
KMC_FUSED_BIN=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_c5_launches_exact.csv python tools/bench_configs.py --cases c5 --steps 1 --warmup 1 > /dev/null 2>&1
for rep in 1 2; do
  python tools/bench_configs.py --cases modes,c4,kmer4 --steps 10 2>/dev/null | line cap3
  KMC_ALIGNED_KERNEL=0 python tools/bench_configs.py --cases c4 --steps 10 2>/dev/null | line generic
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q_ragged_launches.csv python tools/bench_configs.py --cases ragged --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bucket_bin -s 1 -c 1 -o gpurun_out/r2q_c5_binkernel -f python tools/bench_configs.py --cases c5 --steps 1 --warmup 1 > gpurun_out/r2q_ncu.log 2>&1
ls -la gpurun_out/r2q_*.ncu-rep

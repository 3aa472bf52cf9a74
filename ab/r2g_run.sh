#!/bin/bash
# round 2, call g: full GPU suite; C3 with the branch-free two-step compaction (v6); the TMA bulk-store A/B on the C2 kernel
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log); tail -6 gpurun_out/r2g_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d['frac_of_measured_peak'],3), d['case'][:90])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lin_v6
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lin_|recode|sums|rescan|compact|valid" -c 40 --csv --log-file gpurun_out/r2g_c3_launches.csv python tools/bench_configs.py --cases c3 --steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lin_compact -s 5 -c 1 -o gpurun_out/r2g_c3_lin_aos -f python tools/bench_configs.py --cases c3 --steps 1 > gpurun_out/r2g_ncu.log 2>&1
# TMA bulk-store variant of the C2 kernel: correctness, then the A/B (alternating processes), then one ncu capture each
(KMERSCUDA_LIB=$PWD/ab/libk_tma.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "uniform or full_size or read_set or c2" > gpurun_out/r2g_tma_pytest.log 2>&1; echo "tma pytest exit $?" >> gpurun_out/r2g_tma_pytest.log); tail -3 gpurun_out/r2g_tma_pytest.log
python ab/ab.py $PWD/kmers.jl_b200/libkmerscuda.so $PWD/ab/libk_tma.so > gpurun_out/r2g_ab_tma.txt 2>&1; cat gpurun_out/r2g_ab_tma.txt
ncu --set full --clock-control none -k regex:extract_kernel -s 3 -c 1 -o gpurun_out/r2g_c2_base -f python bench.py --no-legs --no-cpu --no-e2e --no-sustained --no-check --steps 3 > /dev/null 2>&1
KMERSCUDA_LIB=$PWD/ab/libk_tma.so ncu --set full --clock-control none -k regex:extract_kernel -s 3 -c 1 -o gpurun_out/r2g_c2_tma -f python bench.py --no-legs --no-cpu --no-e2e --no-sustained --no-check --steps 3 > /dev/null 2>&1
ls -la gpurun_out/r2g_*.ncu-rep

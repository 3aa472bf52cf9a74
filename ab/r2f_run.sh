#!/bin/bash
# round 2, call d: all GPU tests (incl. spaced k-mers, ASCII -> 4-bit), C3 A/B with conflict-free staging, ncu of the
# SoA and AoS launches of lin_compact_kernel, the launch list of one C3 call
mkdir -p gpurun_out
(python -m pytest tests/test_gpu_fourbit.py tests/test_gpu_ascii.py tests/test_gpu_lincompact.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log); tail -15 gpurun_out/r2f_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d['frac_of_measured_peak'],3), d['case'][:90])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lin_v5
  KMC_LINEAR=0 python tools/bench_configs.py --cases c3 --steps 10 2>/dev/null | line lookback
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lin_|recode|sums|rescan|compact|valid" -c 40 --csv --log-file gpurun_out/r2f_c3_launches.csv python tools/bench_configs.py --cases c3 --steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lin_compact -s 2 -c 1 -o gpurun_out/r2f_c3_lin_soa -f python tools/bench_configs.py --cases c3 --steps 1 > gpurun_out/r2f_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lin_compact -s 5 -c 1 -o gpurun_out/r2f_c3_lin_aos -f python tools/bench_configs.py --cases c3 --steps 1 >> gpurun_out/r2f_ncu.log 2>&1
python tools/bench_configs.py --cases kmer4 --steps 5 2>/dev/null | line kmer4

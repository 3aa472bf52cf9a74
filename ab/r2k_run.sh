#!/bin/bash
# round 2, call k: software-pipelined source-order compaction (64-bit loads, next trip's loads before this trip's stores)
mkdir -p gpurun_out
(python -m pytest tests/test_gpu_lincompact.py tests/test_gpu_fourbit.py tests/test_gpu_ascii.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest.log); tail -6 gpurun_out/r2k_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d['frac_of_measured_peak'],3), d['case'][:90])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line v8
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lin_|recode|sums|rescan|compact|valid|ascii|extract" -c 60 --csv --log-file gpurun_out/r2k_c3_launches.csv python tools/bench_configs.py --cases c3 --steps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lin_|recode|sums|rescan|compact|valid|ascii|extract|first_error|resolve" -c 60 --csv --log-file gpurun_out/r2k_ascii_launches.csv python tools/bench_configs.py --cases ascii --steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lin_compact -s 1 -c 1 -o gpurun_out/r2k_c3_lin_soa -f python tools/bench_configs.py --cases c3 --steps 1 > gpurun_out/r2k_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lin_compact -s 5 -c 1 -o gpurun_out/r2k_c3_lin_aos -f python tools/bench_configs.py --cases c3 --steps 1 >> gpurun_out/r2k_ncu.log 2>&1
ls -la gpurun_out/r2k_*.ncu-rep

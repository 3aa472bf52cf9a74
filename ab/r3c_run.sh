#!/bin/bash
# round 2, call 3c: ASCII recoding through positioned tables (OR of eight pre-placed entries per byte pair); smoke()
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_ascii.py tests/test_gpu_spaced.py tests/test_gpu_fourbit.py tests/test_gpu_lincompact.py tests/test_gpu_kmer4.py -m gpu -x -q > gpurun_out/r3c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3c_pytest.log); tail -3 gpurun_out/r3c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do python tools/bench_configs.py --cases ascii --steps 10 2>/dev/null | line positioned; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lin_|recode|sums|rescan|compact|valid|ascii|extract|first_error|resolve" -c 60 --csv --log-file gpurun_out/r3c_ascii_launches.csv python tools/bench_configs.py --cases ascii --steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ascii_recode -c 1 -o gpurun_out/r3c_ascii_recode -f python tools/bench_configs.py --cases ascii --steps 1 --warmup 0 > gpurun_out/r3c_ncu.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ascii.py -m gpu -x -q > gpurun_out/r3c_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r3c_memcheck.log); tail -3 gpurun_out/r3c_memcheck.log

#!/bin/bash
# round 2, call 3d: strict iteration over recoded sources: the valid-start bits prefetched in the same bursts as the source
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_ascii.py tests/test_gpu_fourbit.py tests/test_gpu_spaced.py -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3d_pytest.log); tail -3 gpurun_out/r3d_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do python tools/bench_configs.py --cases ascii --steps 10 2>/dev/null | line vprefetch; KMC_PREFETCH=0 python tools/bench_configs.py --cases ascii --steps 10 2>/dev/null | line noprefetch; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ascii|extract" -c 20 --csv --log-file gpurun_out/r3d_ascii_launches.csv python tools/bench_configs.py --cases ascii --steps 1 > /dev/null 2>&1

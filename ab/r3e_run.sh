#!/bin/bash
# round 2, call 3e: the fused bucket count with a spill list instead of the overflow flag + fallback (no host wait any more; pieces removed)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3e_pytest.log); tail -3 gpurun_out/r3e_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line spill; done
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bucket" > gpurun_out/r3e_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r3e_memcheck.log); tail -3 gpurun_out/r3e_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_bins" > gpurun_out/r3e_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/r3e_racecheck.log); tail -3 gpurun_out/r3e_racecheck.log

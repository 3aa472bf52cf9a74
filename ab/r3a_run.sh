#!/bin/bash
# round 2, call 3a: dense minimizers templated on the exact W; mul_c64
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_minimizers.py tests/test_gpu_sketch.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3a_pytest.log); tail -3 gpurun_out/r3a_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do python tools/bench_configs.py --cases minimizers --steps 10 2>/dev/null | line exactW; done
ncu --set full --clock-control none --import-source on -k regex:minimizer_dense -c 1 -o gpurun_out/r3a_minimizers -f python tools/bench_configs.py --cases minimizers --steps 1 --warmup 0 > gpurun_out/r3a_ncu.log 2>&1

#!/bin/bash
# round 2, call 3g: the consumers (MinHash sketch, composition, exact count table) through the aligned work items
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_minimizers.py -m gpu -x -q > gpurun_out/r3g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3g_pytest.log); tail -3 gpurun_out/r3g_pytest.log
(KMC_ALIGNED_KERNEL=0 timeout 900 python -m pytest tests/test_gpu_sketch.py -m gpu -x -q > gpurun_out/r3g_pytest_generic.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3g_pytest_generic.log); tail -3 gpurun_out/r3g_pytest_generic.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do python tools/bench_configs.py --cases sketch,count --steps 10 2>/dev/null | line aligned; KMC_ALIGNED_KERNEL=0 python tools/bench_configs.py --cases sketch,count --steps 10 2>/dev/null | line generic; done

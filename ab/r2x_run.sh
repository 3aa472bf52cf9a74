#!/bin/bash
# round 2, call x: after the fix of the general kernel's aligned path for partial tails: full suite, memcheck again, strict 4-bit / ASCII
# sources through extract_aligned_kernel<STRICT4>
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_pytest.log); tail -3 gpurun_out/r2x_pytest.log
(KMC_ALIGNED_KERNEL=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fourbit.py tests/test_gpu_ascii.py -m gpu -x -q > gpurun_out/r2x_pytest_generic.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_pytest_generic.log); tail -3 gpurun_out/r2x_pytest_generic.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases ascii,c4 --steps 10 2>/dev/null | line aligned
  KMC_ALIGNED_KERNEL=0 python tools/bench_configs.py --cases ascii,c4 --steps 10 2>/dev/null | line generic
done
(timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_fourbit.py tests/test_gpu_ascii.py -m gpu -x -q -k "not full_size and not host_path_single" > gpurun_out/r2x_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r2x_memcheck.log); tail -4 gpurun_out/r2x_memcheck.log

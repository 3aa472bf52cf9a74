#!/bin/bash
# round 2, call 3h: the Julia tuple layouts (Tuple{Kmer,Kmer}, Tuple{Kmer,Int}) of one-limb k-mers through the lean kernel with groups of two windows
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3h_pytest.log); tail -3 gpurun_out/r3h_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'FwRv' in d['case'] or 'Unambig' in d['case']: print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do python tools/bench_configs.py --cases modes --steps 10 2>/dev/null | line aos2; KMC_ALIGNED_KERNEL=0 python tools/bench_configs.py --cases modes --steps 10 2>/dev/null | line generic; done
ncu --set full --clock-control none --import-source on -k regex:extract_aligned -s 4 -c 1 -o gpurun_out/r3h_aos -f python tools/bench_configs.py --cases modes --steps 1 --warmup 0 > gpurun_out/r3h_ncu.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aligned or single_sequence_all_modes" > gpurun_out/r3h_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r3h_memcheck.log); tail -3 gpurun_out/r3h_memcheck.log

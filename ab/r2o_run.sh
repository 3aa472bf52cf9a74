#!/bin/bash
# round 2, call o: extract_aligned_kernel (lean kernel for aligned uniform sets), 3-instruction fx multiply, half-width head mask:
# parity with both kernels, A/B of the two kernels in alternating processes, bench.py, ncu of the new kernel
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log); tail -3 gpurun_out/r2o_pytest.log
(KMC_ALIGNED_KERNEL=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kmer4.py -m gpu -x -q > gpurun_out/r2o_pytest_generic.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest_generic.log); tail -3 gpurun_out/r2o_pytest_generic.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases modes,c4,kmer4 --steps 10 2>/dev/null | line aligned
  KMC_ALIGNED_KERNEL=0 python tools/bench_configs.py --cases modes,c4,kmer4 --steps 10 2>/dev/null | line generic
done
python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line aligned
KMC_ALIGNED_KERNEL=0 python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line generic
for rep in 1 2; do
  python bench.py --no-legs --no-cpu --no-e2e --no-check --steps 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('aligned bench', d['ms_per_step'], d['value']/1e9, r['frac'], r.get('frac_sustained'), r.get('frac_burst'), d['clocks'])"
  KMC_ALIGNED_KERNEL=0 python bench.py --no-legs --no-cpu --no-e2e --no-check --steps 100 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('generic bench', d['ms_per_step'], d['value']/1e9, r['frac'], r.get('frac_sustained'), r.get('frac_burst'), d['clocks'])"
done
timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -c 1500 gpurun_out/r2o_bench.json
ncu --set full --clock-control none --import-source on -k regex:extract_aligned -s 3 -c 1 -o gpurun_out/r2o_c2_aligned -f python bench.py --no-legs --no-cpu --no-e2e --no-sustained --no-check --steps 3 > gpurun_out/r2o_ncu.log 2>&1
ls -la gpurun_out/r2o_*.ncu-rep

#!/bin/bash
# round 2, call 3j: SoA index stream as a template parameter of the lean kernel; full suite; bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3j_pytest.log); tail -3 gpurun_out/r3j_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for rep in 1 2; do python tools/bench_configs.py --cases modes --steps 10 2>/dev/null | line lean; done
timeout 900 python bench.py > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r3j_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('bench', d['ms_per_step'], d['value']/1e9, r['frac'], r['frac_sustained'], r['frac_burst'], 'e2e', d['e2e']['value']/1e9, 'c4', d['c4']['ms_per_step'], 'c5', d['c5']['count_ms'])"

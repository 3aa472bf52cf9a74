#!/bin/bash
# round 2, call 3n: final build (AlignedLocator moved to kmer_core.cuh): full GPU suite, smoke, short bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3n_pytest.log); tail -3 gpurun_out/r3n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu --steps 50 > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r3n_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('bench', d['ms_per_step'], d['value']/1e9, r['frac'], r['frac_sustained'], r['frac_burst'], 'e2e', d['e2e']['value']/1e9, 'c4', d['c4']['ms_per_step'], 'c5', d['c5']['count_ms'])"

#!/bin/bash
# round 2, call n (re-entry): full GPU suite + both bench arms + every bench_configs case on the state of commit db174ef
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log); tail -4 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err; tail -c 600 gpurun_out/r2n_bench_ref.json
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 3000 gpurun_out/r2n_bench.json
timeout 900 python tools/bench_configs.py --cases c3,c3long,c4,c5,modes,ragged,ascii,minimizers,sketch,kmer4,count --steps 10 > gpurun_out/r2n_configs.jsonl 2> gpurun_out/r2n_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2n_configs.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])
PY

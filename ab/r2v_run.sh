#!/bin/bash
# round 2, call v: full GPU suite + both bench arms + every bench_configs case (aligned kernel, fused bucket bins)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2v_pytest.log); tail -4 gpurun_out/r2v_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err; tail -c 600 gpurun_out/r2v_bench_ref.json
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -c 3000 gpurun_out/r2v_bench.json
timeout 900 python tools/bench_configs.py --cases c3,c3long,c4,c5,modes,ragged,ascii,minimizers,sketch,kmer4,count --steps 10 > gpurun_out/r2v_configs.jsonl 2> gpurun_out/r2v_configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2v_configs.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:extract_aligned -s 3 -c 1 -o gpurun_out/r2v_c2_aligned -f python bench.py --no-legs --no-cpu --no-e2e --no-sustained --no-check --steps 3 > gpurun_out/r2v_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:extract_ -c 2 -o gpurun_out/r2v_c4 -f python tools/bench_configs.py --cases c4 --steps 1 --warmup 0 >> gpurun_out/r2v_ncu.log 2>&1
KMC_ALIGNED_KERNEL=0 ncu --set full --clock-control none --import-source on -k regex:extract_ -c 1 -o gpurun_out/r2v_c4_generic -f python tools/bench_configs.py --cases c4 --steps 1 --warmup 0 >> gpurun_out/r2v_ncu.log 2>&1
ls -la gpurun_out/r2v_*.ncu-rep

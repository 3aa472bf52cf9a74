#!/bin/bash
# one GPU call of round 2: tests, the C3 A/B, an ncu capture of the source-order compaction, the H2D probe, C2 bench, modes
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log); tail -15 gpurun_out/r2c_pytest.log
bash ab/r2c_c3.sh > gpurun_out/r2c_c3_ab.txt 2>&1; cat gpurun_out/r2c_c3_ab.txt
ncu --set full --clock-control none --import-source on -k regex:lin_compact -s 2 -c 2 -o gpurun_out/r2c_c3_lin -f python tools/bench_configs.py --cases c3 --steps 1 > gpurun_out/r2c_ncu.log 2>&1
python tools/h2d_probe.py > gpurun_out/r2c_h2d_1gpu.json 2> gpurun_out/r2c_h2d.err; tail -5 gpurun_out/r2c_h2d.err
python bench.py --no-legs --no-cpu --steps 100 > gpurun_out/r2c_bench.json 2>gpurun_out/r2c_bench.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2c_bench.json') if l.startswith('{')][0]
print('C2 value', d['value'], 'ms', d['ms_per_step'], 'sus', d['roofline']['sustained']['kernel_ms'], 'gap', d['roofline']['idle_gaps']['kernel_ms'], 'e2e', d['e2e']['value'])"
python tools/bench_configs.py --cases modes,c4 --steps 10 2>/dev/null > gpurun_out/r2c_modes.jsonl
python -c "
import json
for l in open('gpurun_out/r2c_modes.jsonl'):
    d=json.loads(l); print(round(d['ms_median'],3), round(d['frac_of_measured_peak'],3), d['case'][:90])"

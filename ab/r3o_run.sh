#!/bin/bash
# round 2, call 3o: ncu launch list of bench.py on the final build; one full capture of a bin_apply_fused_kernel launch
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3o_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bin_apply_fused -s 10 -c 1 -o gpurun_out/r3o_c5_apply -f python tools/bench_configs.py --cases c5 --steps 1 --warmup 0 > gpurun_out/r3o_ncu.log 2>&1
ls -la gpurun_out/r3o_*

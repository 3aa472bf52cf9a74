#!/bin/bash
# round 2, call 3p: ncu captures of the lean kernel's launches in the "modes" case (FwKmers, FwRv SoA, FwRv AoS with groups of two, ...)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:extract_aligned -c 6 -o gpurun_out/r3p_modes -f python tools/bench_configs.py --cases modes --steps 1 --warmup 1 > gpurun_out/r3p_ncu.log 2>&1
tail -3 gpurun_out/r3p_ncu.log; ls -la gpurun_out/r3p_*

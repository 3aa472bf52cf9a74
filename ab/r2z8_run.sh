#!/bin/bash
# round 2, call z8: the C5 job from plain C on eight GPUs (kmc_group_bucket_count, one host thread per device for the counts)
mkdir -p gpurun_out
gcc -std=c99 -O2 -Iinclude examples/c5_group_count.c -Lkmers.jl_b200 -lkmerscuda -Wl,-rpath,$PWD/kmers.jl_b200 -o /tmp/c5_group_count && /tmp/c5_group_count 25000000 28 > gpurun_out/r2z8_c5_example.txt 2>&1; cat gpurun_out/r2z8_c5_example.txt

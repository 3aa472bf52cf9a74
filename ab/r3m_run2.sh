#!/bin/bash
# round 2, call 3m: the C5 job from plain C and the multi-GPU tests on two GPUs; counts enqueued from one host thread per device again
mkdir -p gpurun_out
gcc -std=c99 -O2 -Iinclude examples/c5_group_count.c -Lkmers.jl_b200 -lkmerscuda -Wl,-rpath,$PWD/kmers.jl_b200 -o /tmp/c5_group_count && /tmp/c5_group_count 25000000 28 > gpurun_out/r3m_c5_example.txt 2>&1; cat gpurun_out/r3m_c5_example.txt
(python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3m_pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3m_pytest_multi.log); tail -3 gpurun_out/r3m_pytest_multi.log

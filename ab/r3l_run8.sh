#!/bin/bash
# round 2, third 8-GPU call: final build (spill-list bucket count, tuple-layout kernels): bench.py at N = 8 (both arms: C2 weak scaling, e2e, the c4 / c5 legs), the host -> device ceiling of the
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r3l_topo8.txt 2>&1; lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name" >> gpurun_out/r3l_topo8.txt; free -g | head -2 >> gpurun_out/r3l_topo8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --impl reference --steps 5 --warmup 1 > gpurun_out/r3l_ref_n8.json 2> gpurun_out/r3l_ref_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3l_bench_n8.json 2> gpurun_out/r3l_bench_n8.err
echo "bench exit $?"; tail -3 gpurun_out/r3l_bench_n8.err
python - <<'PY'
import json
for f in ("gpurun_out/r3l_ref_n8.json", "gpurun_out/r3l_bench_n8.json"):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][0]
        print(f, d.get("impl","ours"), "value", d["value"], "cores", d.get("cpu_baseline",{}).get("cores"), "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("h2d_GBps_per_gpu"))
        for k in ("c4","c5"):
            if k in d: print(k, {x: d[k][x] for x in d[k] if x not in ("workload","parity","collective")})
    except Exception as e: print(f, "ERR", e)
PY
gcc -std=c99 -O2 -Iinclude examples/c5_group_count.c -Lkmers.jl_b200 -lkmerscuda -Wl,-rpath,$PWD/kmers.jl_b200 -o /tmp/c5_group_count && /tmp/c5_group_count 25000000 28 > gpurun_out/r3l_c5_example.txt 2>&1; cat gpurun_out/r3l_c5_example.txt

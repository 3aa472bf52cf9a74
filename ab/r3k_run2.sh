#!/bin/bash
# round 2, call 3k: bench.py at N = 2 (both arms) on the final build
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r3k_ref_n2.json 2> gpurun_out/r3k_ref_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3k_bench_n2.json 2> gpurun_out/r3k_bench_n2.err
echo "bench exit $?"; tail -3 gpurun_out/r3k_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r3k_ref_n2.json", "gpurun_out/r3k_bench_n2.json"):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][0]
        print(f, d.get("impl","ours"), "value", d["value"], "cores", d.get("cpu_baseline",{}).get("cores"), "e2e", d.get("e2e",{}).get("value"))
        for k in ("c4","c5"):
            if k in d: print(k, {x: d[k][x] for x in d[k] if x not in ("workload","parity","collective")})
    except Exception as e: print(f, "ERR", e)
PY

#!/bin/bash
# round 2, call s: L2 hints (binned keys evict_first, table slices and increments evict_last), pieces again
mkdir -p gpurun_out
(KMC_BIN_PIECES=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_sketch.py -m gpu -x -q -k "bucket or multi or group or merge or table or count" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2s_pytest.log); tail -3 gpurun_out/r2s_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'B=28' in d['case'] or 'exact' in d['case']: print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for P in 1 2 3 4; do
  KMC_BIN_PIECES=$P python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line pieces$P
done
for cfg in "128 8" "256 4" "256 8"; do
  set -- $cfg
  for P in 2 4; do
  KMC_BIN_PIECES=$P KMC_APPLY_BLOCK=$1 KMC_APPLY_GRID=$2 python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line "pieces${P}_block$1_persm$2"
  done
done
KMC_FUSED_BIN=0 python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line exact
python tools/bench_configs.py --cases count --steps 5 2>/dev/null | line count
KMC_BIN_PIECES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2s_c5_launches_p1.csv python tools/bench_configs.py --cases c5 --steps 1 --warmup 1 > /dev/null 2>&1

#!/bin/bash
# round 2, call r: the fused bucket count in pieces -- the increments of piece i on a high-priority stream beside the binning of piece i + 1
mkdir -p gpurun_out
(KMC_BIN_PIECES=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q -k "bucket or multi or group or merge" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2r_pytest.log); tail -3 gpurun_out/r2r_pytest.log
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'B=28' in d['case']: print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for P in 1 2 3 4 6 8; do
  KMC_BIN_PIECES=$P python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line pieces$P
done
for cfg in "64 4" "64 8" "128 2" "128 8" "256 2" "256 4"; do
  set -- $cfg
  KMC_BIN_PIECES=4 KMC_APPLY_BLOCK=$1 KMC_APPLY_GRID=$2 python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line "pieces4_block$1_persm$2"
done
KMC_BIN_PIECES=4 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_c5_launches_p4.csv python tools/bench_configs.py --cases c5 --steps 1 --warmup 1 > /dev/null 2>&1

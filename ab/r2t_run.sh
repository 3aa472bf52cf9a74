#!/bin/bash
# round 2, call t: apply / warm kernels under the binning kernel's shared-memory carve-out (can two kernels share an SM?)
mkdir -p gpurun_out
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'B=28' in d['case'] or 'exact' in d['case']: print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d.get('frac_of_measured_peak',0),3), d['case'][:100])"; }
for P in 1 2 4; do
  KMC_BIN_PIECES=$P python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line pieces$P
done
for cfg in "128 8" "128 2" "64 8" "256 4"; do
  set -- $cfg
  for P in 2 4; do
  KMC_BIN_PIECES=$P KMC_APPLY_BLOCK=$1 KMC_APPLY_GRID=$2 python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line "pieces${P}_block$1_persm$2"
  done
done

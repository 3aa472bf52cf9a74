#!/bin/bash
# round 2, C3: the source-order compaction (lin_compact_kernel) against the look-back kernel, register variants, prefetch
cd "$(dirname "$0")/.."
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d['frac_of_measured_peak'],3), d['case'][:90])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lin4
  KMERSCUDA_LIB=$PWD/ab/libk_lin3.so python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lin3
  KMERSCUDA_LIB=$PWD/ab/libk_lin5.so python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lin5
  KMC_PREFETCH=0 python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lin4_nopf
  KMC_LINEAR=0 python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lookback
done

#!/bin/bash
# round 2, C3: source-order compaction v2 (packed positions, shared staging addresses, cheap read-out) vs the look-back kernel
cd "$(dirname "$0")/.."
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['ms_min'],3), round(d['frac_of_measured_peak'],3), d['case'][:90])"; }
for rep in 1 2; do
  python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lin_v2
  KMC_LINEAR=0 python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line lookback
done

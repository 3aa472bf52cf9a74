#!/bin/bash
# round 2, call 3b: bins applied per launch, again, now that table slices are evict_last and the ids evict_first
mkdir -p gpurun_out
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'B=28' in d['case']: print('$1', round(d['ms_median'],3), round(d['ms_min'],3), d['case'][:60])"; }
for g in 1 2 3 4 6 8; do KMC_APPLY_GROUP=$g python tools/bench_configs.py --cases c5 --steps 5 2>/dev/null | line group$g; done

run() { # label env...
  label=$1; shift
  env "$@" python tools/bench_configs.py --cases modes,c4,ragged,kmer4,ascii,c3 --steps 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('$label', round(d['ms_median'],3), round(d['frac_of_measured_peak'],3), d['case'][:80])
"
}
run base KMC_PREFETCH=0 KMERSCUDA_LIB=$PWD/ab/libk_noef.so
run ef KMC_PREFETCH=0
run pf_noef KMC_PREFETCH=1 KMERSCUDA_LIB=$PWD/ab/libk_noef.so
run pf_ef KMC_PREFETCH=1

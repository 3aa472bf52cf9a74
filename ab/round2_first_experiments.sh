#!/bin/bash
# The first GPU experiments queued for the next round (none of them has been measured yet).
#   bash ab/round2_first_experiments.sh build     # here (CPU): build the variant libraries into ab/
#   gpurun -- 'bash ab/round2_first_experiments.sh run > gpurun_out/round2_ab.txt 2>&1'
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
case "$1" in
build)
    # 1. burst prefetch for the UnambiguousKmers compaction kernel (C3), DESIGN.md 8 item 3
    make -C kmers.jl_b200/csrc -j8 OBJDIR=/tmp/build_cpf LIB="$ROOT/ab/libk_compact_prefetch.so" EXTRA=-DKMC_COMPACT_PREFETCH=1
    ;;
run)
    line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$1', round(d['ms_median'],3), round(d['frac_of_measured_peak'],3), d['case'][:80])"; }
    for rep in 1 2; do
        python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line base
        KMERSCUDA_LIB="$ROOT/ab/libk_compact_prefetch.so" python tools/bench_configs.py --cases c3,c3long,ascii --steps 10 2>/dev/null | line compact_prefetch
    done
    # 2. how much of the e2e step is the DMA competing with the kernels: the same step with the prefetch off
    python bench.py --no-cpu --steps 20 --host-steps 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']['ms_per_step'], 'value', d['value'])"
    KMC_PREFETCH=0 python bench.py --no-cpu --steps 20 --host-steps 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e prefetch off', d['e2e']['ms_per_step'], 'value', d['value'])"
    ;;
*)
    echo "usage: $0 build|run"; exit 2 ;;
esac

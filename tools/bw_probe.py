"""Device bandwidth probes: the write-only ceiling (kmc_store_probe: 256-bit streaming stores,
nothing read) next to torch's copy (what MEASURED_PEAKS.json hbm_gbs is) and memset.
Usage (GPU box): python tools/bw_probe.py > gpurun_out/bw_probe.json"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmers.jl_b200"))
import torch  # noqa: E402

import kmerscuda as kc  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ctx = kc.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    nbytes = 19_200_000_000
    buf = torch.empty(nbytes // 8, dtype=torch.int64, device="cuda")
    out = {}
    ms = timed(lambda: ctx._check(ctx.lib.kmc_store_probe(ctx.handle, buf.data_ptr(), nbytes)))
    out["store_probe_256bit_GBps"] = nbytes / ms / 1e6
    ms = timed(lambda: buf.zero_())
    out["torch_memset_GBps"] = nbytes / ms / 1e6
    half = buf.numel() // 2
    ms = timed(lambda: buf[:half].copy_(buf[half:2 * half]))
    out["torch_copy_read_plus_write_GBps"] = 2 * half * 8 / ms / 1e6
    # host -> device over PCIe from pinned memory: the ceiling of bench.py's e2e leg (400 MB per step)
    h = torch.empty(400_000_000 // 8, dtype=torch.int64).pin_memory()
    d = torch.empty_like(h, device="cuda")
    ms = timed(lambda: d.copy_(h, non_blocking=True))
    out["h2d_pinned_400MB_GBps"] = h.numel() * 8 / ms / 1e6
    for mb in (10, 32, 100):
        step = mb * 1_000_000 // 8

        def chunks():
            for o in range(0, h.numel(), step):
                d[o:o + step].copy_(h[o:o + step], non_blocking=True)
        ms = timed(chunks)
        out["h2d_pinned_400MB_in_%dMB_copies_GBps" % mb] = h.numel() * 8 / ms / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()

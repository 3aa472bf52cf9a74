#!/usr/bin/env python
"""Host -> device ceiling of the box with N GPUs copying at once (the e2e leg of bench.py scales
1.00 / 1.00 / 0.54 / 0.44 at 1 / 2 / 4 / 8 GPUs; is that the pipeline, or the host?).

For n in 1, 2, 4, 8 (up to the GPUs present) every GPU copies a 400 MB pinned buffer (one bench.py step's reads) to its
device memory `reps` times; reported: aggregate and per-GPU GB/s.  Variants:
  threads   one process, one host thread per GPU (what kmc_group_extract_host does)
  procs     one process per GPU (what torchrun + bench.py does)
  wc        pinned memory allocated write-combined (cudaHostAllocWriteCombined)
  d2h       the opposite direction, for reference
plus, once: nvidia-smi topo -m, the NUMA layout (lscpu) and the PCIe link of every GPU.
Plain cudart through ctypes: no library of this repository is involved (this measures the box, not the product).

  python tools/h2d_probe.py > profiles/r02_h2d_probe.json
"""
import ctypes as C
import json
import multiprocessing as mp
import subprocess
import sys
import threading
import time

BYTES = 400_000_000
H2D, D2H = 1, 2
WC = 4  # cudaHostAllocWriteCombined


def cudart():
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    import glob
    import os
    import torch  # noqa: F401  (its wheel bundles cudart)
    base = os.path.dirname(os.path.dirname(torch.__file__))
    for p in glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        return C.CDLL(p)
    raise OSError("libcudart not found")


def check(rt, st, what):
    if st != 0:
        rt.cudaGetErrorString.restype = C.c_char_p
        raise RuntimeError(f"{what}: {rt.cudaGetErrorString(st).decode()}")


class Lane:
    """One GPU: a pinned host buffer, a device buffer, a stream."""

    def __init__(self, rt, dev, flags):
        self.rt, self.dev = rt, dev
        check(rt, rt.cudaSetDevice(dev), "cudaSetDevice")
        self.h, self.d, self.s = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(rt, rt.cudaHostAlloc(C.byref(self.h), C.c_size_t(BYTES), C.c_uint(flags)), "cudaHostAlloc")
        C.memset(self.h, 1, BYTES)  # first touch by the thread / process that will issue the copies
        check(rt, rt.cudaMalloc(C.byref(self.d), C.c_size_t(BYTES)), "cudaMalloc")
        check(rt, rt.cudaStreamCreate(C.byref(self.s)), "cudaStreamCreate")

    def copy(self, reps, kind):
        rt = self.rt
        check(rt, rt.cudaSetDevice(self.dev), "cudaSetDevice")
        for _ in range(reps):
            if kind == H2D:
                st = rt.cudaMemcpyAsync(self.d, self.h, C.c_size_t(BYTES), C.c_int(1), self.s)
            else:
                st = rt.cudaMemcpyAsync(self.h, self.d, C.c_size_t(BYTES), C.c_int(2), self.s)
            check(rt, st, "cudaMemcpyAsync")
        check(rt, rt.cudaStreamSynchronize(self.s), "cudaStreamSynchronize")


def run_threads(n, reps, flags, kind):
    rt = cudart()
    lanes = [Lane(rt, i, flags) for i in range(n)]
    for ln in lanes:
        ln.copy(2, kind)
    bar = threading.Barrier(n + 1)
    done = threading.Barrier(n + 1)

    def work(ln):
        bar.wait()
        ln.copy(reps, kind)
        done.wait()
    th = [threading.Thread(target=work, args=(ln,)) for ln in lanes]
    for t in th:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    return n * reps * BYTES / dt / 1e9


def _proc(dev, reps, flags, kind, bar, done):
    rt = cudart()
    ln = Lane(rt, dev, flags)
    ln.copy(2, kind)
    bar.wait()
    ln.copy(reps, kind)
    done.wait()


def run_procs(n, reps, flags, kind):
    ctx = mp.get_context("spawn")
    bar, done = ctx.Barrier(n + 1), ctx.Barrier(n + 1)
    ps = [ctx.Process(target=_proc, args=(i, reps, flags, kind, bar, done)) for i in range(n)]
    for p in ps:
        p.start()
    bar.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for p in ps:
        p.join()
    return n * reps * BYTES / dt / 1e9


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout
    except Exception as e:  # noqa: BLE001
        return f"{cmd}: {e}"


def main():
    rt = cudart()
    n_dev = C.c_int(0)
    check(rt, rt.cudaGetDeviceCount(C.byref(n_dev)), "cudaGetDeviceCount")
    reps = 10
    out = {"bytes_per_copy": BYTES, "reps": reps, "gpus": n_dev.value, "rows": []}
    for n in (1, 2, 4, 8):
        if n > n_dev.value:
            break
        for name, fn, flags, kind in (("threads", run_procs if False else run_threads, 0, H2D), ("procs", run_procs, 0, H2D),
                                      ("procs_wc", run_procs, WC, H2D), ("procs_d2h", run_procs, 0, D2H)):
            # (threads in a child process, so that every variant starts from a fresh CUDA context)
            if name == "threads":
                ctx = mp.get_context("spawn")
                q = ctx.Queue()
                p = ctx.Process(target=_threads_child, args=(n, reps, flags, kind, q))
                p.start()
                gbs = q.get()
                p.join()
            else:
                gbs = fn(n, reps, flags, kind)
            out["rows"].append({"gpus": n, "variant": name, "aggregate_GBps": gbs, "per_gpu_GBps": gbs / n})
            print(f"# {n} GPU(s) {name}: {gbs:.1f} GB/s aggregate, {gbs / n:.1f} per GPU", file=sys.stderr, flush=True)
    out["topo"] = sh("nvidia-smi topo -m")
    out["numa"] = sh("lscpu | grep -i -E 'numa|socket|^CPU\\(s\\)|model name'")
    out["pcie"] = sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv")
    out["mem"] = sh("free -g | head -2")
    print(json.dumps(out))


def _threads_child(n, reps, flags, kind, q):
    q.put(run_threads(n, reps, flags, kind))


if __name__ == "__main__":
    main()

"""Is the C2 kernel power-limited in a back-to-back loop?  Times the same launch (a) back to back, (b) with idle gaps."""
import ctypes as C, os, sys, time, subprocess, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmers.jl_b200"))
import numpy as np, torch
import kmerscuda as kc
from kmerscuda import _abi
ctx = kc.Context(0)
n_reads = 10_000_000; n = n_reads * 120
g = torch.Generator(device="cuda").manual_seed(1)
words = torch.randint(-2**63, 2**63 - 1, (n_reads * 5,), dtype=torch.int64, device="cuda", generator=g)
a = torch.empty(n, dtype=torch.int64, device="cuda"); h = torch.empty(n, dtype=torch.int64, device="cuda")
desc = _abi.kmc_seqs(words.data_ptr(), words.numel(), n_reads, None, None, 150, 5, 2, 0)
res = _abi.kmc_result()
def step(mode, hash_):
    out = _abi.kmc_out(a.data_ptr(), h.data_ptr() if mode == 1 else None, h.data_ptr() if hash_ else None, None, None, n, 0)
    st = ctx.lib.kmc_extract(ctx.handle, C.byref(desc), 31, mode, (1 if hash_ else 0) | 4, C.byref(out), C.byref(res))
    assert st == 0
samples = []
stop = False
def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
    while not stop:
        l = p.stdout.readline()
        if l: samples.append((time.time(), l.strip()))
    p.kill()
th = threading.Thread(target=sampler); th.start()
time.sleep(1.0)
for name, mode, hash_ in (("canon+hash", 2, True), ("fwrv soa", 1, False)):
    for gap in (0.0, 0.02):
        for _ in range(3): step(mode, hash_)
        t0 = time.time(); ms = []
        for i in range(60 if gap else 400):
            ctx.timer_begin()
            step(mode, hash_)
            ms.append(ctx.timer_end())
            if gap: time.sleep(gap)
        t1 = time.time()
        clk = [s[1] for s in samples if t0 <= s[0] <= t1]
        sm = [float(c.split(",")[0]) for c in clk]; pw = [float(c.split(",")[1]) for c in clk]
        print(f"{name:12s} gap {gap*1e3:5.0f} ms: kernel median {np.median(ms):.3f} min {np.min(ms):.3f} ms; sm clock min/median {min(sm) if sm else 0:.0f}/{np.median(sm) if sm else 0:.0f} MHz; power max {max(pw) if pw else 0:.0f} W ({len(clk)} samples)", flush=True)
stop = True; th.join()

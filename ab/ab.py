"""A/B of two builds of the library on the same box: C2 kernel time (CUDA events), alternating.
usage: python ab/ab.py libA.so libB.so"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, ctypes as C
import numpy as np, torch
lib = C.CDLL(sys.argv[1])
class S(C.Structure):
    _fields_=[("words",C.c_void_p),("n_words",C.c_uint64),("n_seqs",C.c_uint64),("off",C.c_void_p),("len",C.c_void_p),("ulen",C.c_uint64),("ustride",C.c_uint64),("bits",C.c_uint32),("first",C.c_uint32)]
class O(C.Structure):
    _fields_=[("a",C.c_void_p),("b",C.c_void_p),("hash",C.c_void_p),("index",C.c_void_p),("so",C.c_void_p),("cap",C.c_uint64),("ib",C.c_int64)]
class R(C.Structure):
    _fields_=[("n",C.c_uint64),("es",C.c_uint64),("ep",C.c_uint64),("esym",C.c_uint32),("ms",C.c_float),("dig",C.c_uint64*4)]
lib.kmc_ctx_create.argtypes=[C.c_int32,C.POINTER(C.c_void_p)]
lib.kmc_extract.argtypes=[C.c_void_p,C.POINTER(S),C.c_int32,C.c_int32,C.c_uint32,C.POINTER(O),C.POINTER(R)]
lib.kmc_timer_begin.argtypes=[C.c_void_p]; lib.kmc_timer_end.argtypes=[C.c_void_p,C.POINTER(C.c_float)]
lib.kmc_sync.argtypes=[C.c_void_p]
ctx=C.c_void_p(); assert lib.kmc_ctx_create(0,C.byref(ctx))==0
n_reads=10_000_000; n=n_reads*120
g = torch.Generator(device="cuda").manual_seed(1)
words = torch.randint(-2**63, 2**63-1, (n_reads*5,), dtype=torch.int64, device="cuda", generator=g)
a = torch.empty(n, dtype=torch.int64, device="cuda"); h = torch.empty(n, dtype=torch.int64, device="cuda")
mode=int(sys.argv[2]); flags=int(sys.argv[3])
desc = S(words.data_ptr(), words.numel(), n_reads, None, None, 150, 5, 2, 0)
out = O(a.data_ptr(), None, h.data_ptr() if flags&1 else None, None, None, n, 0)
res = R()
torch.cuda.synchronize()
def step():
    assert lib.kmc_extract(ctx, C.byref(desc), 31, mode, flags|4, C.byref(out), C.byref(res)) == 0
for _ in range(5): step()
lib.kmc_sync(ctx)
ms=[]
for _ in range(30):
    lib.kmc_timer_begin(ctx); step(); t=C.c_float(); lib.kmc_timer_end(ctx,C.byref(t)); ms.append(t.value)
print("median %.4f min %.4f ms" % (float(np.median(ms)), float(np.min(ms))))
'''
libs = sys.argv[1:]
cases = [(2, 1, "canon+hash"), (2, 0, "canon"), (0, 0, "fw")]
for mode, flags, nm in cases:
    for rep in range(2):
        for lib in libs:
            r = subprocess.run([sys.executable, "-c", code, lib, str(mode), str(flags)], capture_output=True, text=True)
            print(nm, rep, os.path.basename(lib), r.stdout.strip(), r.stderr.strip()[-200:], flush=True)

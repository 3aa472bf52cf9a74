#!/usr/bin/env python
"""bench.py -- canonical 31-mers + fx_hash per second (BASELINE.json metric, config C2).

One "step" is one pass of the hot path over one batch of synthetic reads:
CanonicalDNAMers{31} + fx_hash over 10 M x 150 bp 2-bit reads per GPU (1.2 G k-mers, 400 MB in,
19.2 GB out), weak scaling (every rank owns its own 10 M reads; no data-path collective).

  value     device-resident throughput (inputs in HBM, outputs stay in HBM), CUDA events, max over ranks
  e2e       the same metric through kmc_extract_host with pinned HOST sequence buffers: H2D of the
            step's reads inside, streams left in HBM (KMC_OUT_DEVICE), 32-byte result fingerprint read
            back; e2e_host_streams = the same with both full streams copied to pinned host memory
  roofline  algorithmic bytes (16.3125 B / k-mer) / measured kernel time vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference   the CPU restatement of the reference's per-symbol recurrence
            (oracle/, OpenMP over reads, all host cores) -- Julia is not installed, so the reference
            itself cannot run; see DESIGN.md.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "kmers.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

K = 31
READ_LEN = 150
STRIDE = 5  # words per read (word-aligned CSR)
WPR = READ_LEN - K + 1
BYTES_PER_KMER = 0.25 * READ_LEN / WPR + 8 + 8  # 16.3125
SEED = 439824  # test/benchmark.jl:19
METRIC = "canonical 31-mers+hash/sec"
UNIT = "kmers/s"


def splitmix64(x):
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_reads(n_reads, rank=0, out=None):
    """word[j] = splitmix64(seed + j); trailing bits of each read's last word zeroed."""
    n = n_reads * STRIDE
    base = np.uint64(SEED) + np.uint64(rank) * np.uint64(1 << 40)
    words = out if out is not None else np.empty(n, dtype=np.uint64)
    step = 1 << 24
    for s in range(0, n, step):
        e = min(n, s + step)
        words[s:e] = splitmix64(np.arange(s, e, dtype=np.uint64) + base)
    tail = READ_LEN - 32 * (STRIDE - 1)
    words.reshape(n_reads, STRIDE)[:, STRIDE - 1] &= np.uint64((1 << (2 * tail)) - 1)
    return words


def workload_name(n_reads):
    return (f"C2: CanonicalDNAMers{{{K}}} + fx_hash over {n_reads:,} x {READ_LEN} bp 2-bit reads per GPU "
            f"({n_reads * WPR:,} k-mers, SoA canon u64 + hash u64)")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    # (timestamp last: it contains no comma-separated sub-fields, but keeps the indices below stable)
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        """Median SM clock, maximum power and the throttle reasons over the samples taken between the wall-clock
        times t0 and t1 (the timed region), all samples if none fall inside."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    row = {"sm": float(f[1]), "mx": float(f[2]), "w": float(f[3]), "t": None,
                           "reasons": {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                                "sw_power_cap"), f[5:9]) if v.lower().startswith("active")}}
                except ValueError:
                    continue
                if len(f) > 9:
                    try:
                        row["t"] = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except ValueError:
                        pass
                rows.append(row)
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if t0 is not None and r["t"] is not None and t0 <= r["t"] <= t1]
        use = inside or rows
        if use:
            reasons = set().union(*[r["reasons"] for r in use])
            out.update(sm_mhz=float(np.median([r["sm"] for r in use])), sm_max_mhz=float(max(r["mx"] for r in use)),
                       reasons=sorted(reasons), samples=len(use), power_w_max=float(max(r["w"] for r in use)),
                       window="timed region" if inside else "whole run (no sample carried a timestamp inside the timed region)")
        return out


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's literal per-symbol recurrence, OpenMP over reads (test infrastructure
# used here only as the measured CPU baseline, never on the product path)
# --------------------------------------------------------------------------------------------
def cpu_throughput(sample_reads, min_seconds, threads=0):
    from oracle import oracle as ko
    cores = ko.max_threads() if threads <= 0 else threads
    words = synth_reads(sample_reads)
    n = sample_reads * WPR
    a = np.empty((n, 1), dtype=np.uint64)
    h = np.empty(n, dtype=np.uint64)
    ko.batch_iterate(words, sample_reads, K, ko.CANON, uniform_len=READ_LEN, uniform_stride=STRIDE,
                     want_hash=True, threads=cores, out=(a, None, h))  # warm-up (page faults, threads)
    passes, t0 = 0, time.perf_counter()
    while True:
        ko.batch_iterate(words, sample_reads, K, ko.CANON, uniform_len=READ_LEN, uniform_stride=STRIDE,
                         want_hash=True, threads=cores, out=(a, None, h))
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return n * passes / dt, cores, passes, dt


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    sample = min(args.reads, args.cpu_sample_reads)
    from oracle import oracle as ko
    cores = ko.max_threads()
    words = synth_reads(sample)
    n = sample * WPR
    a = np.empty((n, 1), dtype=np.uint64)
    h = np.empty(n, dtype=np.uint64)

    def step():
        ko.batch_iterate(words, sample, K, ko.CANON, uniform_len=READ_LEN, uniform_stride=STRIDE,
                         want_hash=True, threads=cores, out=(a, None, h))
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample_desc = (f"{sample:,} of the {args.reads:,} reads per step ({n:,} k-mers/step), "
                   f"literal per-symbol recurrence, OpenMP parallel-for over reads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.reads), "k": K, "read_len": READ_LEN,
                   "note": "Julia is not installed: this arm times the C restatement of the reference algorithm (oracle/)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import ctypes as C

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print a banner on stdout, which must carry exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import kmerscuda as kc
    from kmerscuda import _abi

    ctx = kc.Context(local_rank)
    lib = ctx.lib
    n_reads = args.reads
    n_kmers = n_reads * WPR

    # ---- inputs: generated on the host once, resident in HBM before the timed region ----------
    pinned_in = ctx.pinned(n_reads * STRIDE * 8, np.uint64)
    synth_reads(n_reads, rank, out=pinned_in)
    d_words = torch.empty(n_reads * STRIDE, dtype=torch.int64, device="cuda")
    d_canon = torch.empty(n_kmers, dtype=torch.int64, device="cuda")
    d_hash = torch.empty(n_kmers, dtype=torch.int64, device="cuda")
    # a dedicated (non-default) torch stream: the library launches on it and torch.cuda.Event
    # records on it, so the events bracket exactly the kernels they are meant to time
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ctx._check(lib.kmc_upload(ctx.handle, d_words.data_ptr(), pinned_in.ctypes.data, pinned_in.nbytes))
    ctx.sync()
    desc = _abi.kmc_seqs(d_words.data_ptr(), d_words.numel(), n_reads, None, None, READ_LEN, STRIDE, 2, 0)
    out = _abi.kmc_out(d_canon.data_ptr(), None, d_hash.data_ptr(), None, None, n_kmers, 0)
    res = _abi.kmc_result()
    flags = _abi.KMC_HASH_FX | _abi.KMC_NO_SYNC

    def step():
        st = lib.kmc_extract(ctx.handle, C.byref(desc), K, _abi.KMC_CANON, flags, C.byref(out), C.byref(res))
        if st != 0:
            raise RuntimeError(f"kmc_extract failed: {lib.kmc_last_error(ctx.handle).decode()}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    t_begin.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
    t_end.record(stream)
    barrier()
    wall1 = time.time()
    total_ms = t_begin.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    assert res.n_written == n_kmers

    # The same launch with idle gaps (outside the timed region, reported beside the roofline): back to back the
    # kernel runs into the board's power cap and the SM clock drops (tools/diag_power.py: 1000 W, ~1650 of 1965 MHz);
    # with 25 ms of idle between launches it runs at full clock.  The gap between the two is power, not the kernel.
    gap_ms = []
    if rank == 0:
        for _ in range(12):
            time.sleep(0.025)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step()
            b.record(stream)
            torch.cuda.synchronize()
            gap_ms.append(a.elapsed_time(b))
    barrier()

    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = n_kmers * world * args.steps / (max_ms / 1e3)

    # ---- sanity: sampled reads against the oracle (outside the timed region) -------------------
    if rank == 0 and not args.no_check:
        from oracle import oracle as ko
        rng = np.random.default_rng(7)
        for r in rng.choice(n_reads, size=64, replace=False):
            w = pinned_in[r * STRIDE:(r + 1) * STRIDE]
            a, _, h = ko.iterate(w, READ_LEN, K, ko.CANON, want_hash=True)
            got_a = d_canon[r * WPR:(r + 1) * WPR].cpu().numpy().view(np.uint64)
            got_h = d_hash[r * WPR:(r + 1) * WPR].cpu().numpy().view(np.uint64)
            if not (np.array_equal(got_a, a[:, 0]) and np.array_equal(got_h, h)):
                raise SystemExit("bench.py: GPU output differs from the oracle; refusing to report a number")

    # ---- end to end through the C ABI with HOST sequence buffers ---------------------------------
    # e2e (headline): every step uploads that step's reads from pinned host memory (chunked H2D
    # overlapped with the kernels inside kmc_extract_host), leaves the canonical + hash streams in
    # HBM for a device consumer (KMC_OUT_DEVICE -- the deployment north_star describes: each GPU
    # emits its streams, only tables / sketches / fingerprints leave the GPU), and reads back the
    # step's result fingerprint (KMC_DIGEST: xor + wrapping sum of both streams, 32 bytes).
    # e2e_host_streams: the same call materialising both full streams in pinned HOST memory
    # (19.2 GB over PCIe per step) -- reported beside it; it is PCIe-bound by construction.
    e2e = e2e_host = None
    if not args.no_e2e:
        hdesc = _abi.kmc_seqs(pinned_in.ctypes.data, n_reads * STRIDE, n_reads, None, None, READ_LEN, STRIDE, 2, 0)
        dout = _abi.kmc_out(d_canon.data_ptr(), None, d_hash.data_ptr(), None, None, n_kmers, 0)
        hres = _abi.kmc_result()
        want_dig = (ctx.digest(d_canon.data_ptr(), n_kmers), ctx.digest(d_hash.data_ptr(), n_kmers))  # from the resident run
        d_canon.zero_()
        d_hash.zero_()
        torch.cuda.synchronize()

        def e2e_step():
            st = lib.kmc_extract_host(ctx.handle, C.byref(hdesc), K, _abi.KMC_CANON,
                                      _abi.KMC_HASH_FX | _abi.KMC_OUT_DEVICE | _abi.KMC_DIGEST, C.byref(dout), C.byref(hres))
            if st != 0:
                raise RuntimeError(f"e2e step failed: {lib.kmc_last_error(ctx.handle).decode()}")
        for _ in range(2):
            e2e_step()  # warm-up (allocates the pipeline slots)
        dig = hres.digest
        got_dig = ((int(dig[0]), int(dig[1])), (int(dig[2]), int(dig[3])))
        if got_dig != want_dig:
            raise SystemExit("bench.py: e2e result fingerprint differs from the device-resident run")
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_kmers * world * args.e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": n_reads * STRIDE * 8, "d2h_bytes_per_step": 32, "steps": args.e2e_steps,
               "ms_per_step": float(dt.item()) / args.e2e_steps * 1e3,
               "timer": "host wall clock around the calls (each synchronises), max over ranks",
               "path": "one kmc_extract_host call: pinned host words -> chunked H2D overlapped with the kernels -> canonical "
                       "+ hash streams resident in HBM (KMC_OUT_DEVICE), fingerprinted chunk by chunk (KMC_DIGEST) -> "
                       "32-byte result to host",
               "result_check": "digest equals the device-resident run's; sampled reads equal the oracle"}

        # -- the same call with both streams written to pinned host memory
        del d_canon, d_hash
        torch.cuda.empty_cache()
        # 16 bytes of pinned host memory per k-mer: the full 19.2 GB on one GPU, a 1/world share per rank
        # otherwise (all ranks share one host's memory and one root complex; the number is PCIe-bound anyway)
        e2e_reads = max(1, n_reads // world)
        try:
            h_canon = ctx.pinned(e2e_reads * WPR * 8, np.uint64)
            h_hash = ctx.pinned(e2e_reads * WPR * 8, np.uint64)
        except kc.KmersCUDAError:
            e2e_reads = max(1, e2e_reads // 10)
            h_canon = ctx.pinned(e2e_reads * WPR * 8, np.uint64)
            h_hash = ctx.pinned(e2e_reads * WPR * 8, np.uint64)
        hdesc2 = _abi.kmc_seqs(pinned_in.ctypes.data, e2e_reads * STRIDE, e2e_reads, None, None, READ_LEN, STRIDE, 2, 0)
        hout = _abi.kmc_out(h_canon.ctypes.data, None, h_hash.ctypes.data, None, None, e2e_reads * WPR, 0)

        def host_step():
            st = lib.kmc_extract_host(ctx.handle, C.byref(hdesc2), K, _abi.KMC_CANON, _abi.KMC_HASH_FX, C.byref(hout), C.byref(hres))
            if st != 0:
                raise RuntimeError(f"kmc_extract_host failed: {lib.kmc_last_error(ctx.handle).decode()}")
        host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.host_steps):
            host_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rank == 0 and not args.no_check:
            from oracle import oracle as ko
            for r in (0, e2e_reads // 2, e2e_reads - 1):
                a, _, h = ko.iterate(pinned_in[r * STRIDE:(r + 1) * STRIDE], READ_LEN, K, ko.CANON, want_hash=True)
                if not (np.array_equal(h_canon[r * WPR:(r + 1) * WPR], a[:, 0]) and np.array_equal(h_hash[r * WPR:(r + 1) * WPR], h)):
                    raise SystemExit("bench.py: host-path output differs from the oracle")
        e2e_host = {"value": e2e_reads * WPR * world * args.host_steps / float(dt.item()), "unit": UNIT,
                    "h2d_bytes_per_step": e2e_reads * STRIDE * 8, "d2h_bytes_per_step": e2e_reads * WPR * 16,
                    "steps": args.host_steps, "reads_per_gpu": e2e_reads,
                    "path": "kmc_extract_host: pinned host words -> 3-slot H2D/kernel/D2H pipeline -> BOTH full streams in "
                            "pinned host memory (PCIe-bound: 16 B per k-mer over the link)"}

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = float(np.mean(kernel_ms))
        achieved = BYTES_PER_KMER * n_kmers / (k_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("extract_kernel_c2_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(n_reads), "k": K, "read_len": READ_LEN, "reads_per_gpu": n_reads,
                       "parallelism": f"reads sharded over {world} GPU(s), no data-path collective",
                       "l2": "no explicit flush: each step streams 0.4 GB in + 19.2 GB out, far larger than the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "extract_kernel<N=1,NX=3,CANON,HASH,uniform,G=8>",
                         "kernel_ms": k_ms, "bytes_per_kmer": BYTES_PER_KMER, "peak_source": peak_src,
                         "idle_gaps": {"kernel_ms": float(np.median(gap_ms)),
                                       "achieved": BYTES_PER_KMER * n_kmers / (float(np.median(gap_ms)) / 1e3) / 1e9,
                                       "how": "the same launch, 12 times with 25 ms of idle before each (full SM clock, no "
                                              "power cap); not the timed region"} if gap_ms else None,
                         "write_ceiling_note": "the peak is a copy (half reads); this kernel is 98 % writes, whose ceiling "
                                               "measures 7480 GB/s (profiles/r01_bw_probe_v2.json, tools/bw_probe.py)"},
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
            line["e2e_host_streams"] = e2e_host
        if world == 1 and not args.no_cpu:
            v, cores, passes, dt = cpu_throughput(min(n_reads, args.cpu_sample_reads), args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": (f"{min(n_reads, args.cpu_sample_reads):,} of the {n_reads:,} reads x {passes} passes in {dt:.1f} s; "
                           "C restatement of the reference's per-symbol recurrence, OpenMP over reads")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--host-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and "RANK" not in os.environ:
        # launched directly: re-exec under torchrun, one rank per GPU
        port = 29500 + os.getpid() % 2000
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())

"""The multi-GPU surface of the C ABI (comm.cu): NCCL communicators attached to contexts (one process per GPU),
the single-process device group, the bucket-count merge (the path's only collective) and the exact k-mer table
exchange.  The one-rank / one-device forms run on any GPU box (a real NCCL communicator of one rank); the two-GPU
tests are skipped on boxes with fewer than two GPUs (a 2-GPU run is kept under profiles/)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "kmers.jl_b200"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import kmertools as kt
    from oracle import oracle as ko
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import kmerscuda as kc
    from kmerscuda import sharding
    ctx = kc.Context(rank)
    ok = True
    k, bits = 31, 20
    rng = np.random.default_rng(77)
    # (1) read set: canonical + hash per shard, bucket table merged with NCCL
    n_reads, length, stride = 50_000, 150, 5
    words = rng.integers(0, 2**64, size=n_reads * stride, dtype=np.uint64)
    rs = kc.ReadSet(2, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    sub = sharding.read_shard(rs, world, rank)
    e = kc.extract(kc.KMC_CANON, sub, k, hash=True, ctx=ctx)
    a, _, h, _ = ko.batch_iterate(words, n_reads, k, ko.CANON, uniform_len=length, uniform_stride=stride, want_hash=True)
    plan = sharding.plan_read_shards(rs, world)
    lo = plan[rank].seq0 * (length - k + 1)
    ok &= bool(np.array_equal(e.kmers, a[lo:lo + e.n]) and np.array_equal(e.hash, h[lo:lo + e.n]))
    table = torch.zeros(1 << bits, dtype=torch.int32, device="cuda")
    drs = kc.DeviceReadSet(ctx, sub)
    import ctypes as C
    from kmerscuda import _abi
    res = _abi.kmc_result()
    ctx._check(ctx.lib.kmc_bucket_count(ctx.handle, C.byref(drs.desc), k, bits, table.data_ptr(), C.byref(res)))
    sharding.allreduce_table(table)
    want = np.bincount((h >> np.uint64(64 - bits)).astype(np.int64), minlength=1 << bits).astype(np.int32)
    ok &= bool(np.array_equal(table.cpu().numpy(), want))
    # (1b) a table beyond L2, counted slice after slice with the merge of finished ranges overlapping the count
    bits2 = 26
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx.set_stream(stream.cuda_stream)
        table2 = torch.zeros(1 << bits2, dtype=torch.int32, device="cuda")
        n_counted = sharding.count_and_merge_table(ctx, drs.desc, k, bits2, table2, n_parts=8)
        stream.synchronize()
        ctx.set_stream(None)
    ok &= n_counted == e.n
    want2 = np.bincount((h >> np.uint64(64 - bits2)).astype(np.int64), minlength=1 << bits2).astype(np.int32)
    ok &= bool(np.array_equal(table2.cpu().numpy(), want2))
    # (2) one long 4-bit sequence: unambiguous k-mers with global indices
    n = 400_001
    codes = np.where(rng.random(n) < 0.01, np.uint64(15), np.uint64(1) << rng.integers(0, 4, size=n).astype(np.uint64))
    w4 = kt.pack_codes(codes, 4)
    km, pos = ko.unambiguous(w4, n, k, src_bits=4)
    srs, base = sharding.sequence_shard(4, w4, n, k, world, rank)
    e = kc.extract(kc.KMC_UNAMBIG, srs, k, ctx=ctx, index_base=base)
    counts = sharding.gather_counts(e.n)
    off = sum(counts[:rank])
    ok &= sum(counts) == km.shape[0]
    ok &= bool(np.array_equal(e.kmers, km[off:off + e.n]) and np.array_equal(e.index, pos[off:off + e.n]))
    # (3) the collective inside the library: a communicator attached to the context (id broadcast by torch), the
    #     count + overlapped merge in one C call -- an L2-sized table and one beyond L2
    ctx.comm_init_torch()
    ok &= ctx.comm_info() == (rank, world)
    for b, w in ((bits, want), (bits2, want2)):
        t = ctx.alloc(4 << b)
        ctx._check(ctx.lib.kmc_memset(ctx.handle, t.ptr, 0, 4 << b))
        n_counted, _ = ctx.bucket_count_merge(drs.desc, k, b, t.ptr)
        ok &= n_counted == plan[rank].n_seqs * (length - k + 1)
        ok &= bool(np.array_equal(t.download(np.uint32, 1 << b).astype(np.int32), w))
        t.free()
    # (4) exact k-mer table across GPUs: every rank counts its shard, entries travel to their owners
    tab = kc.KmerTable(18, ctx=ctx)
    gcodes = rng.integers(0, 4, size=20_000).astype(np.uint64)  # reads drawn from a small genome: repeated k-mers
    starts = rng.integers(0, 20_000 - length, size=4000)
    rcodes = np.zeros((4000, stride * 32), dtype=np.uint64)
    rcodes[:, :length] = gcodes[starts[:, None] + np.arange(length)[None, :]]
    rwords = kt.pack_codes(rcodes.reshape(-1), 2)
    rs2 = kc.ReadSet(2, rwords.reshape(-1), 4000, uniform_len=length, uniform_stride_words=stride)
    tab.count(sharding.read_shard(rs2, world, rank), k, canonical=True)
    owned = tab.exchange(18)
    ka, _, _, _ = ko.batch_iterate(rs2.words, 4000, k, ko.CANON, uniform_len=length, uniform_stride=stride)
    uk, uc = np.unique(ka[:, 0], return_counts=True)
    mine = np.array([ctx.lib.kmc_kmer_owner(int(x), world) == rank for x in uk])
    gk, gv = owned.items()
    ok &= bool(np.array_equal(gk, uk[mine]) and np.array_equal(gv.astype(np.int64), uc[mine]))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_and_table_merge():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        assert ret.get(0) is True and ret.get(1) is True


def test_one_rank_communicator_and_one_device_group():
    """Runs on any GPU box: a real NCCL communicator of ONE rank attached to a context -- kmc_bucket_count_merge,
    kmc_allreduce, kmc_kmer_table_exchange go through NCCL and must equal the single-GPU results -- and a kmc_group of
    one device."""
    import ctypes as C
    import kmerscuda as kc
    from kmerscuda import _abi
    from oracle import oracle as ko
    lib = _abi.load()
    v = C.c_int32()
    assert lib.kmc_nccl_version(C.byref(v)) == 0 and v.value >= 21000
    ctx = kc.Context(0)
    ctx.comm_init(1, 0, kc.Context.comm_unique_id())
    assert ctx.comm_info() == (0, 1)
    rng = np.random.default_rng(5)
    k, n_reads, length, stride = 31, 20_000, 150, 5
    words = rng.integers(0, 2**64, size=n_reads * stride, dtype=np.uint64)
    rs = kc.ReadSet(2, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    drs = kc.DeviceReadSet(ctx, rs)
    _, _, h, _ = ko.batch_iterate(words, n_reads, k, ko.CANON, uniform_len=length, uniform_stride=stride, want_hash=True)
    for bits in (6, 20, 26):
        want = np.bincount((h >> np.uint64(64 - bits)).astype(np.int64), minlength=1 << bits).astype(np.uint32)
        t = ctx.alloc(4 << bits)
        ctx._check(lib.kmc_memset(ctx.handle, t.ptr, 0, 4 << bits))
        n, ms = ctx.bucket_count_merge(drs.desc, k, bits, t.ptr)
        assert n == h.size and ms > 0
        assert np.array_equal(t.download(np.uint32, 1 << bits), want)
        ctx.allreduce(t.ptr, 1 << bits)  # one rank: the identity, through ncclAllReduce
        ctx.sync()
        assert np.array_equal(t.download(np.uint32, 1 << bits), want)
        t.free()
    tab = kc.KmerTable(20, ctx=ctx)
    small = kc.ReadSet(2, words[: 2000 * stride], 2000, uniform_len=length, uniform_stride_words=stride)
    tab.count(small, k)
    owned = tab.exchange(20)
    k0, v0 = tab.items()
    k1, v1 = owned.items()
    assert np.array_equal(k0, k1) and np.array_equal(v0, v1)
    ctx.comm_destroy()
    ctx.close()

    g = kc.Group([0])
    assert len(g) == 1
    merged, n_per, ms, tables = g.bucket_count([kc.DeviceReadSet(g.ctx[0], rs)], k, 20)
    assert n_per == [h.size]
    assert np.array_equal(merged, np.bincount((h >> np.uint64(44)).astype(np.int64), minlength=1 << 20).astype(np.uint32))
    e = g.extract_host(kc.KMC_CANON, [rs], k, hash=True)[0]
    a, _, hh, _ = ko.batch_iterate(words, n_reads, k, ko.CANON, uniform_len=length, uniform_stride=stride, want_hash=True)
    assert np.array_equal(e.kmers, a) and np.array_equal(e.hash, hh)
    g.close()


def test_group_two_gpus_single_process():
    """One process, two GPUs (what a Julia session is): kmc_group_create (ncclCommInitAll), C5's count + merge, the
    sharded `collect`, and the k-mer table exchange, against the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import kmertools as kt
    import kmerscuda as kc
    from kmerscuda import sharding
    from oracle import oracle as ko
    g = kc.Group([0, 1])
    rng = np.random.default_rng(99)
    k, n_reads, length, stride = 31, 60_001, 150, 5
    words = rng.integers(0, 2**64, size=n_reads * stride, dtype=np.uint64)
    rs = kc.ReadSet(2, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    shards = [sharding.read_shard(rs, 2, r) for r in range(2)]
    a, _, h, _ = ko.batch_iterate(words, n_reads, k, ko.CANON, uniform_len=length, uniform_stride=stride, want_hash=True)
    for bits in (20, 26):
        merged, n_per, ms, tables = g.bucket_count([kc.DeviceReadSet(c, s) for c, s in zip(g.ctx, shards)], k, bits)
        want = np.bincount((h >> np.uint64(64 - bits)).astype(np.int64), minlength=1 << bits).astype(np.uint32)
        assert sum(n_per) == h.size
        assert np.array_equal(merged, want)
        assert np.array_equal(tables[1].download(np.uint32, 1 << bits), want)  # every device holds the merged table
    parts = g.extract_host(kc.KMC_CANON, shards, k, hash=True)
    assert np.array_equal(np.concatenate([p.kmers for p in parts]), a)
    assert np.array_equal(np.concatenate([p.hash for p in parts]), h)
    # one long 4-bit sequence, UnambiguousKmers with global indices
    n = 300_001
    codes = np.where(rng.random(n) < 0.01, np.uint64(15), np.uint64(1) << rng.integers(0, 4, size=n).astype(np.uint64))
    w4 = kt.pack_codes(codes, 4)
    km, pos = ko.unambiguous(w4, n, k, src_bits=4)
    sh = [sharding.sequence_shard(4, w4, n, k, 2, r) for r in range(2)]
    parts = g.extract_host(kc.KMC_UNAMBIG, [x[0] for x in sh], k, index_bases=[x[1] for x in sh])
    assert np.array_equal(np.concatenate([p.kmers for p in parts]), km)
    assert np.array_equal(np.concatenate([p.index for p in parts]), pos)
    # exact k-mer table: per-device counts, exchange by owner
    small = kc.ReadSet(2, np.tile(words[: 3000 * stride], 2), 6000, uniform_len=length, uniform_stride_words=stride)
    tabs = []
    for r, c in enumerate(g.ctx):
        t = kc.KmerTable(20, ctx=c)
        t.count(sharding.read_shard(small, 2, r), k)
        tabs.append(t)
    owned = g.kmer_table_exchange(tabs, 20)
    ka, _, _, _ = ko.batch_iterate(small.words, 6000, k, ko.CANON, uniform_len=length, uniform_stride=stride)
    uk, uc = np.unique(ka[:, 0], return_counts=True)
    own = np.array([g.lib.kmc_kmer_owner(int(x), 2) for x in uk])
    for r in range(2):
        gk, gv = owned[r].items()
        assert np.array_equal(gk, uk[own == r]) and np.array_equal(gv.astype(np.int64), uc[own == r])
    g.close()


def test_c_example_counts_and_merges_without_python(tmp_path):
    """examples/c5_group_count.c: the C5 job (count on every GPU + NCCL merge inside the library) from plain C, on all
    the GPUs of the box, at a small size; it checks the merged total and the equality of the tables itself."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "kmers.jl_b200")
    exe = str(tmp_path / "c5_group_count")
    subprocess.run(["gcc", "-std=c99", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c5_group_count.c"),
                    "-L", lib_dir, "-lkmerscuda", "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    for bits in ("20", "27"):
        r = subprocess.run([exe, "200000", bits], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "ok; identical on all" in r.stdout

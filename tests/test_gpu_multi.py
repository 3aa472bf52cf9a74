"""Two GPUs, one process each (NCCL): shards extracted through the C ABI concatenate to the
unsharded result, emitted indices are global, and the bucket-count tables merge with the path's
only collective.  Skipped on boxes with fewer than two GPUs."""
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

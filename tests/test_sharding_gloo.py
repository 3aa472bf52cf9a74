"""Multi-GPU host logic on CPU: the shard planners and the one collective of the path (the sum of
the per-rank bucket-count tables), exercised with world_size = 2 over gloo.  The per-shard
"compute" here is the oracle (test infrastructure) -- what is under test is that shards tile the
input exactly, that rank-order concatenation reproduces the unsharded (= reference) order, and that
the all-reduce merges the tables."""
import os
import socket

import numpy as np
import pytest

import kmertools as kt
from oracle import oracle as ko


@pytest.fixture(scope="module")
def sh():
    from kmerscuda import sharding
    return sharding


@pytest.fixture(scope="module")
def kc():
    import kmerscuda
    return kmerscuda


def ragged_set(kc, rng, n, bits=2, amb=0.0):
    lens = rng.integers(0, 300, size=n)
    spw = 64 // bits
    nw = (lens + spw - 1) // spw
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(nw)
    if bits == 2:
        words = rng.integers(0, 2**64, size=int(off[-1]) + 1, dtype=np.uint64)
    else:
        codes = np.uint64(1) << rng.integers(0, 4, size=(int(off[-1]) + 1) * 16).astype(np.uint64)
        codes = np.where(rng.random(codes.size) < amb, np.uint64(15), codes)
        words = kt.pack_codes(codes, 4)
    return kc.ReadSet(bits, words, n, seq_word_offset=off[:-1].copy(), seq_len=lens.astype(np.uint64))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_read_shards_tile_and_concatenate(kc, sh, world):
    rng = np.random.default_rng(world)
    k = 31
    for rs in (ragged_set(kc, rng, 500),
               kc.ReadSet(2, rng.integers(0, 2**64, size=300 * 5, dtype=np.uint64), 300, uniform_len=150, uniform_stride_words=5)):
        plan = sh.plan_read_shards(rs, world)
        assert [p.rank for p in plan] == list(range(world))
        assert plan[0].seq0 == 0 and sum(p.n_seqs for p in plan) == rs.n_seqs
        for a, b in zip(plan, plan[1:]):
            assert a.seq0 + a.n_seqs == b.seq0
        if rs.seq_len is not None and world > 1:  # balanced by symbols within one read's length
            tot = [int(rs.seq_len[p.seq0:p.seq0 + p.n_seqs].sum()) for p in plan]
            assert max(tot) - min(tot) <= 2 * int(rs.seq_len.max())
        whole = ko.batch_iterate(rs.words, rs.n_seqs, k, ko.CANON, word_off=rs.seq_word_offset, seq_len=rs.seq_len,
                                 uniform_len=rs.uniform_len, uniform_stride=rs.uniform_stride_words, want_hash=True)
        parts_a, parts_h = [], []
        for g in range(world):
            sub = sh.read_shard(rs, world, g)
            if sub.n_seqs == 0:
                continue
            a, _, h, _ = ko.batch_iterate(sub.words, sub.n_seqs, k, ko.CANON, word_off=sub.seq_word_offset, seq_len=sub.seq_len,
                                          uniform_len=sub.uniform_len, uniform_stride=sub.uniform_stride_words, want_hash=True)
            parts_a.append(a)
            parts_h.append(h)
        assert np.array_equal(np.concatenate(parts_a), whole[0]) and np.array_equal(np.concatenate(parts_h), whole[2])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("k", [31, 63])
def test_sequence_shards_with_halo(kc, sh, world, k):
    rng = np.random.default_rng(10 * world + k)
    n = 100_003
    # 2-bit: canonical k-mers + hashes
    w = rng.integers(0, 2**64, size=(n + 31) // 32, dtype=np.uint64)
    a, _, h = ko.iterate(w, n, k, ko.CANON, want_hash=True)
    pa, ph = [], []
    plan = sh.plan_sequence_shards(n, k, 2, world)
    assert sum(p.n_windows for p in plan) == n - k + 1
    for g in range(world):
        rs, base = sh.sequence_shard(2, w, n, k, world, g)
        x, _, y = ko.iterate(rs.words, rs.uniform_len, k, ko.CANON, first=rs.first_symbol_offset, want_hash=True)
        assert x.shape[0] == plan[g].n_windows and base == plan[g].window0
        pa.append(x)
        ph.append(y)
    assert np.array_equal(np.concatenate(pa), a) and np.array_equal(np.concatenate(ph), h)
    # 4-bit with N: unambiguous k-mers, indices made global by index_base
    codes = np.where(rng.random(n) < 0.01, np.uint64(15), np.uint64(1) << rng.integers(0, 4, size=n).astype(np.uint64))
    w4 = kt.pack_codes(codes, 4)
    km, pos = ko.unambiguous(w4, n, k, src_bits=4)
    pk, pp = [], []
    for g in range(world):
        rs, base = sh.sequence_shard(4, w4, n, k, world, g)
        x, y = ko.unambiguous(rs.words, rs.uniform_len, k, src_bits=4, first=rs.first_symbol_offset)
        pk.append(x)
        pp.append(y + base)
    assert np.array_equal(np.concatenate(pk), km) and np.array_equal(np.concatenate(pp), pos)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_ascii_read_shards_tile_and_concatenate(kc, sh, world):
    """ASCII sets (bits = 8) count BYTES in every word quantity: one symbol per unit.  The planner used 64 // bits
    = 8 symbols per unit and truncated the last read of every shard.  Ragged strings and the advisor's
    own reproduction (8 x 'ACGT' * 25)."""
    rng = np.random.default_rng(100 + world)
    k = 31
    sets = [kc.ReadSet.from_strings(["ACGT" * 25] * 8),
            kc.ReadSet.from_strings(["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(0, 300, size=200)])]
    for rs in sets:
        plan = sh.plan_read_shards(rs, world)
        assert plan[0].seq0 == 0 and sum(p.n_seqs for p in plan) == rs.n_seqs
        assert sum(p.n_words for p in plan) == int(rs.seq_len.sum())  # every byte belongs to exactly one shard
        whole_a, whole_h = [], []
        for r in range(rs.n_seqs):
            o, n = int(rs.seq_word_offset[r]), int(rs.seq_len[r])
            a, _, h = ko.ascii_iterate(bytes(rs.words[o:o + n]), k, ko.CANON, want_hash=True)
            whole_a.append(a)
            whole_h.append(h)
        parts_a, parts_h = [], []
        for g in range(world):
            sub = sh.read_shard(rs, world, g)
            for r in range(sub.n_seqs):
                o, n = int(sub.seq_word_offset[r]), int(sub.seq_len[r])
                assert o + n <= sub.words.size
                a, _, h = ko.ascii_iterate(bytes(sub.words[o:o + n]), k, ko.CANON, want_hash=True)
                parts_a.append(a)
                parts_h.append(h)
        assert np.array_equal(np.concatenate(parts_a), np.concatenate(whole_a))
        assert np.array_equal(np.concatenate(parts_h), np.concatenate(whole_h))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_ascii_sequence_shards_with_halo(kc, sh, world):
    rng = np.random.default_rng(7 + world)
    n, k = 1000, 31
    s = "".join(rng.choice(list("ACGT"), size=n)).encode()
    a, _, h = ko.ascii_iterate(s, k, ko.CANON, want_hash=True)
    plan = sh.plan_sequence_shards(n, k, 8, world)
    assert sum(p.n_windows for p in plan) == n - k + 1
    if world == 2:  # the advisor's numbers: rank 1 starts at byte 485 and needs 515 bytes
        assert (plan[1].word0, plan[1].n_words, plan[1].first_symbol_offset) == (485, 515, 0)
    pa, ph = [], []
    data = np.frombuffer(s, dtype=np.uint8)
    for g in range(world):
        rs, base = sh.sequence_shard(8, data, n, k, world, g)
        assert base == plan[g].window0
        view = bytes(rs.words[rs.first_symbol_offset: rs.first_symbol_offset + rs.uniform_len])
        x, _, y = ko.ascii_iterate(view, k, ko.CANON, want_hash=True)
        assert x.shape[0] == plan[g].n_windows
        pa.append(x)
        ph.append(y)
    assert np.array_equal(np.concatenate(pa), a) and np.array_equal(np.concatenate(ph), h)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, bits, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kmerscuda as kc
    from kmerscuda import sharding
    k = 31
    rng = np.random.default_rng(1234)  # every rank builds the same full set, then takes its shard
    n_reads, length, stride = 4000, 150, 5
    words = rng.integers(0, 2**64, size=n_reads * stride, dtype=np.uint64)
    rs = kc.ReadSet(2, words, n_reads, uniform_len=length, uniform_stride_words=stride)
    sub = sharding.read_shard(rs, world, rank)
    _, _, h, _ = ko.batch_iterate(sub.words, sub.n_seqs, k, ko.CANON, uniform_len=length, uniform_stride=stride, want_hash=True)
    table = torch.from_numpy(np.bincount((h >> np.uint64(64 - bits)).astype(np.int64), minlength=1 << bits).astype(np.int32))
    sharding.allreduce_table(table)
    counts = sharding.gather_counts(int(h.size))
    if rank == 0:
        _, _, hh, _ = ko.batch_iterate(words, n_reads, k, ko.CANON, uniform_len=length, uniform_stride=stride, want_hash=True)
        want = np.bincount((hh >> np.uint64(64 - bits)).astype(np.int64), minlength=1 << bits).astype(np.int32)
        ret["ok"] = bool(np.array_equal(table.numpy(), want)) and sum(counts) == hh.size and len(counts) == world
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_table_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 16, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert ret.get("ok") is True

"""Multi-GPU sharding of the k-mer extraction path: one process per GPU.

Every window depends only on its own K symbols, so the path shards with NO data-path collective:

  * read sets     : contiguous read ranges per rank, balanced by total symbols; reads are never
                    split; the concatenation of the ranks' outputs in rank order is the single-GPU
                    (= reference) order.
  * one sequence  : rank g owns the window starts [g*ceil(n/G), (g+1)*ceil(n/G)) and reads those
                    symbols plus a (K-1)-symbol halo; emitted indices are made global through
                    kmc_out.index_base.

The only collective of the whole path is the sum of the optional per-GPU bucket-count tables
(`allreduce_table`: torch.distributed all_reduce(SUM) -- NCCL over NVLink on GPUs, gloo in the CPU
tests).  The table lives in device memory owned by torch; libkmerscuda accumulates into it through
its raw pointer.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .api import ReadSet


@dataclass(frozen=True)
class ReadShard:
    rank: int
    seq0: int          # first read of the shard
    n_seqs: int
    word0: int         # first word of the shard in the parent buffer
    n_words: int


def symbols_per_unit(bits: int) -> int:
    """Symbols per offset unit of a source: 32 / 16 per 64-bit LongSequence word (2- / 4-bit alphabets); ASCII
    sources (bits == 8) count BYTES in every "word" quantity (include/kmerscuda.h, kmc_seqs), so one symbol per unit
    -- the same rule as host_pipeline.cu and tiles_upper_bound."""
    return 1 if bits == 8 else 64 // bits


def plan_read_shards(rs: ReadSet, world: int) -> list[ReadShard]:
    """Contiguous read ranges, balanced by symbols (uniform sets: by reads)."""
    n = rs.n_seqs
    bits = rs.bits
    spw = symbols_per_unit(bits)
    if rs.seq_len is None:
        bounds = [(n * g) // world for g in range(world + 1)]
    else:
        cum = np.concatenate([[0], np.cumsum(rs.seq_len.astype(np.int64))])
        total = int(cum[-1])
        bounds = [0]
        for g in range(1, world):
            bounds.append(int(np.searchsorted(cum, total * g / world, side="left")))
        bounds.append(n)
        bounds = [min(max(b, 0), n) for b in bounds]
        for i in range(1, len(bounds)):
            bounds[i] = max(bounds[i], bounds[i - 1])
    shards = []
    for g in range(world):
        s0, s1 = bounds[g], bounds[g + 1]
        if s1 == s0:
            shards.append(ReadShard(g, s0, 0, 0, 0))
            continue
        if rs.seq_word_offset is not None:
            w0 = int(rs.seq_word_offset[s0])
            last_len = int(rs.seq_len[s1 - 1]) if rs.seq_len is not None else rs.uniform_len
            w1 = int(rs.seq_word_offset[s1 - 1]) + (rs.first_symbol_offset + last_len + spw - 1) // spw
        else:
            w0, w1 = s0 * rs.uniform_stride_words, s1 * rs.uniform_stride_words
        shards.append(ReadShard(g, s0, s1 - s0, w0, max(0, min(w1, rs.words.size) - w0)))
    return shards


def read_shard(rs: ReadSet, world: int, rank: int) -> ReadSet:
    """The sub-ReadSet rank `rank` extracts (a view of the parent's buffers, offsets rebased)."""
    sh = plan_read_shards(rs, world)[rank]
    words = rs.words[sh.word0: sh.word0 + sh.n_words]
    if sh.n_seqs == 0:
        return ReadSet(rs.bits, np.zeros(0, np.uint64), 0, uniform_len=rs.uniform_len,
                       uniform_stride_words=rs.uniform_stride_words, first_symbol_offset=rs.first_symbol_offset)
    off = None if rs.seq_word_offset is None else (rs.seq_word_offset[sh.seq0: sh.seq0 + sh.n_seqs] - np.uint64(sh.word0))
    ln = None if rs.seq_len is None else rs.seq_len[sh.seq0: sh.seq0 + sh.n_seqs]
    return ReadSet(rs.bits, words, sh.n_seqs, seq_word_offset=off, seq_len=ln, uniform_len=rs.uniform_len,
                   uniform_stride_words=rs.uniform_stride_words, first_symbol_offset=rs.first_symbol_offset)


@dataclass(frozen=True)
class SequenceShard:
    rank: int
    window0: int              # first window (0-based start symbol) the rank owns
    n_windows: int
    word0: int                # first word to upload
    n_words: int
    first_symbol_offset: int  # of the shard view inside word0
    length: int               # symbols in the view: n_windows + K - 1 (the halo)
    index_base: int           # added to emitted 1-based indices (= window0)


def plan_sequence_shards(length: int, K: int, bits: int, world: int, first_symbol_offset: int = 0) -> list[SequenceShard]:
    """Window ranges of ONE long sequence with a K-1-symbol halo per rank."""
    spw = symbols_per_unit(bits)
    n = max(0, length - K + 1)
    per = (n + world - 1) // world if n else 0
    out = []
    for g in range(world):
        w0 = min(n, g * per)
        w1 = min(n, (g + 1) * per)
        nw = w1 - w0
        if nw == 0:
            out.append(SequenceShard(g, w0, 0, 0, 0, 0, 0, w0))
            continue
        sym0 = first_symbol_offset + w0
        sym1 = sym0 + nw + K - 1
        word0 = sym0 // spw
        out.append(SequenceShard(g, w0, nw, word0, (sym1 + spw - 1) // spw - word0, sym0 % spw, nw + K - 1, w0))
    return out


def sequence_shard(bits: int, words: np.ndarray, length: int, K: int, world: int, rank: int,
                   first_symbol_offset: int = 0):
    """(ReadSet view, index_base) for rank `rank` of a single long sequence."""
    sh = plan_sequence_shards(length, K, bits, world, first_symbol_offset)[rank]
    w = words[sh.word0: sh.word0 + sh.n_words]
    rs = ReadSet(bits, w, 1, uniform_len=sh.length, uniform_stride_words=max(int(w.size), 1),
                 first_symbol_offset=sh.first_symbol_offset)
    if sh.n_windows == 0:
        rs = ReadSet(bits, np.zeros(1, np.uint64), 1, uniform_len=0, uniform_stride_words=1)
    return rs, sh.index_base


def allreduce_table(table, group=None):
    """Sum the per-rank bucket-count tables in place (the path's only collective).  `table` is a
    torch tensor (CUDA: NCCL over NVLink; CPU: gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    return table


def count_and_merge_table(ctx, desc, K: int, bucket_bits: int, table, n_parts: int = 8, group=None, comm_stream=None):
    """Bucket-count this rank's reads into `table` (a zeroed CUDA int32 torch tensor of 2^bucket_bits entries) and
    sum the tables of all ranks, with the merge overlapping the count: kmc_bucket_count_async records one
    event per finished range of the table, and every range is all-reduced on `comm_stream` as soon as its
    event has fired, while the context's stream is still counting the later ranges.  Returns the number of
    k-mers this rank counted.  `ctx`'s stream must be torch's current stream.  With one rank (or an
    L2-sized table, which is final only at the end) this degenerates to count, then merge."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import _abi
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    main = torch.cuda.current_stream()
    # torch creates the CUDA event on first use: record once so that the handle exists
    events = [torch.cuda.Event() for _ in range(n_parts)]
    for e in events:
        e.record(main)
    handles = (C.c_void_p * n_parts)(*[C.c_void_p(e.cuda_event) for e in events])
    res = _abi.kmc_result()
    ctx._check(ctx.lib.kmc_bucket_count_async(ctx.handle, C.byref(desc), K, bucket_bits, table.data_ptr(), n_parts, handles,
                                              C.byref(res)))
    if multi:
        comm = comm_stream or torch.cuda.Stream()
        part = table.numel() // n_parts
        with torch.cuda.stream(comm):
            for i, e in enumerate(events):
                comm.wait_event(e)
                dist.all_reduce(table[i * part:(i + 1) * part], op=dist.ReduceOp.SUM, group=group)
        main.wait_stream(comm)
        table.record_stream(comm)
    return int(res.n_written)


def gather_counts(n_local: int, group=None) -> Optional[list[int]]:
    """Element counts of every rank (for the global offsets of variable-length outputs:
    UnambiguousKmers); a host-side all_gather of one integer, not a data-path collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [n_local]
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([n_local], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [int(x.item()) for x in out]

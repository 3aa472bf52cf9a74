#!/usr/bin/env python
"""C5: canonical 31-mer hash-bucket count table over 150 bp reads, one process per GPU, tables
merged with the path's only collective (NCCL all-reduce over NVLink).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5.py [--reads-per-gpu 25000000] [--bits 28]

Rank 0 prints one JSON line: whole-job k-mers/s (count + merge), the count and merge times
(CUDA events, max over ranks) and a parity check of the merged table total."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kmers.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads-per-gpu", type=int, default=25_000_000)
    ap.add_argument("--bits", type=int, default=28)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--overlap", type=int, default=0,
                    help="N > 0: kmc_bucket_count_async with N progress events, finished table ranges all-reduced while "
                         "later ranges are still being counted (sharding.count_and_merge_table)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import kmerscuda as kc
    from kmerscuda import _abi, sharding
    ctx = kc.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    k, length, stride = 31, 150, 5
    n_reads = args.reads_per_gpu
    wpr = length - k + 1
    g = torch.Generator(device="cuda").manual_seed(439824 + rank)
    words = torch.randint(-2**63, 2**63 - 1, (n_reads * stride,), dtype=torch.int64, device="cuda", generator=g)
    words.view(n_reads, stride)[:, stride - 1] &= (1 << (2 * (length - 32 * (stride - 1)))) - 1
    desc = _abi.kmc_seqs(words.data_ptr(), words.numel(), n_reads, None, None, length, stride, 2, 0)
    table = torch.zeros(1 << args.bits, dtype=torch.int32, device="cuda")
    res = _abi.kmc_result()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    comm = torch.cuda.Stream()
    t_count, t_merge = [], []
    for step in range(args.steps + 2):
        table.zero_()
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        if args.overlap:
            sharding.count_and_merge_table(ctx, desc, k, args.bits, table, n_parts=args.overlap, comm_stream=comm)
            e1.record(stream)  # (count and merge are not separable here: merge_ms reads ~0)
        else:
            st = ctx.lib.kmc_bucket_count(ctx.handle, C.byref(desc), k, args.bits, table.data_ptr(), C.byref(res))
            assert st == 0, ctx.lib.kmc_last_error(ctx.handle)
            e1.record(stream)
            sharding.allreduce_table(table)
        e2.record(stream)
        barrier()
        if step >= 2:
            t_count.append(e0.elapsed_time(e1))
            t_merge.append(e1.elapsed_time(e2))
    total = int(table.sum(dtype=torch.int64).item())
    t = torch.tensor([sum(t_count) / len(t_count), sum(t_merge) / len(t_merge)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n = n_reads * wpr * world
        ms_c, ms_m = float(t[0]), float(t[1])
        tb = 4 << args.bits
        print(json.dumps({
            "case": f"C5 canonical 31-mer bucket-count table, B={args.bits}, {n_reads:,} x 150 bp reads per GPU, {world} GPU(s)",
            "n_gpus": world, "kmers": n, "count_ms": ms_c, "merge_ms": ms_m,
            "kmers_per_s": n / ((ms_c + ms_m) / 1e3), "kmers_per_s_count_only": n / (ms_c / 1e3),
            "table_bytes": tb, "allreduce_busbw_GBps": (2 * (world - 1) / world * tb / (ms_m / 1e3) / 1e9) if world > 1 and not args.overlap else None,
            "merged_total_ok": total == n, "overlap_parts": args.overlap, "total_ms": ms_c + ms_m, "collective": "torch.distributed all_reduce(SUM), NCCL" if world > 1 else None}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Gallery/document sharding across the GPUs of one box: one process per GPU, queries replicated,
gallery columns split into contiguous balanced shards, exactly two collectives per scoring pass:

  1. all-reduce(MAX) of the per-query GT thresholds (4 bytes per query; the GT column lives on one shard only) — the tiny
     pre-pass of SURVEY 8e(i);
  2. ONE all-gather of each shard's bitmask columns (or, in LIST mode for galleries whose bitmask does not fit, of each
     shard's truncated above-GT column list); the counts are the popcounts of the merged rows, no second reduction.
When the retrieved items themselves are wanted, ONE all-gather of a fixed-size record per query — the shard's top-k
scores and global column ids, 8k bytes — followed by a local merge (mmx_merge_topk) gives the global top-k.

Counts and bit columns are additive / concatenable over gallery shards, so the merged structures are bit-identical
to an unsharded pass (tests/test_gpu_parity.py::test_dense_arbitrary_gt_and_sharded_gallery checks the arithmetic,
tests/test_sharded_gloo.py the collectives with world_size 2 on CPU). The functions below only move tensors: they
work on any torch.distributed backend (nccl on the B200s, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .ops import ProtocolRanks

SHARD_ALIGN = 1    # shards are balanced to one column; merge_ranks shifts the bit columns into place


def shard_bounds(n_cols: int, world: int, align: int = SHARD_ALIGN) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) column ranges. align = 1: sizes differ by at most one column; otherwise every boundary
    except the last is a multiple of `align`."""
    if align <= 1:
        base, extra = divmod(n_cols, world)
        out, a = [], 0
        for r in range(world):
            b = a + base + (1 if r < extra else 0)
            out.append((a, b))
            a = b
        return out
    per = -(-n_cols // world)
    per = -(-per // align) * align
    return [(min(r * per, n_cols), min((r + 1) * per, n_cols)) for r in range(world)]


def merge_thresholds(thr: torch.Tensor, group=None) -> torch.Tensor:
    """In-place max over ranks (a shard without the GT column contributes -inf)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(thr, op=dist.ReduceOp.MAX, group=group)
    return thr


def _all_gather_stacked(mine: torch.Tensor, group=None) -> torch.Tensor:
    """[world, *mine.shape]: every rank's equally shaped tensor, gathered straight into one buffer (NCCL: no staging copy;
    the list form is kept for the gloo host-logic tests)."""
    world = dist.get_world_size(group)
    mine = mine.contiguous()
    out = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
    if mine.is_cuda:
        dist.all_gather_into_tensor(out, mine, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), mine, group=group)
    return out


def merge_ranks(local: ProtocolRanks, bounds: List[Tuple[int, int]], group=None) -> ProtocolRanks:
    """ONE all-gather per scoring pass: every shard's bit columns travel once, each rank then shifts them to
    their global bit offsets, ORs them together and takes the merged rows' popcounts as the counts (mmx_place_bit_columns;
    the same result on every rank). Without bit columns (want_bits=False) the counts are all-reduced instead; in LIST mode
    the record is the shard's column list."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    cnt, bits = None, None
    if local.bits is not None:
        nq = local.bits.shape[0]
        wmax = max((e - b + 31) // 32 for b, e in bounds)
        mine = local.bits
        if mine.shape[1] != wmax:
            mine = torch.zeros((nq, wmax), dtype=local.bits.dtype, device=local.bits.device)
            mine[:, : local.bits.shape[1]] = local.bits
        gathered = _all_gather_stacked(mine, group)
        if gathered.is_cuda:
            from . import ops
            bits, cnt = ops.place_bit_columns(gathered, bounds)
        else:                                            # gloo host-logic tests: the same merge in torch
            bits = place_bit_columns(list(gathered.unbind(0)), bounds)
            cnt = _popcount_rows(bits)
    elif local.list_i is None:
        cnt = local.cnt_above.clone()
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    list_i = list_n = None
    if local.list_i is not None:
        # LIST mode: two all-gathers, both straight into the layout K3 reads — the shards' column lists [world, nq, cap]
        # (the bulk: no packing or unpacking copies of a 2 GB record array) and one small (exact count, cursor) pair per
        # query, whose first column summed over the shards is the global count (no separate all-reduce)
        list_i = _all_gather_stacked(local.list_i[0], group)
        meta = _all_gather_stacked(torch.stack([local.cnt_above, local.list_n[0]], dim=1), group)     # [world, nq, 2]
        cnt = meta[:, :, 0].sum(dim=0, dtype=torch.int32)
        list_n = meta[:, :, 1].contiguous()
    return ProtocolRanks(thr=local.thr, cnt_above=cnt, bits=bits, n_cols=bounds[-1][1], col2unit=local.col2unit,
                         q_unit=local.q_unit, list_i=list_i, list_n=list_n)


def _popcount_rows(bits: torch.Tensor) -> torch.Tensor:
    """Row popcounts of an int32 bitmask on the CPU (host-logic tests only; the GPU path counts inside the merge kernel)."""
    import numpy as np
    u8 = bits.contiguous().numpy().view(np.uint8)
    return torch.from_numpy(np.unpackbits(u8, axis=1).sum(axis=1).astype(np.int32))


def unlisted_round(local_resolve, pairs, mult_rows, group=None, src_rank: int = 0) -> torch.Tensor:
    """The exact recount of LIST-mode K3's unsettled (iteration, query) pairs across gallery shards. Rank 0 (which ran
    K3) passes the pairs and the multiplicity rows they refer to, every other rank None; EVERY rank calls this once
    per K3 call (rank 0 with pairs=None when nothing was left unsettled). The pairs and rows are broadcast, each
    rank recounts over its own columns (ops.dense_unlisted_resolver) and the partial sums are all-reduced."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_resolve(pairs, mult_rows) if pairs is not None else None
    rank = dist.get_rank(group)
    src = dist.get_global_rank(group, src_rank) if group is not None else src_rank
    dev = mult_rows.device if mult_rows is not None else _group_device(group)
    head = torch.zeros(3, dtype=torch.int64, device=dev)
    if rank == src_rank and pairs is not None:
        head = torch.tensor([pairs.shape[0], mult_rows.shape[0], mult_rows.shape[1]], dtype=torch.int64, device=dev)
    dist.broadcast(head, src=src, group=group)
    n, r, u = (int(x) for x in head.tolist())
    if n == 0:
        return None
    if rank != src_rank:
        pairs = torch.empty((n, 2), dtype=torch.int32, device=dev)
        mult_rows = torch.empty((r, u), dtype=torch.uint8, device=dev)
    pairs, mult_rows = pairs.contiguous(), mult_rows.contiguous()
    dist.broadcast(pairs, src=src, group=group)
    dist.broadcast(mult_rows, src=src, group=group)
    c = local_resolve(pairs, mult_rows)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return c


def _group_device(group=None) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def place_bit_columns(parts: List[torch.Tensor], bounds: List[Tuple[int, int]]) -> torch.Tensor:
    """Shard r's bit j is global column bounds[r][0] + j: shift every shard's words to their global bit offset and OR
    them together (a shard may start in the middle of a 32-column word). Bits beyond a shard's width are zero."""
    nq = parts[0].shape[0]
    n_cols = bounds[-1][1]
    out = torch.zeros((nq, (n_cols + 31) // 32 + 1), dtype=torch.int64, device=parts[0].device)
    for p, (b, e) in zip(parts, bounds):
        wl = (e - b + 31) // 32
        if wl == 0:
            continue
        w = p[:, :wl].to(torch.int64) & 0xFFFFFFFF
        w0, sh = b // 32, b % 32
        out[:, w0:w0 + wl] |= (w << sh) & 0xFFFFFFFF
        if sh:
            out[:, w0 + 1:w0 + wl + 1] |= w >> (32 - sh)
    out = out[:, : (n_cols + 31) // 32]
    out = torch.where(out >= (1 << 31), out - (1 << 32), out)          # back to the int32 bit pattern
    return out.to(parts[0].dtype).contiguous()


def gather_rows(host_rows: torch.Tensor, dev: torch.device, group=None) -> torch.Tensor:
    """A replicated host matrix onto every GPU at NVLink cost: rank r uploads rows [r n / N, (r + 1) n / N) over its own
    PCIe link and ONE all-gather spreads the slices (the 128 MB query matrix of configs[3] crossed PCIe eight times when
    every rank uploaded all of it)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1 or host_rows.is_cuda:
        return host_rows.to(dev, non_blocking=True)
    rank = dist.get_rank(group)
    n = host_rows.shape[0]
    per = -(-n // world)
    out = torch.empty((world * per,) + tuple(host_rows.shape[1:]), dtype=host_rows.dtype, device=dev)
    a, b = min(n, rank * per), min(n, (rank + 1) * per)
    mine = out[rank * per: rank * per + per]
    if b > a:
        mine[: b - a].copy_(host_rows[a:b], non_blocking=True)
    dist.all_gather(list(out.view(world, per, *host_rows.shape[1:]).unbind(0)), mine, group=group)
    return out[:n]


def list_boot_fn(recount, group=None, stats: Optional[dict] = None):
    """boot_fn for bootstrap_stats in LIST mode (split over ranks): K3 on the lists for this rank's iteration block, then
    ONE tiny all-reduce asks whether any rank has pairs its truncated lists could not settle (probability ~1e-17 each at
    cap 64); only then every rank's pairs are recounted exactly, rank by rank, over all gallery shards (unlisted_round)."""
    from . import ops

    def fn(ranks: ProtocolRanks, mult: torch.Tensor) -> torch.Tensor:
        correct, pairs = ops.boot_counts_lists_raw(ranks, mult)
        n_mine = 0 if pairs is None else int(pairs.shape[0])
        flag = torch.tensor([n_mine], dtype=torch.int64, device=mult.device)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            tot = flag.clone()
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
            n_all = int(tot.item())
        else:
            n_all = n_mine
        if stats is not None:
            stats["unresolved_pairs"] = n_all
        if n_all:
            rank = dist.get_rank(group) if world > 1 else 0
            for r in range(world):
                mine = rank == r and pairs is not None
                req = ops.unlisted_request(pairs, mult) if mine else (None, None)
                c = unlisted_round(recount, req[0], req[1], group, src_rank=r)
                if mine:
                    ops.apply_unlisted(correct, pairs, c, ranks, mult)
        return correct
    return fn


def stats_block(n_iterations: int, group=None) -> Tuple[int, int]:
    """[b0, b1): the bootstrap iterations this rank takes when the statistics are split over the ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    per = -(-n_iterations // world)
    return min(n_iterations, rank * per), min(n_iterations, (rank + 1) * per)


def bootstrap_stats(ranks: ProtocolRanks, sample_indices: torch.Tensor, ci_idx: torch.Tensor, q_unit: torch.Tensor,
                    n_units: int, group=None, boot_fn=None, mark=None, n_iterations: Optional[int] = None):
    """The statistics tail of a sharded pass — K3 over the bootstrap iterations, K4 over the CI resample rows — split
    over the ranks instead of serialised on rank 0: every rank holds the merged rank structures and the same numpy
    stream (sample_indices int32 [B, n_units], ci_idx int32 [M, R, B]), takes its block of iterations / resample rows,
    and two small all-gathers put the pieces together. With n_iterations = B given, sample_indices holds only THIS
    rank's rows (stats_block(B): DeviceRng.randint(window=...) keeps just those). Returns (values fp64 [B, M], means fp64 [M, R]) on every rank,
    bit-identical to the single-GPU tail (integer counts; the means are computed row by row either way).
    boot_fn(ranks, mult) -> int32 [b, M] defaults to ops.boot_counts. mark(name) (optional) is called after every phase
    (api.run_retrieval_bootstrap records a CUDA event there when timings are asked for)."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    boot_fn = boot_fn or ops.boot_counts
    mark = mark or (lambda name: None)
    own_rows_only = n_iterations is not None
    B = n_iterations if own_rows_only else sample_indices.shape[0]
    per = -(-B // world)
    b0, b1 = min(B, rank * per), min(B, (rank + 1) * per)
    if own_rows_only and sample_indices.shape[0] != b1 - b0:
        raise ValueError("sample_indices must hold this rank's block of iterations (stats_block)")
    dev = sample_indices.device
    rec = torch.zeros((per, 4), dtype=torch.int64, device=dev)       # correct@1/5/10 + query copies drawn
    if b1 > b0:
        mult = ops.bincount(sample_indices if own_rows_only else sample_indices[b0:b1].contiguous(), n_units)
        mark("bincount")
        rec[: b1 - b0, :3] = boot_fn(ranks, mult)
        mark("bootstrap_counts")
        rec[: b1 - b0, 3] = (ops.query_copies_drawn(mult, q_unit) if mult.is_cuda      # (CPU: the gloo host-logic tests)
                             else mult[:, q_unit.long()].sum(dim=1, dtype=torch.int64))
        mark("query_copies_drawn")
    if world > 1:
        allrec = torch.empty((world, per, 4), dtype=torch.int64, device=dev)
        dist.all_gather(list(allrec.unbind(0)), rec, group=group)
        rec = allrec.view(world * per, 4)
        mark("counts_all_gather")
    rec = rec[:B]
    values = 100.0 * rec[:, :3].to(torch.float64) / rec[:, 3:4].to(torch.float64)         # main.py:358, IEEE division
    M, R, _ = ci_idx.shape
    perr = -(-R // world)
    r0, r1 = min(R, rank * perr), min(R, (rank + 1) * perr)
    means = torch.zeros((M, perr), dtype=torch.float64, device=dev)
    if r1 > r0:
        means[:, : r1 - r0] = ops.ci_means_batched(values.t().contiguous(), ci_idx[:, r0:r1].contiguous())
    if world > 1:
        allm = torch.empty((world, M, perr), dtype=torch.float64, device=dev)
        dist.all_gather(list(allm.unbind(0)), means, group=group)
        means = allm.permute(1, 0, 2).reshape(M, world * perr)
    mark("ci_means_and_all_gather")
    return values, means[:, :R].contiguous()


def merge_topk(top_s: torch.Tensor, top_i: torch.Tensor, group=None, merge_fn=None):
    """Global top-k from per-shard top-k lists with a single all-gather: every rank contributes one int32 record
    [nq, 2k] (score bit patterns + global column ids), then merges the world's records locally.
    `merge_fn(gathered_s, gathered_i, k)` defaults to the CUDA merge (ops.merge_topk); the gloo tests pass a CPU one."""
    k = top_s.shape[1]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return top_s, top_i
    world = dist.get_world_size(group)
    rec = torch.cat([top_s.contiguous().view(torch.int32), top_i], dim=1).contiguous()
    parts = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(parts, rec, group=group)
    gs = torch.cat([p[:, :k] for p in parts], dim=1).contiguous().view(torch.float32)
    gi = torch.cat([p[:, k:] for p in parts], dim=1).contiguous()
    if merge_fn is None:
        from . import ops
        merge_fn = ops.merge_topk
    return merge_fn(gs, gi, k)

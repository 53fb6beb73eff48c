"""Gallery/document sharding across the GPUs of one box: one process per GPU, queries replicated,
gallery columns split into contiguous balanced shards, exactly two tiny collectives per scoring pass:

  1. all-reduce(MAX) of the per-query GT thresholds (the GT column of a query lives on one shard only);
  2. all-reduce(SUM) of the strict-greater counts + all-gather of each shard's bitmask columns (or, in LIST mode for
     galleries whose bitmask does not fit, of each shard's truncated above-GT column list).
When the retrieved items themselves are wanted, ONE all-gather of a fixed-size record per query — the shard's top-k
scores and global column ids, 8k bytes — followed by a local merge (mmx_merge_topk) gives the global top-k.

Counts and bit columns are additive / concatenable over gallery shards, so the merged structures are bit-identical
to an unsharded pass (tests/test_gpu_parity.py::test_dense_arbitrary_gt_and_sharded_gallery checks the arithmetic,
tests/test_sharded_gloo.py the collectives with world_size 2 on CPU). The functions below only move tensors: they
work on any torch.distributed backend (nccl on the B200s, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

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


def merge_ranks(local: ProtocolRanks, bounds: List[Tuple[int, int]], group=None) -> ProtocolRanks:
    """Sum the counts, concatenate the bit columns of all shards (same result on every rank)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    cnt = local.cnt_above.clone()
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    bits = None
    if local.bits is not None:
        nq = local.bits.shape[0]
        wmax = max((e - b + 31) // 32 for b, e in bounds)
        mine = torch.zeros((nq, wmax), dtype=local.bits.dtype, device=local.bits.device)
        mine[:, : local.bits.shape[1]] = local.bits
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        bits = place_bit_columns(parts, bounds)
    list_i = list_n = None
    if local.list_i is not None:
        # LIST mode: ONE all-gather of a fixed-size record per query — the shard's `cap` column ids and its cursor
        cap = local.list_i.shape[2]
        rec = torch.cat([local.list_i[0], local.list_n[0][:, None]], dim=1).contiguous()      # [nq, cap + 1]
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec, group=group)
        list_i = torch.stack([p[:, :cap] for p in parts]).contiguous()
        list_n = torch.stack([p[:, cap] for p in parts]).contiguous()
    return ProtocolRanks(thr=local.thr, cnt_above=cnt, bits=bits, n_cols=bounds[-1][1], col2unit=local.col2unit,
                         q_unit=local.q_unit, list_i=list_i, list_n=list_n)


def unlisted_round(local_resolve, pairs, mult_rows, group=None) -> torch.Tensor:
    """The exact recount of LIST-mode K3's unsettled (iteration, query) pairs across gallery shards. Rank 0 (which ran
    K3) passes the pairs and the multiplicity rows they refer to, every other rank None; EVERY rank calls this once
    per K3 call (rank 0 with pairs=None when nothing was left unsettled). The pairs and rows are broadcast, each
    rank recounts over its own columns (ops.dense_unlisted_resolver) and the partial sums are all-reduced."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_resolve(pairs, mult_rows) if pairs is not None else None
    rank = dist.get_rank(group)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dev = mult_rows.device if mult_rows is not None else _group_device(group)
    head = torch.zeros(3, dtype=torch.int64, device=dev)
    if rank == 0 and pairs is not None:
        head = torch.tensor([pairs.shape[0], mult_rows.shape[0], mult_rows.shape[1]], dtype=torch.int64, device=dev)
    dist.broadcast(head, src=src, group=group)
    n, r, u = (int(x) for x in head.tolist())
    if n == 0:
        return None
    if rank != 0:
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

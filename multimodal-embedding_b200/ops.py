"""Device-level wrappers: one Python function per libmmx entry point, taking CUDA tensors.

torch is plumbing here (device memory, streams); all arithmetic happens inside libmmx.so. Every wrapper raises when
the library or a CUDA device is missing — there is no eager/CPU fallback."""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

MODE_THRESHOLD, MODE_COUNT = 0, 1
LIST_LEN = 32          # entries kept per above-GT list (rows with more fall back to the exact bitmask walk)
LIST_CAP = 64          # LIST mode (no bitmask): entries kept per query and shard; P(64 draws of Poisson(1) sum < 10) ~ 1e-17
UNRESOLVED_CAP = 1 << 16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MmxError("libmmx operates on CUDA tensors only (no CPU fallback)")


def csr_from_sets(gt_sets: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """Ground-truth sets -> CSR (int32 offsets [n+1], int32 indices)."""
    off = np.zeros(len(gt_sets) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(g) for g in gt_sets])
    idx = np.fromiter((int(c) for g in gt_sets for c in g), dtype=np.int32, count=int(off[-1]))
    return off, idx


@dataclass
class ProtocolRanks:
    """Result of one scoring pass for one protocol, resident on the device.

    thr [nq] fp32 GT threshold, cnt_above [nq] int32 strict-greater count (= rank-1 of the best GT item),
    bits [nq, wpr] uint32-as-int32 above-GT bitmask over gallery columns, col2unit / q_unit map columns / queries to
    bootstrap units (images); None = identity."""
    thr: torch.Tensor
    cnt_above: torch.Tensor
    bits: Optional[torch.Tensor]
    n_cols: int
    col2unit: Optional[torch.Tensor] = None
    q_unit: Optional[torch.Tensor] = None
    scores: Optional[torch.Tensor] = None
    ties: Optional[torch.Tensor] = None       # [nq] int32 non-GT columns scoring EXACTLY the threshold (tie class T2)
    t6: Optional[dict] = None                 # ColPali: how many scores depend on pad_sequence's zero row (tie class T6)
    top_s: Optional[torch.Tensor] = None      # [nq, k] fp32 top-k scores (KT), higher first, ties by lower column
    top_i: Optional[torch.Tensor] = None      # [nq, k] int32 global column ids, -1 = padding
    # LIST mode (galleries whose bitmask does not fit): bits is None and the above-GT sets are kept as
    # per-shard column lists [n_shards, nq, cap] (global ids, first arrivals) with cursors [n_shards, nq]
    list_i: Optional[torch.Tensor] = None
    list_n: Optional[torch.Tensor] = None

    @property
    def nq(self) -> int:
        return int(self.cnt_above.shape[0])

    def recall(self) -> dict:
        """100.0 * correct / n exactly as main.py:358 (python float arithmetic on integer counts)."""
        cnt = self.cnt_above
        n = self.nq
        correct = torch.stack([(cnt < k).sum() for k in (1, 5, 10)]).tolist()      # one device->host read, no upload
        return {k: 100.0 * c / n for k, c in zip((1, 5, 10), correct)}


MAXSIM_WORKSPACE_LIMIT = 16 << 30   # bytes; larger problems are scored in blocks of whole queries


def _maxsim_call(q_tok, q_off, nq, d_tok, d_off, nd, d_zero, out) -> None:
    need = ctypes.c_size_t()
    h = _lib.lib()
    _lib.check(h.mmx_maxsim_workspace_bytes(q_tok.shape[0], nd, d_tok.shape[0], ctypes.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=q_tok.device)
    _lib.check(h.mmx_maxsim_scores(q_tok.data_ptr(), q_off.data_ptr(), nq, q_tok.shape[0], d_tok.data_ptr(),
                                   d_off.data_ptr(), nd, d_tok.shape[0], q_tok.shape[1] if q_tok.dim() == 2 else 0,
                                   _ptr(d_zero), out.data_ptr(), out.stride(0), ws.data_ptr(), need.value, _stream()))


def maxsim_scores(q_tok: torch.Tensor, q_off: torch.Tensor, d_tok: torch.Tensor, d_off: torch.Tensor,
                  d_zero: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K2. q_tok [n_qtok,128] / d_tok [n_dtok,128] bf16 CUDA, offsets int32 CUDA -> fp32 [nq, nd] CUDA.

    The kernel's workspace holds one partial sum per (128-row query tile, document) for queries that span several
    tiles; when that exceeds MAXSIM_WORKSPACE_LIMIT (long queries x many documents: the I2T orientation at scale) the
    queries are scored in blocks of whole queries (one small device->host read of the offsets)."""
    _need_cuda(q_tok, q_off, d_tok, d_off, d_zero)
    if q_tok.dtype != torch.bfloat16 or d_tok.dtype != torch.bfloat16:
        raise TypeError("MaxSim kernel takes bf16 token embeddings")
    q_tok, d_tok = q_tok.contiguous(), d_tok.contiguous()
    nq, nd = q_off.numel() - 1, d_off.numel() - 1
    if out is None:
        out = torch.empty((nq, nd), dtype=torch.float32, device=q_tok.device)
    if nq == 0 or nd == 0:
        return out                                         # an empty shard (a rank that only helps with the statistics)
    need = ctypes.c_size_t()
    _lib.check(_lib.lib().mmx_maxsim_workspace_bytes(q_tok.shape[0], nd, d_tok.shape[0], ctypes.byref(need)))
    if need.value <= MAXSIM_WORKSPACE_LIMIT or nq <= 1:
        _maxsim_call(q_tok, q_off, nq, d_tok, d_off, nd, d_zero, out)
        return out
    off_h = q_off.cpu().numpy().astype(np.int64)
    rows_per_block = max(128, int(MAXSIM_WORKSPACE_LIMIT // 2 // max(1, nd * 4)) * 128)
    a = 0
    while a < nq:
        b = int(np.searchsorted(off_h, off_h[a] + rows_per_block, side="right")) - 1
        b = min(nq, max(b, a + 1))
        sub_off = (q_off[a:b + 1] - q_off[a]).contiguous()
        _maxsim_call(q_tok[int(off_h[a]):int(off_h[b])], sub_off, b - a, d_tok, d_off, nd, d_zero, out[a:b])
        a = b
    return out


MAXSIM_RANK_BLOCK_BYTES = 1 << 30   # score block of the fused MaxSim ranking (queries per block = this / (4 n_docs))


def maxsim_rank(q_tok: torch.Tensor, q_lens, d_tok: torch.Tensor, d_off: torch.Tensor, gt_off: torch.Tensor,
                gt_idx: torch.Tensor, d_zero: Optional[torch.Tensor] = None, want_bits: bool = True, g_base: int = 0,
                block_bytes: Optional[int] = None) -> ProtocolRanks:
    """K2 + KR without the [n_queries, n_docs] score matrix (SURVEY 8b ``mmx_maxsim_rank``; main.py:467 + :331-358): the
    queries are scored in blocks of whole queries into ONE reusable block buffer (<= block_bytes), each block is ranked
    at once — thresholds from its ground-truth columns, then strict-greater counts and bitmask rows — and its scores are
    dropped. A 25 000 x 100 000 problem needs 1 GB of scores at a time instead of 10 GB; the patches are re-read once
    per block (12 x 26 GB = 50 ms of HBM time against 11 s of tensor work). Same kernels; a block's queries sit at other
    rows of the 128-row tiles than in a whole-matrix call (another fixed fp32 summation tree), so scores can differ in the
    last bit — thresholds and counts are taken from the SAME block scores, the ranking is self-consistent. For an
    UNSHARDED document set: both ranking passes run on a block while its scores
    exist, which needs every ground-truth column to be local (sharded callers keep their 1/N-size matrix instead).
    q_lens: host int array of tokens per query (the blocks are cut on the host)."""
    _need_cuda(q_tok, d_tok, d_off, gt_off, gt_idx, d_zero)
    lens = np.asarray(q_lens, dtype=np.int64)
    nq, nd = len(lens), d_off.numel() - 1
    off_h = np.zeros(nq + 1, dtype=np.int64)
    off_h[1:] = np.cumsum(lens)
    dev = q_tok.device
    thr, cnt, bits, wpr = _alloc_rank(nq, nd, dev, want_bits)
    per = max(1, int((block_bytes or MAXSIM_RANK_BLOCK_BYTES) // max(1, 4 * nd)))
    buf = torch.empty((min(per, nq), nd), dtype=torch.float32, device=dev)
    h = _lib.lib()
    q_off_all = torch.from_numpy(off_h.astype(np.int32)).to(dev)
    for a in range(0, nq, per):
        b = min(nq, a + per)
        S = buf[: b - a]
        sub_off = (q_off_all[a:b + 1] - int(off_h[a])).contiguous()
        maxsim_scores(q_tok[int(off_h[a]):int(off_h[b])], sub_off, d_tok, d_off, d_zero, out=S)
        # this block's rows of the GT structure: offsets rebased, column ids unchanged
        g0 = gt_off[a:b + 1]
        go = (g0 - g0[:1]).contiguous()
        gi = gt_idx[int(g0[0].item()):int(g0[-1].item())].contiguous() if a or b < nq else gt_idx
        for mode in (MODE_THRESHOLD, MODE_COUNT):
            _lib.check(h.mmx_rank_from_scores(S.data_ptr(), b - a, nd, S.stride(0), go.data_ptr(), gi.data_ptr(), g_base, mode,
                                              thr[a:b].data_ptr(), cnt[a:b].data_ptr() if mode else None,
                                              bits[a:b].data_ptr() if (mode and bits is not None) else None, wpr, _stream()))
    return ProtocolRanks(thr=thr, cnt_above=cnt, bits=bits, n_cols=nd)


def _alloc_rank(nq: int, n_cols: int, device, want_bits: bool):
    thr = torch.full((nq,), float("-inf"), dtype=torch.float32, device=device)
    cnt = torch.zeros((nq,), dtype=torch.int32, device=device)
    wpr = (n_cols + 31) // 32
    bits = torch.empty((nq, wpr), dtype=torch.int32, device=device) if want_bits else None
    return thr, cnt, bits, wpr


def dense_rank(Q: torch.Tensor, G: torch.Tensor, gt_off: torch.Tensor, gt_idx: torch.Tensor, want_bits: bool = True,
               want_scores: bool = False, g_base: int = 0, thr: Optional[torch.Tensor] = None,
               threshold_only: bool = False, top_k: Optional[int] = None) -> ProtocolRanks:
    """K1. Q [nq,d], G [ng,d] bf16 CUDA; GT CSR int32 CUDA (global column ids) -> thresholds, counts, bitmask.

    With `thr` given the threshold pass is skipped (multi-GPU: thresholds were max-reduced over shards first);
    `threshold_only` stops after the threshold pass. `top_k` also returns the top-k lists (``top_s`` / ``top_i``): the
    COUNT pass stores the group maxima on its way, so they cost one more GEMM pass, not two."""
    _need_cuda(Q, G, gt_off, gt_idx, thr)
    if Q.dtype != torch.bfloat16 or G.dtype != torch.bfloat16:
        raise TypeError("dense rank kernel takes bf16 embeddings")
    Q, G = Q.contiguous(), G.contiguous()
    nq, d = Q.shape
    ng = G.shape[0]
    h = _lib.lib()
    thr0, cnt, bits, wpr = _alloc_rank(nq, ng, Q.device, want_bits and not threshold_only)
    scores = torch.empty((nq, ng), dtype=torch.float32, device=Q.device) if want_scores else None
    if thr is None:
        thr = thr0
        need = ctypes.c_size_t()
        _lib.check(h.mmx_dense_workspace_bytes(nq, ng, ctypes.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=Q.device)
        _lib.check(h.mmx_dense_rank(Q.data_ptr(), nq, G.data_ptr(), ng, d, gt_off.data_ptr(), gt_idx.data_ptr(), g_base,
                                    MODE_THRESHOLD, thr.data_ptr(), None, None, 0, None, 0, ws.data_ptr(), need.value,
                                    _stream()))
    top_s = top_i = None
    if not threshold_only and top_k is not None and not want_scores:
        need = ctypes.c_size_t()
        _lib.check(h.mmx_dense_topk_workspace_bytes(nq, ng, ctypes.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=Q.device)
        top_s = torch.empty((nq, top_k), dtype=torch.float32, device=Q.device)
        top_i = torch.empty((nq, top_k), dtype=torch.int32, device=Q.device)
        _lib.check(h.mmx_dense_rank_topk(Q.data_ptr(), nq, G.data_ptr(), ng, d, gt_off.data_ptr(), gt_idx.data_ptr(),
                                         g_base, thr.data_ptr(), cnt.data_ptr(), _ptr(bits), wpr, top_k, top_s.data_ptr(),
                                         top_i.data_ptr(), ws.data_ptr(), need.value, _stream()))
    elif not threshold_only:
        _lib.check(h.mmx_dense_rank(Q.data_ptr(), nq, G.data_ptr(), ng, d, gt_off.data_ptr(), gt_idx.data_ptr(), g_base,
                                    MODE_COUNT, thr.data_ptr(), cnt.data_ptr(), _ptr(bits), wpr, _ptr(scores),
                                    ng if want_scores else 0, None, 0, _stream()))
        if top_k is not None:
            top_s, top_i = topk_from_scores(scores, top_k, g_base)
    return ProtocolRanks(thr=thr, cnt_above=cnt, bits=bits, n_cols=ng, scores=scores, top_s=top_s, top_i=top_i)


def dense_rank_lists(Q: torch.Tensor, G: torch.Tensor, gt_off: torch.Tensor, gt_idx: torch.Tensor, g_base: int = 0,
                     thr: Optional[torch.Tensor] = None, cap: int = LIST_CAP) -> ProtocolRanks:
    """K1 in LIST mode: thresholds (unless given: shard-merged), exact counts, and per query the first `cap` columns
    to score above the threshold instead of the [nq, ng] bitmask — for galleries where nq * ng bits do not fit."""
    if thr is None:
        thr = dense_rank(Q, G, gt_off, gt_idx, want_bits=False, g_base=g_base, threshold_only=True).thr
    _need_cuda(Q, G, thr)
    if Q.dtype != torch.bfloat16 or G.dtype != torch.bfloat16:
        raise TypeError("dense rank kernel takes bf16 embeddings")
    Q, G = Q.contiguous(), G.contiguous()
    nq, d = Q.shape
    ng = G.shape[0]
    cnt = torch.empty((nq,), dtype=torch.int32, device=Q.device)
    list_i = torch.empty((1, nq, cap), dtype=torch.int32, device=Q.device)
    list_n = torch.empty((1, nq), dtype=torch.int32, device=Q.device)
    _lib.check(_lib.lib().mmx_dense_rank_lists(Q.data_ptr(), nq, G.data_ptr(), ng, d, g_base, thr.data_ptr(),
                                               cnt.data_ptr(), list_i.data_ptr(), list_n.data_ptr(), cap, _stream()))
    return ProtocolRanks(thr=thr, cnt_above=cnt, bits=None, n_cols=ng, list_i=list_i, list_n=list_n)


def dense_unlisted_resolver(Q: torch.Tensor, G: torch.Tensor, thr: torch.Tensor,
                            col2unit: Optional[torch.Tensor] = None, g_base: int = 0):
    """Exact recount for the (iteration, query) pairs LIST-mode K3 could not settle from a truncated list. Returns
    ``resolve(pairs[n, 2], mult_rows[r, n_units]) -> c int64 [n]`` with pairs[:, 0] an index into mult_rows and
    c = sum over THIS shard's columns of mult_rows[b, unit(col)] * [S[q, col] > thr[q]] — the scores come from K1's own
    MMA path (a few materialised rows at a time), so the comparison is the one the COUNT pass made. Partial sums of
    gallery shards add (sharded.unlisted_round)."""
    ng = G.shape[0]
    cols = torch.arange(g_base, g_base + ng, device=G.device)
    units = cols if col2unit is None else col2unit[cols].long()

    def resolve(pairs: torch.Tensor, mult_rows: torch.Tensor) -> torch.Tensor:
        c = torch.zeros(pairs.shape[0], dtype=torch.int64, device=G.device)
        if ng == 0 or pairs.shape[0] == 0:
            return c
        order = torch.argsort(pairs[:, 1], stable=True)
        q_sorted = pairs[order, 1].long().contiguous()
        qs = torch.unique(q_sorted)
        rows_per_block = max(1, min(64, (256 << 20) // (4 * ng)))
        for i in range(0, qs.numel(), rows_per_block):
            blk = qs[i:i + rows_per_block]
            above = dense_scores(Q[blk].contiguous(), G) > thr[blk][:, None]             # [rows, ng]
            lo = int(torch.searchsorted(q_sorted, blk[0]).item())
            hi = int(torch.searchsorted(q_sorted, blk[-1], right=True).item())
            for j in range(lo, hi, 64):
                sel = order[j:min(j + 64, hi)]
                row_of = torch.searchsorted(blk, pairs[sel, 1].long())
                m = mult_rows[pairs[sel, 0].long()][:, units]                             # [n, ng] uint8
                c[sel] = (m * above[row_of]).sum(dim=1, dtype=torch.int64)
        return c
    return resolve


def dense_scores(Q: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """K1 with the score output switched on and ranking neutralised (thr = +inf): fp32 CUDA [nq, ng] = Q @ G^T."""
    _need_cuda(Q, G)
    if Q.dtype != torch.bfloat16 or G.dtype != torch.bfloat16:
        raise TypeError("dense score kernel takes bf16 embeddings")
    Q, G = Q.contiguous(), G.contiguous()
    nq, d = Q.shape
    ng = G.shape[0]
    thr = torch.full((nq,), float("inf"), dtype=torch.float32, device=Q.device)
    cnt = torch.zeros((nq,), dtype=torch.int32, device=Q.device)
    scores = torch.empty((nq, ng), dtype=torch.float32, device=Q.device)
    _lib.check(_lib.lib().mmx_dense_rank(Q.data_ptr(), nq, G.data_ptr(), ng, d, None, None, 0, MODE_COUNT,
                                         thr.data_ptr(), cnt.data_ptr(), None, 0, scores.data_ptr(), ng, None, 0,
                                         _stream()))
    return scores


def rank_from_scores(scores: torch.Tensor, gt_off: torch.Tensor, gt_idx: torch.Tensor, want_bits: bool = True,
                     g_base: int = 0, thr: Optional[torch.Tensor] = None, threshold_only: bool = False) -> ProtocolRanks:
    """KR. fp32 CUDA score matrix [nq, ng] (any row pitch, unit column stride) -> thresholds, counts, bitmask."""
    _need_cuda(scores, gt_off, gt_idx, thr)
    if scores.dtype != torch.float32:
        raise TypeError("rank_from_scores takes fp32 scores")
    if scores.dim() != 2 or (scores.shape[1] > 1 and scores.stride(1) != 1):
        scores = scores.contiguous()
    nq, ng = scores.shape
    ld = scores.stride(0) if nq > 1 else max(ng, 1)
    h = _lib.lib()
    thr0, cnt, bits, wpr = _alloc_rank(nq, ng, scores.device, want_bits and not threshold_only)
    if thr is None:
        thr = thr0
        _lib.check(h.mmx_rank_from_scores(scores.data_ptr(), nq, ng, ld, gt_off.data_ptr(), gt_idx.data_ptr(), g_base,
                                          MODE_THRESHOLD, thr.data_ptr(), None, None, 0, _stream()))
    if not threshold_only:
        _lib.check(h.mmx_rank_from_scores(scores.data_ptr(), nq, ng, ld, gt_off.data_ptr(), gt_idx.data_ptr(), g_base,
                                          MODE_COUNT, thr.data_ptr(), cnt.data_ptr(), _ptr(bits), wpr, _stream()))
    return ProtocolRanks(thr=thr, cnt_above=cnt, bits=bits, n_cols=ng)


def _just_below(thr: torch.Tensor) -> torch.Tensor:
    """The largest float below each threshold: `S > just_below(t)` is `S >= t` (-inf stays -inf)."""
    return torch.nextafter(thr, torch.full_like(thr, float("-inf")))


def _gt_slices(gt_off: torch.Tensor, gt_idx: torch.Tensor):
    """The c-th ground-truth column of every query as single-column CSR pairs (queries with fewer columns repeat
    their last one; queries without any are skipped by the kernels)."""
    sizes = (gt_off[1:] - gt_off[:-1]).long()
    nq = sizes.numel()
    for c in range(int(sizes.max().item()) if nq else 0):
        has = sizes > 0
        pos = gt_off[:-1].long() + torch.minimum(torch.full_like(sizes, c), (sizes - 1).clamp_min(0))
        idx_c = torch.where(has, gt_idx[pos.clamp_max(max(gt_idx.numel() - 1, 0))].long(), torch.zeros_like(pos)).to(torch.int32)
        off_c = torch.cat([torch.zeros(1, dtype=torch.int64, device=sizes.device), has.long().cumsum(0)]).to(torch.int32)
        yield c, sizes, off_c, idx_c[has].contiguous()


def exact_tie_counts(r: ProtocolRanks, gt_off: torch.Tensor, gt_idx: torch.Tensor, *, Q: Optional[torch.Tensor] = None,
                     G: Optional[torch.Tensor] = None, scores: Optional[torch.Tensor] = None, g_base: int = 0) -> torch.Tensor:
    """Tie report (tie class T2): for every query the number of NON-ground-truth columns whose score equals
    the ground-truth threshold bit for bit — the only place where `gt in torch.topk(row, k)` (main.py:333) depends on
    torch's unspecified order of tied scores. Uses the ranking kernels only: one more COUNT pass with the bar just
    below the threshold gives #{S >= thr}; one THRESHOLD pass per ground-truth slot gives every GT column's own score
    (same MMA path, bit-identical), hence how many GT columns sit exactly at the threshold.
    Pass (Q, G) for the dense path (K1) or `scores` for a materialised matrix (KR)."""
    ge = (dense_rank(Q, G, gt_off, gt_idx, want_bits=False, g_base=g_base, thr=_just_below(r.thr)) if scores is None
          else rank_from_scores(scores, gt_off, gt_idx, want_bits=False, g_base=g_base, thr=_just_below(r.thr)))
    n_gt_at_thr = torch.zeros_like(r.cnt_above)
    for c, sizes, off_c, idx_c in _gt_slices(gt_off, gt_idx):
        one = (dense_rank(Q, G, off_c, idx_c, g_base=g_base, threshold_only=True) if scores is None
               else rank_from_scores(scores, off_c, idx_c, g_base=g_base, threshold_only=True))
        n_gt_at_thr += ((one.thr == r.thr) & (sizes > c)).to(torch.int32)
    return ge.cnt_above - r.cnt_above - n_gt_at_thr


def tie_summary(r: ProtocolRanks) -> dict:
    """Queries whose Recall@k could flip with the order of exactly tied scores: cnt_above < k <= cnt_above + ties."""
    if r.ties is None:
        raise ValueError("no tie counts (score_and_rank(..., want_ties=True))")
    out = {"queries_with_exact_ties": int((r.ties > 0).sum().item())}
    for k in (1, 5, 10):
        out[f"R@{k}_depends_on_tie_order"] = int(((r.cnt_above < k) & (r.cnt_above + r.ties >= k)).sum().item())
    return out


def topk_from_scores(scores: torch.Tensor, k: int = 10, g_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """KT. fp32 CUDA score matrix [nq, ng] -> (top_s fp32 [nq, k], top_i int32 [nq, k] global column ids):
    ``torch.topk(scores[i], k)`` of main.py:333 for every row, ties ordered by column."""
    _need_cuda(scores)
    if scores.dtype != torch.float32:
        raise TypeError("topk_from_scores takes fp32 scores")
    if scores.dim() != 2 or (scores.shape[1] > 1 and scores.stride(1) != 1):
        scores = scores.contiguous()
    nq, ng = scores.shape
    ld = scores.stride(0) if nq > 1 else max(ng, 1)
    top_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    top_i = torch.empty((nq, k), dtype=torch.int32, device=scores.device)
    _lib.check(_lib.lib().mmx_topk_from_scores(scores.data_ptr(), nq, ng, ld, g_base, k, top_s.data_ptr(),
                                               top_i.data_ptr(), _stream()))
    return top_s, top_i


def dense_topk(Q: torch.Tensor, G: torch.Tensor, k: int = 10, g_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1 top-k without the score matrix. Q [nq,d], G [ng,d] bf16 CUDA -> (top_s [nq,k], top_i [nq,k])."""
    _need_cuda(Q, G)
    if Q.dtype != torch.bfloat16 or G.dtype != torch.bfloat16:
        raise TypeError("dense top-k kernel takes bf16 embeddings")
    Q, G = Q.contiguous(), G.contiguous()
    nq, d = Q.shape
    ng = G.shape[0]
    h = _lib.lib()
    need = ctypes.c_size_t()
    _lib.check(h.mmx_dense_topk_workspace_bytes(nq, ng, ctypes.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=Q.device)
    top_s = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
    top_i = torch.empty((nq, k), dtype=torch.int32, device=Q.device)
    _lib.check(h.mmx_dense_topk(Q.data_ptr(), nq, G.data_ptr(), ng, d, g_base, k, top_s.data_ptr(), top_i.data_ptr(),
                                ws.data_ptr(), need.value, _stream()))
    return top_s, top_i


def merge_topk(gathered_s: torch.Tensor, gathered_i: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-shard lists concatenated along columns ([nq, n_shards * k], global ids) -> global top-k."""
    _need_cuda(gathered_s, gathered_i)
    gathered_s, gathered_i = gathered_s.contiguous(), gathered_i.contiguous()
    nq, width = gathered_s.shape
    if width % k or gathered_i.shape != gathered_s.shape:
        raise ValueError("gathered lists must be [nq, n_shards * k]")
    top_s = torch.empty((nq, k), dtype=torch.float32, device=gathered_s.device)
    top_i = torch.empty((nq, k), dtype=torch.int32, device=gathered_s.device)
    _lib.check(_lib.lib().mmx_merge_topk(gathered_s.data_ptr(), gathered_i.data_ptr(), width // k, nq, k,
                                         top_s.data_ptr(), top_i.data_ptr(), _stream()))
    return top_s, top_i


def place_bit_columns(gathered: torch.Tensor, bounds: Sequence[Tuple[int, int]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Shard merge: gathered int32 CUDA [n_shards, nq, wmax] bit columns of every shard (what one
    all-gather delivers) + the shards' global [begin, end) column ranges -> (bits int32 [nq, ceil(n_cols / 32)],
    cnt_above int32 [nq] = popcount of every merged row). One kernel; a shard may start inside a 32-column word."""
    _need_cuda(gathered)
    gathered = gathered.contiguous()
    n_shards, nq, wmax = gathered.shape
    begin = np.array([b for b, _ in bounds] + [bounds[-1][1]], dtype=np.int64)
    n_cols = int(begin[-1])
    wpr = (n_cols + 31) // 32
    bits = torch.empty((nq, wpr), dtype=torch.int32, device=gathered.device)
    cnt = torch.empty((nq,), dtype=torch.int32, device=gathered.device)
    _lib.check(_lib.lib().mmx_place_bit_columns(gathered.data_ptr(), n_shards, nq, wmax, begin.ctypes.data,
                                                bits.data_ptr(), wpr, cnt.data_ptr(), _stream()))
    return bits, cnt


def bincount(sample_idx: torch.Tensor, n_units: int) -> torch.Tensor:
    """int32 CUDA [B, n_draws] -> uint8 CUDA [B, n_units] multiplicities (saturated at 255)."""
    _need_cuda(sample_idx)
    if sample_idx.dtype != torch.int32:
        raise TypeError("sample indices must be int32 on the device")
    sample_idx = sample_idx.contiguous()
    B, n_draws = sample_idx.shape
    scratch = torch.empty((B, n_units), dtype=torch.int32, device=sample_idx.device)
    mult = torch.empty((B, n_units), dtype=torch.uint8, device=sample_idx.device)
    _lib.check(_lib.lib().mmx_bincount(sample_idx.data_ptr(), B, n_draws, n_units, scratch.data_ptr(), mult.data_ptr(),
                                       _stream()))
    return mult


def boot_counts(r: ProtocolRanks, mult: torch.Tensor, list_len: int = LIST_LEN,
                q_hits1: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K3 for one protocol: mult uint8 CUDA [B, n_units] -> int32 CUDA [B, 3] correct counts for k = 1, 5, 10."""
    if r.bits is None and r.list_i is not None:
        raise ValueError("LIST-mode ranks: use boot_counts_lists (it needs the exact-recount callback)")
    _need_cuda(mult, r.bits)
    if r.bits is None:
        raise ValueError("bootstrap needs the above-GT bitmask (want_bits=True)")
    B, n_units = mult.shape
    nq, wpr = r.bits.shape
    h = _lib.lib()
    list_t = torch.empty((list_len, nq), dtype=torch.int32, device=mult.device)
    _lib.check(h.mmx_above_lists(r.bits.data_ptr(), nq, wpr, _ptr(r.col2unit), list_len, list_t.data_ptr(), _stream()))
    correct = torch.empty((B, 3), dtype=torch.int32, device=mult.device)
    _lib.check(h.mmx_boot_counts(list_t.data_ptr(), list_len, r.cnt_above.data_ptr(), r.bits.data_ptr(), wpr,
                                 _ptr(r.col2unit), _ptr(r.q_unit), mult.data_ptr(), B, n_units, nq, correct.data_ptr(),
                                 _ptr(q_hits1), _stream()))
    return correct


def query_copies_drawn(mult: torch.Tensor, q_unit: Optional[torch.Tensor], nq: Optional[int] = None) -> torch.Tensor:
    """int64 CUDA [B]: how many query copies each bootstrap iteration drew = sum_q mult[b, q_unit[q]] — the denominator of
    main.py:358's recall (len(valid_indices)). q_unit None = identity over nq queries."""
    _need_cuda(mult)
    B, n_units = mult.shape
    n = int(q_unit.shape[0]) if q_unit is not None else int(nq)
    out = torch.empty((B,), dtype=torch.int64, device=mult.device)
    _lib.check(_lib.lib().mmx_copies_drawn(mult.data_ptr(), B, n_units, _ptr(q_unit), n, out.data_ptr(), _stream()))
    return out


def boot_counts_lists_raw(r: ProtocolRanks, mult: torch.Tensor, q_hits1: Optional[torch.Tensor] = None,
                          list_len: Optional[int] = None):
    """K3 in LIST mode without the exact recount: returns (correct int32 [B, 3] WITHOUT the unsettled pairs' share,
    pairs int32 [n, 2] (iteration, query) the truncated lists could not settle — None when there are none)."""
    _need_cuda(mult, r.list_i, r.list_n, r.cnt_above)
    n_shards, nq, cap = r.list_i.shape
    L = cap if list_len is None else min(list_len, cap)
    B, n_units = mult.shape
    h = _lib.lib()
    dev = mult.device
    need = ctypes.c_size_t()
    _lib.check(h.mmx_boot_counts_lists_workspace_bytes(B, n_units, ctypes.byref(need)))
    ws = torch.empty(max(need.value, 16), dtype=torch.uint8, device=dev)      # the transposed multiplicities
    correct = torch.empty((B, 3), dtype=torch.int32, device=dev)
    unres_cap = UNRESOLVED_CAP
    while True:
        unres = torch.empty((unres_cap, 2), dtype=torch.int32, device=dev)
        unres_n = torch.empty((1,), dtype=torch.int32, device=dev)
        hits = None if q_hits1 is None else q_hits1.clone()
        _lib.check(h.mmx_boot_counts_lists(r.list_i.data_ptr(), r.list_n.data_ptr(), n_shards, nq, cap, _ptr(r.col2unit), L,
                                           r.cnt_above.data_ptr(), _ptr(r.q_unit), mult.data_ptr(), B, n_units,
                                           correct.data_ptr(), _ptr(hits), unres.data_ptr(), unres_n.data_ptr(),
                                           unres_cap, ws.data_ptr(), ws.numel(), _stream()))
        n_un = int(unres_n.item())
        if n_un <= unres_cap:
            break
        unres_cap = n_un                                  # more unsettled pairs than room: run again with enough
    if q_hits1 is not None:
        q_hits1.copy_(hits)
    return correct, (unres[:n_un] if n_un else None)


def unlisted_request(pairs: torch.Tensor, mult: torch.Tensor):
    """What an exact recount needs for the unsettled pairs: (pairs with the iteration replaced by a row index into
    mult_rows, mult_rows = only those iterations' multiplicity rows)."""
    b_used, b_pos = torch.unique(pairs[:, 0], return_inverse=True)
    return torch.stack([b_pos.to(torch.int32), pairs[:, 1]], dim=1), mult[b_used.long()].contiguous()


def apply_unlisted(correct: torch.Tensor, pairs: torch.Tensor, c: torch.Tensor, r: ProtocolRanks, mult: torch.Tensor,
                   q_hits1: Optional[torch.Tensor] = None) -> None:
    """Add the exactly recounted pairs' share to `correct` in place: min(m_q, max(0, k - c)) per pair."""
    B, dev = correct.shape[0], correct.device
    b, q = pairs[:, 0].long(), pairs[:, 1].long()
    mq = mult[b, q if r.q_unit is None else r.q_unit[q].long()].long()
    for col, k in enumerate((1, 5, 10)):
        add = torch.minimum(mq, (k - c).clamp_min(0))
        correct[:, col] += torch.zeros(B, dtype=torch.int64, device=dev).index_add_(0, b, add).to(torch.int32)
        if k == 1 and q_hits1 is not None:
            q_hits1 += torch.zeros(r.nq, dtype=torch.int64, device=dev).index_add_(0, q, add).to(q_hits1.dtype)


def boot_counts_lists(r: ProtocolRanks, mult: torch.Tensor, resolve, q_hits1: Optional[torch.Tensor] = None,
                      list_len: Optional[int] = None, stats: Optional[dict] = None) -> torch.Tensor:
    """K3 in LIST mode: `r` from dense_rank_lists (per-shard lists stacked along dim 0 after the all-gather, cnt_above
    = the shard-summed exact counts). Pairs the truncated lists cannot settle are recounted exactly through
    `resolve(pairs[n, 2], mult_rows) -> c[n]` (dense_unlisted_resolver; sharded.unlisted_round across gallery shards)
    and added back: the result equals the bitmask path bit for bit."""
    correct, pairs = boot_counts_lists_raw(r, mult, q_hits1, list_len)
    if stats is not None:
        stats["unresolved_pairs"] = 0 if pairs is None else int(pairs.shape[0])
    if pairs is not None:
        c = resolve(*unlisted_request(pairs, mult))
        apply_unlisted(correct, pairs, c, r, mult, q_hits1)
    return correct


def ci_means(data: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """K4. data fp64 CUDA [n], idx int32 CUDA [R, n] -> fp64 CUDA [R] = np.mean(data[idx[r]]) bit for bit."""
    _need_cuda(data, idx)
    if data.dtype != torch.float64 or idx.dtype != torch.int32:
        raise TypeError("ci_means takes fp64 data and int32 indices")
    data, idx = data.contiguous(), idx.contiguous()
    R, n = idx.shape
    if n != data.numel():
        raise ValueError("idx rows must have len(data) entries")
    out = torch.empty((R,), dtype=torch.float64, device=data.device)
    _lib.check(_lib.lib().mmx_ci_means(data.data_ptr(), n, idx.data_ptr(), R, out.data_ptr(), _stream()))
    return out


def ci_means_batched(data: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """K4 for several data vectors at once. data fp64 CUDA [M, n], idx int32 CUDA [M, R, n] -> fp64 CUDA [M, R]."""
    _need_cuda(data, idx)
    if data.dtype != torch.float64 or idx.dtype != torch.int32:
        raise TypeError("ci_means_batched takes fp64 data and int32 indices")
    data, idx = data.contiguous(), idx.contiguous()
    M, R, n = idx.shape
    if tuple(data.shape) != (M, n):
        raise ValueError("data must be [M, n] for idx [M, R, n]")
    out = torch.empty((M, R), dtype=torch.float64, device=data.device)
    _lib.check(_lib.lib().mmx_ci_means_batched(data.data_ptr(), n, idx.data_ptr(), R, M, out.data_ptr(), _stream()))
    return out


def perm_stats(combined: torch.Tensor, n1: int, perm: torch.Tensor) -> torch.Tensor:
    """K5. combined fp64 CUDA [n1+n2], perm int32 CUDA [P, n1+n2] -> fp64 CUDA [P] |mean(first n1) - mean(rest)|."""
    _need_cuda(combined, perm)
    if combined.dtype != torch.float64 or perm.dtype != torch.int32:
        raise TypeError("perm_stats takes fp64 data and int32 permutations")
    combined, perm = combined.contiguous(), perm.contiguous()
    P, n = perm.shape
    if n != combined.numel():
        raise ValueError("perm rows must have len(combined) entries")
    out = torch.empty((P,), dtype=torch.float64, device=combined.device)
    _lib.check(_lib.lib().mmx_perm_stats(combined.data_ptr(), n1, n - n1, perm.data_ptr(), P, out.data_ptr(), _stream()))
    return out


class DeviceRng:
    """numpy's legacy MT19937 stream continued on the device (K6).

    ``DeviceRng(np.random)`` / ``DeviceRng(RandomState(42))`` snapshots the generator state onto the GPU; ``randint``
    then yields exactly the values the numpy object would have produced next (the ``choice(n, n)`` /
    ``choice(data, n)`` draws of main.py:588,269) as a CUDA int32 tensor; ``commit()`` writes the advanced state back
    into the numpy object, so the host stream continues as if numpy itself had made the draws."""

    PAR_MIN_COUNT = 200_000      # below this one small CTA (K6) is as fast as the jump-ahead launch sequence (K6p)
    _tables: dict = {}           # (device, long) -> (polys, scratch, bit positions, bit counts, seg_blocks)

    def __init__(self, source, device: Optional[torch.device] = None):
        self.source = np.random if source is None else source
        st = self.source.get_state()
        if st[0] != "MT19937":
            raise ValueError("DeviceRng needs numpy's legacy MT19937 RandomState")
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        host = np.concatenate([np.asarray(st[1], dtype=np.uint32), np.array([st[2]], dtype=np.uint32)])
        self.state = torch.from_numpy(host.view(np.int32).copy()).to(dev)

    LONG_MIN_WORDS = 24_000_000  # raw words from which the 512-block table pays (two launches of the 64-block one)

    def _par_buffers(self, long: bool = False):
        """(polys, scratch, positions, counts, seg_blocks) of the jump table for this device: the 64-block table or
        (long) the 512-block one, uploaded and registered with the library on first use."""
        dev = self.state.device
        key = (dev, long)
        if key not in DeviceRng._tables:
            from . import mtjump
            table = mtjump.load_table(build_if_missing=False, long=long)
            polys = torch.from_numpy(table.view(np.int32)).to(dev)
            need = ctypes.c_size_t()
            _lib.check(_lib.lib().mmx_mt19937_par_scratch_bytes(mtjump.N_POLY, ctypes.byref(need)))
            # set-bit positions of every polynomial (the jump convolution gathers x[n + i] over them)
            pos, cnt = mtjump.bit_positions(table)
            pos_d = torch.from_numpy(pos.view(np.int16)).to(dev)
            cnt_d = torch.from_numpy(cnt).to(dev)
            _lib.check(_lib.lib().mmx_mt19937_par_set_bit_positions(polys.data_ptr(), pos_d.data_ptr(),
                                                                    cnt_d.data_ptr(), pos.shape[1]))
            DeviceRng._tables[key] = (polys, need.value, pos_d, cnt_d,
                                      mtjump.SEG_BLOCKS_LONG if long else mtjump.SEG_BLOCKS)
        polys, need_bytes, pos_d, cnt_d, seg_blocks = DeviceRng._tables[key]
        # the jump tables are shared per device; the scratch is per generator: two generators may draw at the same time
        # on different streams (ShufflePrefetch next to a model's bootstrap draws)
        if not hasattr(self, "_scratch"):
            self._scratch = {}
        if long not in self._scratch:
            self._scratch[long] = torch.empty(need_bytes, dtype=torch.uint8, device=dev)
        return polys, self._scratch[long], pos_d, cnt_d, seg_blocks

    def _table_for(self, raw_words: float):
        return self._par_buffers(long=raw_words >= self.LONG_MIN_WORDS)

    def randint(self, high: int, count: int, parallel: Optional[bool] = None,
                window: Optional[Tuple[int, int]] = None) -> Optional[torch.Tensor]:
        """The next `count` values of ``randint(0, high)``. Large requests use the jump-ahead kernels (K6p, all SMs),
        small ones the single-CTA kernel (K6); both continue the same stream, so they can be mixed freely.
        window=(begin, n): the generator advances by all `count` values but only values [begin, begin + n) are kept
        and returned (None when n == 0) — a rank's block of a stream every rank replays, or a plain skip."""
        dev = self.state.device
        if window is None:
            window = (0, count)
        w0, wn = int(window[0]), int(window[1])
        if w0 < 0 or wn < 0 or w0 + wn > count:
            raise ValueError("window outside the request")
        if parallel is None:
            parallel = count >= self.PAR_MIN_COUNT
        if parallel:
            from . import mtjump
            out = torch.empty((wn,), dtype=torch.int32, device=dev) if wn else None
            mask = (1 << max(int(high) - 1, 1).bit_length()) - 1                   # numpy's masked rejection
            polys, scratch, _, _, seg_blocks = self._table_for(count * (mask + 1.0) / float(high))
            _lib.check(_lib.lib().mmx_mt19937_randint_par_window(
                self.state.data_ptr(), int(high), int(count), _ptr(out), w0, wn, polys.data_ptr(), mtjump.N_POLY,
                seg_blocks, scratch.data_ptr(), scratch.numel(), _stream()))
            return out
        out = torch.empty((count,), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().mmx_mt19937_randint(self.state.data_ptr(), int(high), int(count), out.data_ptr(), _stream()))
        return out[w0:w0 + wn] if wn else None

    def raw_words(self, count: int, advance: bool = True) -> Optional[torch.Tensor]:
        """The next `count` raw 32-bit outputs (as int32 bit patterns). advance=False leaves the generator where it
        is (the draws are made from a copy of the state); ``skip`` then moves it by the number actually consumed."""
        from . import mtjump
        polys, scratch, _, _, seg_blocks = self._table_for(count)
        st = self.state if advance else self.state.clone()
        out = torch.empty((count,), dtype=torch.int32, device=self.state.device)
        _lib.check(_lib.lib().mmx_mt19937_raw_par(st.data_ptr(), int(count), out.data_ptr(), polys.data_ptr(),
                                                  mtjump.N_POLY, seg_blocks, scratch.data_ptr(),
                                                  scratch.numel(), _stream()))
        return out

    def skip(self, count: int) -> None:
        """Advance the generator by `count` raw words without materialising them."""
        from . import mtjump
        polys, scratch, _, _, seg_blocks = self._table_for(count)
        _lib.check(_lib.lib().mmx_mt19937_raw_par(self.state.data_ptr(), int(count), None, polys.data_ptr(),
                                                  mtjump.N_POLY, seg_blocks, scratch.data_ptr(),
                                                  scratch.numel(), _stream()))

    @staticmethod
    def shuffle_words_needed(n: int, n_shuffles: int) -> int:
        """Raw words handed to the replay of n_shuffles shuffles of n items (mean + 8 sigma): step i accepts with
        probability (i + 1) / 2^ceil(log2(i + 1))."""
        i = np.arange(1, n, dtype=np.float64)
        pacc = (i + 1.0) / np.exp2(np.ceil(np.log2(i + 1.0)))
        mean = n_shuffles * float(np.sum(1.0 / pacc))
        sigma = float(np.sqrt(n_shuffles * np.sum((1.0 - pacc) / pacc ** 2)))
        return int(mean + 8.0 * sigma) + 2048

    def shuffle_indices(self, n: int, n_shuffles: int) -> torch.Tensor:
        """K6s. int32 CUDA [n_shuffles, n]: an index array after each of `n_shuffles` cumulative in-place
        ``shuffle`` calls (legacy RandomState.shuffle, main.py:296), consuming exactly the words numpy would."""
        h = _lib.lib()
        dev = self.state.device
        n_words = self.shuffle_words_needed(n, n_shuffles)
        perm = torch.empty((n_shuffles, n), dtype=torch.int32, device=dev)
        info = torch.empty((2,), dtype=torch.int64, device=dev)
        while True:
            words = self.raw_words(n_words, advance=False)
            need = ctypes.c_size_t()
            _lib.check(h.mmx_perm_shuffle_workspace_bytes(n, n_shuffles, n_words, ctypes.byref(need)))
            ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
            _lib.check(h.mmx_perm_shuffle(words.data_ptr(), n_words, n, n_shuffles, perm.data_ptr(), info.data_ptr(),
                                          ws.data_ptr(), need.value, _stream()))
            # wait on the stream first: a device->host copy into pageable memory that has to wait for its stream does so
            # inside the driver, where it holds up the launches of every other host thread (ShufflePrefetch runs this
            # on a worker thread next to the scoring thread)
            torch.cuda.current_stream().synchronize()
            consumed = int(info[1].item())
            if consumed >= 0:
                break
            if consumed == -3:
                raise _lib.MmxError("shuffle replay: a mask epoch did not fit the 65 536-word window")
            n_words = int(n_words * 1.25)          # 8 sigma was not enough: replay with a longer stream
        self.skip(consumed)
        return perm

    def commit(self) -> None:
        host = self.state.cpu().numpy().view(np.uint32)
        st = self.source.get_state()
        self.source.set_state(("MT19937", host[:624].copy(), int(host[624]), st[3], st[4]))


class ShufflePrefetch:
    """The index rows of a KNOWN sequence of permutation tests, replayed ahead of time on a side stream.

    numpy's stream does not depend on the data: which words a test's 10 000 shuffles consume is fixed by the generator
    state alone. So the whole chain of tests — serial, ~0.07 ms per 50 000-element permutation on ONE SM — can run while
    the GPU (or the other GPUs) still score the models whose vectors the tests will compare (BASELINE configs[4]: 21
    per-query tests = 14.5 s of chain next to 8-10 s of MaxSim scoring). ``skip(drng)`` first advances the snapshot
    past the draws that happen before the tests (the models' bootstrap draws, ``api.skip_bootstrap_draws``).

        pre = ShufflePrefetch(rng, [(50000, 10000)] * 21, skip=...); pre.start()
        ...                                                   # score the models (they advance `rng` themselves)
        assert pre.starts_where(rng)                          # the dry run consumed exactly what the models did
        p = api.permutation_test(a, b, index_rows=pre.index_rows(k))
        pre.commit(rng)                                       # `rng` continues behind the last prefetched test
    """

    def __init__(self, source, tests: Sequence[Tuple[int, int]], device: Optional[torch.device] = None, skip=None):
        import threading
        self.tests = [(int(n), int(P)) for n, P in tests]
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.drng = DeviceRng(source, self.dev)
        self.drng._par_buffers(False)                      # jump tables are uploaded once, not from the worker thread
        self.drng._par_buffers(True)
        if skip is not None:
            skip(self.drng)
        torch.cuda.synchronize(self.dev)
        self.start_state = self.drng.state.clone()
        self.stream = torch.cuda.Stream(device=self.dev)
        # cudaMalloc synchronises the whole device: an allocation made by the worker while the main stream sits in a
        # collective (waiting for the other ranks' scoring) would stall the chain until that collective ends. So every
        # buffer the chain will need is allocated once now, on its stream, and handed back to the caching allocator.
        with torch.cuda.stream(self.stream):
            warm = [torch.empty((P, n), dtype=torch.int32, device=self.dev) for n, P in self.tests]
            warm.append(torch.empty(1024, dtype=torch.int64, device=self.dev))     # the small-block pool of this stream
            for n, P in sorted(set(self.tests)):
                n_words = self.drng.shuffle_words_needed(n, P)
                need = ctypes.c_size_t()
                _lib.check(_lib.lib().mmx_perm_shuffle_workspace_bytes(n, P, n_words, ctypes.byref(need)))
                warm.append(torch.empty(n_words, dtype=torch.int32, device=self.dev))
                warm.append(torch.empty(need.value, dtype=torch.uint8, device=self.dev))
            del warm
        torch.cuda.synchronize(self.dev)
        self.out = [None] * len(self.tests)
        self.state_after = [None] * len(self.tests)         # generator state behind every test (625 words each)
        self.ready = [threading.Event() for _ in self.tests]
        self.error: Optional[BaseException] = None
        self.trace: List[Tuple[int, float, float]] = []      # (test, host time.perf_counter() at start, at end)
        self.thread = threading.Thread(target=self._run, daemon=True)

    def start(self) -> "ShufflePrefetch":
        self.thread.start()
        return self

    def _run(self) -> None:
        try:
            torch.cuda.set_device(self.dev)
            with torch.cuda.stream(self.stream):
                for k, (n, P) in enumerate(self.tests):
                    t0 = time.perf_counter()
                    rows = self.drng.shuffle_indices(n, P)
                    self.trace.append((k, t0, time.perf_counter()))
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                    self.state_after[k] = self.drng.state.clone()
                    self.out[k] = (rows, ev)
                    self.ready[k].set()
        except BaseException as exc:  # noqa: BLE001 - reported by the consumer
            self.error = exc
            for r in self.ready:
                r.set()

    @staticmethod
    def _same_state(source, state: torch.Tensor) -> bool:
        st = source.get_state()
        host = state.cpu().numpy().view(np.uint32)
        return bool(np.array_equal(np.asarray(st[1], dtype=np.uint32), host[:624]) and int(st[2]) == int(host[624]))

    def starts_where(self, source) -> bool:
        """True when `source` (numpy generator) stands exactly where the prefetched chain started."""
        return self._same_state(source, self.start_state)

    def stands_after(self, source, k: int) -> bool:
        """True when `source` stands exactly where the chain stood after test k (for checks against numpy's own shuffles)."""
        self.ready[k].wait()
        self.stream.synchronize()
        return self._same_state(source, self.state_after[k])

    def index_rows(self, k: int) -> torch.Tensor:
        """int32 CUDA [P, n] index rows of test k (blocks until its replay has been queued; the current stream waits
        for it on the device)."""
        self.ready[k].wait()
        if self.error is not None:
            raise self.error
        rows, ev = self.out[k]
        torch.cuda.current_stream().wait_event(ev)
        rows.record_stream(torch.cuda.current_stream())
        self.out[k] = None                                 # the caller owns them now
        return rows

    def commit(self, source) -> None:
        """Move the numpy generator behind the last prefetched test."""
        self.thread.join()
        if self.error is not None:
            raise self.error
        self.stream.synchronize()
        self.drng.source = source
        self.drng.commit()

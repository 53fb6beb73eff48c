"""Synthetic workloads of BASELINE.json's larger configurations and the drivers that run them through the
public API — shared by bench.py (the ``cfg3`` / ``cfg5`` sections of every JSON line) and dev/run_cfg5.py.

configs[2]  1024-d dense scoring scaled to a 10^6-image gallery, gallery-sharded (T2I with the 1000-iteration bootstrap in
            LIST mode; I2T = 10^6 images against 5 x 10^6 captions, main.py:622).
configs[4]  ColPali MaxSim against a 100 000-document corpus, doc-sharded, 1000-iteration bootstrap, pairwise
            permutation tests on per-iteration arrays and on per-query vectors (main.py:281-302, :790).

The generators are shard-invariant: a rank builds exactly its slice of the same global tensors whatever the world size
(documents come in blocks of 500 with one seed per block), so the integers can be compared across N = 1, 2, 4, 8.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import api, ops, sharded

DOC_BLOCK = 500


def _unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=-1, keepdim=True)


def _world(group=None) -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _max_over_ranks(x: float, dev: torch.device, group=None) -> float:
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


# ------------------------------------------------------------------------------------------------ configs[4]
def synth_colpali_shard(dev: torch.device, nq: int, nd: int, lq: int, ld: int, d0: int, d1: int, seed: int = 2,
                        sigma: float = 1.0, dim: int = 128):
    """Queries (replicated) and documents [d0, d1) of a corpus of nd documents x ld patches. Query i belongs to document
    q2d[i] = i * nd // nq; its lq tokens are planted into that document as noisy copies (patch rows drawn per token), so
    MaxSim has something to find and Recall@1 is neither 0 nor 100 %. Returns (Q bf16 [nq*lq, dim], D bf16 [(d1-d0)*ld, dim],
    q2d int32 [nq]) on `dev`."""
    g = torch.Generator(device=dev).manual_seed(seed)
    Q = _unit(torch.randn(nq * lq, dim, generator=g, device=dev))
    rows = torch.randint(0, ld, (nq, lq), generator=g, device=dev)
    noise = _unit(torch.randn(nq * lq, dim, generator=g, device=dev))
    planted = _unit(Q + sigma * noise).bfloat16().view(nq, lq, dim)
    q2d = (torch.arange(nq, device=dev, dtype=torch.int64) * nd) // nq
    D = torch.empty(((d1 - d0) * ld, dim), dtype=torch.bfloat16, device=dev)
    for blk in range(d0 // DOC_BLOCK, (d1 + DOC_BLOCK - 1) // DOC_BLOCK):
        a, b = blk * DOC_BLOCK, min(nd, (blk + 1) * DOC_BLOCK)
        gb = torch.Generator(device=dev).manual_seed(1_000_003 * (seed + 1) + blk)
        x = _unit(torch.randn((b - a) * ld, dim, generator=gb, device=dev)).bfloat16()
        lo, hi = max(a, d0), min(b, d1)
        D[(lo - d0) * ld:(hi - d0) * ld] = x[(lo - a) * ld:(hi - a) * ld]
        del x
    mine = (q2d >= d0) & (q2d < d1)
    qi = torch.nonzero(mine).squeeze(1)
    if qi.numel():
        dst = (((q2d[qi] - d0) * ld)[:, None] + rows[qi]).reshape(-1)     # flat patch row of every planted token
        src = planted[qi].reshape(-1, dim)
        # two tokens may land on the same patch row: the LAST one (query order, then token order) wins — decided by a sort,
        # not by the scatter's write order, so the corpus is the same at every shard count and in every run
        order = torch.arange(dst.numel(), device=dev, dtype=torch.int64)
        perm = torch.sort(dst * dst.numel() + order).indices
        ds = dst[perm]
        last = torch.ones_like(ds, dtype=torch.bool)
        last[:-1] = ds[1:] != ds[:-1]
        D[ds[last]] = src[perm[last]]
    return Q.bfloat16().contiguous(), D, q2d.to(torch.int32)


def colpali_model(dev: torch.device, nq: int, nd: int, lq: int, ld: int, n_iterations: int, seed: int, sigma: float,
                  rng, group=None, name: str = "model", bounds=None) -> Tuple[Optional[dict], np.ndarray, Dict[str, float]]:
    """One ColPali 'model' of configs[4], doc-sharded over `group`: synthesize this rank's shard on the device, score +
    rank + bootstrap through ``api.run_retrieval_bootstrap``. Returns (row on rank 0 / None elsewhere, per-query Recall@1
    hit vector float64 [nq] (the same on every rank), phase times in ms)."""
    world, rank = _world(group)
    if bounds is None:
        bounds = sharded.shard_bounds(nd, world)       # `bounds` may leave a rank without documents (it then only helps with the statistics)
    d0, d1 = bounds[rank]
    t = {}
    torch.cuda.current_stream(dev).synchronize()   # not the device: a shuffle chain may be running on a side stream
    t0 = time.perf_counter()
    Q, D, q2d = synth_colpali_shard(dev, nq, nd, lq, ld, d0, d1, seed=seed, sigma=sigma)
    torch.cuda.current_stream(dev).synchronize()   # not the device: a shuffle chain may be running on a side stream
    t["synthesize_ms"] = 1e3 * (time.perf_counter() - t0)
    qv = api.FlatMultiVector(Q, np.full(nq, lq))
    dv = api.FlatMultiVector(D, np.full(d1 - d0, ld))
    keep: dict = {}
    tm: dict = {}
    t0 = time.perf_counter()
    row = api.run_retrieval_bootstrap(qv, dv, q2d.cpu().numpy(), {"type": "colpali", "name": name}, n_iterations, rng=rng,
                                      group=group, timings=tm, shard_sizes=[e - b for b, e in bounds], keep=keep)
    torch.cuda.current_stream(dev).synchronize()   # not the device: a shuffle chain may be running on a side stream
    t["score_rank_bootstrap_ms"] = 1e3 * (time.perf_counter() - t0)
    hits = (keep["ranks"].cnt_above < 1).to(torch.float64).cpu().numpy()     # per-query correct_r1 (main.py:342-356)
    if tm:
        t["maxsim_kernel_ms"] = tm.get("score_ms")
    t["peak_mem_GB"] = torch.cuda.max_memory_allocated(dev) / 1e9
    del Q, D, qv, dv, keep
    return row, hits, t


def cfg5_section(dev: torch.device, group=None, n_iterations: int = 1000, nq: int = 25000, nd: int = 100000,
                 lq: int = 20, ld: int = 1030, n_perm: int = 10000) -> Optional[dict]:
    """configs[4] for bench.py: ONE model scored + bootstrapped (doc-sharded), then on rank 0 one permutation test on
    per-iteration arrays (n = 2 x 1000) and one on per-query vectors (n = 2 x 25 000) — against a copy of the model's
    own vectors with 2 % of the entries perturbed (the cost of a test does not depend on the data). dev/run_cfg5.py runs
    the full seven-model, 42-test configuration."""
    world, rank = _world(group)
    row, hits, t = colpali_model(dev, nq, nd, lq, ld, n_iterations, seed=11, sigma=3.8,
                                 rng=np.random.RandomState(42), group=group, name="cfg5-model")
    total_ms = _max_over_ranks(t["score_rank_bootstrap_ms"], dev, group)
    k2_ms = _max_over_ranks(t.get("maxsim_kernel_ms") or 0.0, dev, group)
    peak = _max_over_ranks(t["peak_mem_GB"], dev, group)
    if rank != 0:
        return None
    flop = 2.0 * lq * ld * 128 * nq * nd
    rs = np.random.RandomState(7)
    v1 = row["_bootstrap_values"][:, 0]
    v2 = v1 + rs.choice([0.0, 0.004, -0.004], size=v1.shape, p=[0.98, 0.01, 0.01])
    h2 = np.where(rs.rand(nq) < 0.02, 1.0 - hits, hits)
    out = {"workload": f"ColPali MaxSim {nq} queries x {nd} docs x {ld} patches, doc-sharded x{world}, {n_iterations}-iter bootstrap",
           "score_rank_bootstrap_ms": total_ms, "maxsim_kernel_ms": k2_ms,
           "maxsim_tflops_per_gpu": flop / world / (k2_ms * 1e-3) / 1e12 if k2_ms else None,
           "pairs_per_s": nq * nd / (total_ms * 1e-3), "peak_mem_GB_per_gpu": peak,
           "base_R@1/5/10": [row["base"][k] for k in (1, 5, 10)],
           "R@1_ci": [row["R@1_mean"], row["R@1_lower"], row["R@1_upper"]]}
    for key, (a, b) in (("permutation_test_per_iteration_n2000", (v1, v2)), ("permutation_test_per_query_n50000", (hits, h2))):
        g = np.random.RandomState(99)
        api.permutation_test(a[:64], b[:64], 16, rng=np.random.RandomState(1))          # lazy initialisation outside the timing
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        p = api.permutation_test(a, b, n_perm, rng=g)
        torch.cuda.synchronize(dev)
        out[key] = {"ms": 1e3 * (time.perf_counter() - t0), "p": float(p), "n_permutations": n_perm}
    return out


# ------------------------------------------------------------------------------------------------ configs[2]
def synth_dense_gallery(dev: torch.device, n: int, d: int, a: int, b: int, sigma: float = 9.0, seed: int = 0, chunk: int = 125_000):
    """Images [a, b) of a gallery of n x d unit vectors and ALL n caption-0 queries (img + sigma * unit noise), generated in
    chunks with one seed per chunk so that every world size sees the same global tensors. Returns (txt bf16 [n, d], G bf16 [b-a, d])."""
    txt = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    G = torch.empty((b - a, d), dtype=torch.bfloat16, device=dev)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        g = torch.Generator(device=dev).manual_seed(seed * 7919 + c0 // chunk + 1)
        img = _unit(torch.randn(c1 - c0, d, generator=g, device=dev))
        txt[c0:c1] = _unit(img + sigma * _unit(torch.randn(c1 - c0, d, generator=g, device=dev))).bfloat16()
        lo, hi = max(a, c0), min(b, c1)
        if hi > lo:
            G[lo - a:hi - a] = img[lo - c0:hi - c0].bfloat16()
        del img
    return txt, G


def synth_caption_shard(dev: torch.device, n_img: int, d: int, a: int, b: int, caps: int = 5, sigma: float = 9.0, seed: int = 0,
                        chunk: int = 125_000):
    """Captions [a, b) of the caps * n_img captions of configs[2]'s I2T protocol (caption 5 i + c belongs to image i) and
    ALL n_img images. Returns (img bf16 [n_img, d], C bf16 [b-a, d])."""
    img_all = torch.empty((n_img, d), dtype=torch.bfloat16, device=dev)
    C = torch.empty((b - a, d), dtype=torch.bfloat16, device=dev)
    for c0 in range(0, n_img, chunk):
        c1 = min(n_img, c0 + chunk)
        g = torch.Generator(device=dev).manual_seed(seed * 7919 + c0 // chunk + 1)
        img = _unit(torch.randn(c1 - c0, d, generator=g, device=dev))
        img_all[c0:c1] = img.bfloat16()
        lo, hi = max(a, c0 * caps), min(b, c1 * caps)
        if hi > lo:
            g2 = torch.Generator(device=dev).manual_seed(seed * 104729 + c0 // chunk + 1)
            for k0 in range(c0, c1, 25_000):                       # 25 000 images = 125 000 captions at a time
                k1 = min(c1, k0 + 25_000)
                noise = _unit(torch.randn((k1 - k0) * caps, d, generator=g2, device=dev))
                cap = _unit(img[k0 - c0:k1 - c0].repeat_interleave(caps, 0) + sigma * noise).bfloat16()
                l2, h2 = max(lo, k0 * caps), min(hi, k1 * caps)
                if h2 > l2:
                    C[l2 - a:h2 - a] = cap[l2 - k0 * caps:h2 - k0 * caps]
                del noise, cap
        del img
    return img_all, C


def cfg3_section(dev: torch.device, group=None, n: int = 1_000_000, d: int = 1024, n_iterations: int = 1000,
                 caps: int = 5, i2t: bool = True, peak_tflops: Optional[float] = None) -> Optional[dict]:
    """configs[2] for bench.py, gallery-sharded over `group`:
      t2i   n caption-0 queries x n images: threshold pass, COUNT + LIST pass, all-gather of the lists, then the
            n_iterations bootstrap (LIST mode: the [n, n] bitmask would be 125 GB) and three CIs — api.run_retrieval_bootstrap;
      i2t   n images x caps * n captions (main.py:622; GT = the image's caps captions): threshold all-reduce(MAX) + COUNT pass +
            counts all-reduce(SUM) -> Recall@K.
    Times are wall clock around device-synchronised calls, max over ranks; TFLOP/s count 2 n_q n_g d per scoring pass."""
    import torch.distributed as dist
    world, rank = _world(group)
    out: dict = {"workload": f"dense {n} x {n} x {d} gallery-sharded x{world} (T2I + {n_iterations}-iter LIST-mode bootstrap)"
                             + (f"; I2T {n} images x {caps * n} captions" if i2t else "")}
    bounds = sharded.shard_bounds(n, world)
    a, b = bounds[rank]
    txt, G = synth_dense_gallery(dev, n, d, a, b)
    # lazy initialisation (kernel attributes, jump tables, NCCL buffers) on a small problem, outside the timing
    m = min(n, 20_000)
    sb = sharded.shard_bounds(m, world)
    api.run_retrieval_bootstrap(txt[:m], txt[sb[rank][0]:sb[rank][1]], np.arange(m), {"type": "dense", "name": "warm"}, 50,
                                rng=np.random.RandomState(1), group=group, list_mode=True,
                                shard_sizes=[e - s for s, e in sb])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(group=group)
    # two calls at full size: the first one also pays the device allocations of every full-size buffer (cudaMalloc of ~10 GB
    # of lists, draws and multiplicities per rank: tens of ms, and it synchronises the device); the second is the steady state
    # a pipeline that scores model after model sees. Both are reported.
    call_ms = []
    for _ in range(2):
        tm: dict = {}
        t0 = time.perf_counter()
        row = api.run_retrieval_bootstrap(txt, G, np.arange(n), {"type": "dense", "name": "cfg3-t2i"}, n_iterations,
                                          rng=np.random.RandomState(42), group=group, timings=tm,
                                          shard_sizes=[e - s for s, e in bounds])
        torch.cuda.synchronize(dev)
        call_ms.append(_max_over_ranks(1e3 * (time.perf_counter() - t0), dev, group))
        if world > 1:
            dist.barrier(group=group)
    t2i_ms = call_ms[1]
    thr_ms = _max_over_ranks(tm.get("score_ms", 0.0), dev, group)
    # the COUNT + LIST pass alone (this rank's shard), for the per-GPU tensor-pipe figure
    gt_off = torch.arange(n + 1, dtype=torch.int32, device=dev)
    gt_idx = torch.arange(n, dtype=torch.int32, device=dev)
    thr = sharded.merge_thresholds(ops.dense_rank(txt, G, gt_off, gt_idx, want_bits=False, g_base=a, threshold_only=True).thr, group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.dense_rank_lists(txt, G, gt_off, gt_idx, g_base=a, thr=thr)
    e1.record()
    torch.cuda.synchronize(dev)
    k1_ms = _max_over_ranks(e0.elapsed_time(e1), dev, group)
    k1_tf = 2.0 * n * (b - a) * d / (k1_ms * 1e-3) / 1e12
    if rank == 0:
        out["t2i"] = {"score_rank_bootstrap_ms": t2i_ms, "first_call_ms (incl. device allocations)": call_ms[0],
                      "threshold_pass_ms": thr_ms, "count_list_pass_ms": k1_ms,
                      "rest_ms (all-gather of the lists, numpy stream, bincount, K3, K4, D2H)": t2i_ms - thr_ms - k1_ms,
                      "tflops_per_gpu_count_list_pass": k1_tf,
                      "frac_of_bf16_peak": k1_tf / peak_tflops if peak_tflops else None,
                      "pairs_per_s": n * n / (t2i_ms * 1e-3), "base_R@1/5/10": [row["base"][k] for k in (1, 5, 10)],
                      "R@1_ci": [row["R@1_mean"], row["R@1_lower"], row["R@1_upper"]], "list_mode": row.get("_list_mode"),
                      "device_ms_by_phase_rank0": tm.get("device_ms"), "host_checkpoints_ms_rank0": tm.get("wall_ms")}
    del txt, G, thr
    torch.cuda.empty_cache()
    if i2t:
        nc = caps * n
        cb = sharded.shard_bounds(nc, world)
        ca, cbe = cb[rank]
        img, C = synth_caption_shard(dev, n, d, ca, cbe, caps=caps)
        b_off = torch.arange(0, nc + 1, caps, dtype=torch.int32, device=dev)
        b_idx = torch.arange(nc, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=group)
        t0 = time.perf_counter()
        e0.record()
        loc = ops.dense_rank(img, C, b_off, b_idx, want_bits=False, g_base=ca, threshold_only=True)
        thr = sharded.merge_thresholds(loc.thr, group)
        loc = ops.dense_rank(img, C, b_off, b_idx, want_bits=False, g_base=ca, thr=thr)
        cnt = loc.cnt_above.clone()
        if world > 1:
            dist.all_reduce(cnt, group=group)
        e1.record()
        correct = [int((cnt < k).sum().item()) for k in (1, 5, 10)]
        torch.cuda.synchronize(dev)
        ms = _max_over_ranks(e0.elapsed_time(e1), dev, group)
        tf = 2.0 * n * (cbe - ca) * d / (ms * 1e-3) / 1e12
        if rank == 0:
            out["i2t"] = {"score_rank_ms": ms, "tflops_per_gpu": tf, "frac_of_bf16_peak": tf / peak_tflops if peak_tflops else None,
                          "pairs_per_s": n * nc / (ms * 1e-3), "R@1/5/10": [100.0 * c / n for c in correct]}
        del img, C
        torch.cuda.empty_cache()
    return out if rank == 0 else None

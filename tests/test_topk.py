"""KT: per-query top-k lists (`torch.topk(scores[i], k)` of main.py:333,362,372) from a score matrix, fused into the
dense GEMM without a matrix, merged across gallery shards. GPU tests call through the C ABI; the gloo test covers the
single all-gather of the sharded merge on CPU."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from helpers import from_bits, synth_dense, unit

ROOT = Path(__file__).resolve().parent.parent
REL_TOL = 1e-3


@pytest.fixture(scope="module")
def mmx_mod():
    import mmx
    assert torch.cuda.is_available(), "GPU tests need a B200"
    mmx._lib.check(mmx._lib.lib().mmx_device_check())
    return mmx


def _check_against_cpu_scores(top_s, top_i, s_cpu, k, g_base=0):
    """GPU lists vs CPU fp32 scores: values at the reported columns agree within REL_TOL, lists are ordered, columns
    unique, and indices equal the oracle's wherever the oracle's top-(k+1) values are separated by more than the
    tolerance (elsewhere GPU-MMA vs CPU summation order may swap neighbours: tie class T3)."""
    top_s, top_i = top_s.cpu(), top_i.cpu().long() - g_base
    n, g = s_cpu.shape
    kk = min(k, g)
    want_v, want_i = oracle.topk_lists(s_cpu, k + 1 if g > k else k)
    assert torch.all(top_i[:, kk:] == -1 - g_base) and torch.all(torch.isinf(top_s[:, kk:]))
    got_at = torch.gather(s_cpu, 1, top_i[:, :kk])
    rel = (got_at - top_s[:, :kk]).abs() / got_at.abs().clamp_min(1e-3)
    assert float(rel.max()) < REL_TOL
    assert torch.all(top_s[:, :kk - 1] >= top_s[:, 1:kk])
    assert all(len(set(r.tolist())) == kk for r in top_i[:, :kk])
    tol = REL_TOL * want_v[:, :kk].abs().clamp_min(1e-3)
    gap_next = (want_v[:, :kk] - want_v[:, 1:kk + 1]) if g > k else torch.full((n, kk), float("inf"))
    gap_next = torch.where(torch.isnan(gap_next), torch.full_like(gap_next, float("inf")), gap_next)
    gap_prev = torch.cat([torch.full((n, 1), float("inf")), want_v[:, :kk - 1] - want_v[:, 1:kk]], dim=1)
    clear = (gap_next > tol) & (gap_prev > tol)
    assert torch.equal(top_i[:, :kk][clear], want_i[:, :kk][clear])
    return float(clear.float().mean())


@pytest.mark.gpu
def test_topk_from_scores_is_torch_topk(mmx_mod, golden_dir):
    dev = torch.device("cuda:0")
    z = np.load(golden_dir / "dense.npz")
    for name in ("scores_t2i", "scores_i2t"):
        s = torch.from_numpy(z[name])
        ts, ti = mmx_mod.ops.topk_from_scores(s.to(dev), 10)
        tv = torch.topk(s, 10, dim=1)                                   # the reference's call (main.py:333)
        assert torch.equal(ts.cpu(), tv.values) and torch.equal(ti.cpu().long(), tv.indices)
    rs = np.random.RandomState(3)
    for nq, ng, k in ((1, 1, 10), (3, 7, 10), (5, 31, 5), (9, 32, 1), (17, 129, 32), (64, 1000, 10), (33, 4097, 10)):
        s = torch.from_numpy(rs.randn(nq, ng).astype(np.float32))
        s[:, ::5] = s[:, :1].clone()                                    # exact ties: ordered by column
        wide = torch.full((nq, ng + 13), 9.0)
        wide[:, :ng] = s                                                # row pitch > ng
        ts, ti = mmx_mod.ops.topk_from_scores(wide.to(dev)[:, :ng], k, g_base=1000)
        wv, wi = oracle.topk_lists(s, k)
        assert torch.equal(ts.cpu(), wv)
        assert torch.equal(ti.cpu().long(), torch.where(wi >= 0, wi + 1000, wi))


@pytest.mark.gpu
@pytest.mark.parametrize("nq,ng,d", [(300, 1500, 768), (257, 5000, 1152), (130, 40000, 256), (5000, 25000, 64)])
def test_dense_topk_without_matrix(mmx_mod, nq, ng, d):
    """Small gallery (matrix in the workspace), 32-column groups, 256-column groups (ng > 32768)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(nq)
    G = unit(torch.randn(ng, d, generator=g)).bfloat16()
    Q = unit(G[torch.randint(0, ng, (nq,), generator=g)].float() + 2.0 * unit(torch.randn(nq, d, generator=g))).bfloat16()
    ts, ti = mmx_mod.ops.dense_topk(Q.to(dev), G.to(dev), 10, g_base=64)
    clear = _check_against_cpu_scores(ts, ti, oracle.dense_scores(G, Q), 10, g_base=64)
    assert clear > 0.5          # most list positions are separated by more than the tolerance and must match exactly
    # the fused lists equal the lists taken from the kernel's own materialised scores (same MMA arithmetic): bit-exact
    S = mmx_mod.ops.dense_scores(Q.to(dev), G.to(dev))
    ks, ki = mmx_mod.ops.topk_from_scores(S, 10, g_base=64)
    assert torch.equal(ts, ks) and torch.equal(ti, ki)


@pytest.mark.gpu
@pytest.mark.parametrize("nq,ng,d", [(300, 1500, 768), (513, 5000, 1152), (130, 40000, 256)])
def test_rank_and_topk_in_one_call(mmx_mod, nq, ng, d):
    """mmx_dense_rank_topk: the ranking COUNT pass stores the group maxima on its way, so thresholds, counts, bitmask
    AND the top-k lists come out of one call — every output bit-identical to the separate calls (small gallery,
    32-column groups, 256-column groups; shard offset g_base; ragged GT sets)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(nq + 1)
    G = unit(torch.randn(ng, d, generator=g)).bfloat16().to(dev)
    gt = torch.randint(0, ng, (nq,), generator=g)
    Q = unit(G[gt.to(dev)].float().cpu() + 2.0 * unit(torch.randn(nq, d, generator=g))).bfloat16().to(dev)
    n_gt = torch.randint(1, 4, (nq,), generator=g)                      # 1..3 GT columns per query
    off = torch.zeros(nq + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(n_gt, 0)
    idx = torch.cat([(gt[i] + torch.arange(n_gt[i])) % ng for i in range(nq)]).int() + 64
    off, idx = off.to(dev), idx.to(dev)
    for k in (1, 10):
        sep = mmx_mod.ops.dense_rank(Q, G, off, idx, g_base=64)
        ts, ti = mmx_mod.ops.dense_topk(Q, G, k, g_base=64)
        one = mmx_mod.ops.dense_rank(Q, G, off, idx, g_base=64, top_k=k)
        assert torch.equal(one.thr, sep.thr) and torch.equal(one.cnt_above, sep.cnt_above)
        assert torch.equal(one.bits, sep.bits)
        assert torch.equal(one.top_s, ts) and torch.equal(one.top_i, ti)


@pytest.mark.gpu
def test_dense_topk_mass_ties_take_the_exact_fallback(mmx_mod):
    """200 identical gallery rows tie at the top: the 64-slot candidate list overflows and the row is recomputed exactly;
    ties resolve by column index."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    ng, d = 6000, 128
    G = unit(torch.randn(ng, d, generator=g)).bfloat16()
    dup = torch.arange(100, 300)
    G[dup] = G[100].clone()
    Q = torch.cat([G[100:101], G[4000:4003]])
    ts, ti = mmx_mod.ops.dense_topk(Q.to(dev), G.to(dev), 10)
    assert ti[0].cpu().tolist() == list(range(100, 110)) and float(ts[0].min()) == float(ts[0].max())
    assert ti[1:, 0].cpu().tolist() == [4000, 4001, 4002]


@pytest.mark.gpu
def test_sharded_topk_merge_equals_unsharded(mmx_mod):
    dev = torch.device("cuda:0")
    img, all_txt = synth_dense(1200, 256, 3.0, seed=8)
    Q, G = img.to(dev), all_txt.to(dev)                                  # [1200] queries vs 6000 columns
    full_s, full_i = mmx_mod.ops.dense_topk(Q, G, 10)
    parts_s, parts_i = [], []
    for a, b in mmx_mod.sharded.shard_bounds(G.shape[0], 3):
        s, i = mmx_mod.ops.dense_topk(Q, G[a:b].contiguous(), 10, g_base=a)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = mmx_mod.ops.merge_topk(torch.cat(parts_s, 1), torch.cat(parts_i, 1), 10)
    assert torch.equal(ms, full_s) and torch.equal(mi, full_i)
    # the API switch fills every protocol
    r = mmx_mod.score_and_rank(img, all_txt[0::5].contiguous(), all_txt, {"type": "dense", "name": "t"}, want_top10=True)
    assert torch.equal(r["I2T"].top_i, full_i)
    hit1 = (r["T2I"].top_i[:, 0].cpu() == torch.arange(1200)).numpy()
    assert np.array_equal(hit1, (r["T2I"].cnt_above.cpu() == 0).numpy())   # top-1 is the GT  <=>  nothing scores above it


# ------------------------------------------------------------------------------------------------ gloo, CPU
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_merge(gs, gi, k):
    v = torch.where(gi >= 0, gs, torch.full_like(gs, float("-inf")))
    order = np.lexsort((gi.numpy(), -v.numpy()), axis=1)[:, :k]                     # score desc, then column asc
    o = torch.from_numpy(order)
    return torch.gather(v, 1, o), torch.gather(gi, 1, o)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmx
    rs = np.random.RandomState(1)
    scores = torch.from_numpy(rs.randn(40, 700).astype(np.float32))
    a, b = mmx.sharded.shard_bounds(700, world)[rank]
    lv, li = oracle.topk_lists(scores[:, a:b], 10)
    ms, mi = mmx.sharded.merge_topk(lv, (li + a).to(torch.int32), merge_fn=_cpu_merge)
    wv, wi = oracle.topk_lists(scores, 10)
    ok = torch.equal(ms, wv) and torch.equal(mi.long(), wi)
    Path(out_dir, f"ok{rank}").write_text("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_single_allgather_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert [Path(tmp_path, f"ok{r}").read_text() for r in range(2)] == ["1", "1"]

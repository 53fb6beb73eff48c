"""Randomised shapes for K1 / KR / KT / K3 / K4 / K6 / K6s against torch / numpy restatements on small sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import mmx
from mmx import ops

dev = torch.device("cuda:0")
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rs = np.random.RandomState(seed)
g = torch.Generator().manual_seed(seed)

def unit(x): return x / x.norm(dim=-1, keepdim=True)

for case in range(n_cases):
    # ---------------- K1 + KR + KT on the same random problem
    nq, ng = int(rs.randint(1, 700)), int(rs.randint(1, 1500))
    d = int(rs.choice([8, 16, 64, 72, 128, 256, 768, 1152]))
    Q = unit(torch.randn(nq, d, generator=g)).bfloat16().to(dev)
    G = unit(torch.randn(ng, d, generator=g)).bfloat16().to(dev)
    if rs.rand() < 0.3 and ng > 3:
        G[rs.randint(0, ng, size=3)] = G[0].clone()                        # exact duplicates -> exact ties
    g_base = int(rs.choice([0, 7, 64, 1000]))
    gt = [sorted((rs.choice(ng, size=min(ng, rs.randint(0, 4)), replace=False) + g_base).tolist()) for _ in range(nq)]
    off, idx = ops.csr_from_sets(gt)
    off_d, idx_d = torch.from_numpy(off).to(dev), torch.from_numpy(idx).to(dev)
    S = ops.dense_scores(Q, G)
    r = ops.dense_rank(Q, G, off_d, idx_d, g_base=g_base)
    Sc = S.cpu()
    thr = torch.tensor([max([float(Sc[q, c - g_base]) for c in gt[q]], default=float("-inf")) for q in range(nq)])
    assert torch.equal(r.thr.cpu(), thr), ("K1 thr", case, nq, ng, d)
    above = Sc > thr[:, None]
    assert torch.equal(r.cnt_above.cpu(), above.sum(1).int()), ("K1 cnt", case)
    bits = np.unpackbits(r.bits.cpu().numpy().view(np.uint8), axis=1, bitorder="little")[:, :ng]
    assert np.array_equal(bits.astype(bool), above.numpy()), ("K1 bits", case)
    rk = ops.rank_from_scores(S, off_d, idx_d, g_base=g_base)
    assert torch.equal(rk.thr, r.thr) and torch.equal(rk.cnt_above, r.cnt_above) and torch.equal(rk.bits, r.bits), ("KR", case)
    k = int(rs.choice([1, 5, 10, 32]))
    ts, ti = ops.dense_topk(Q, G, k, g_base=g_base)
    rt = ops.dense_rank(Q, G, off_d, idx_d, g_base=g_base, top_k=k)                 # fused COUNT + group maxima
    assert torch.equal(rt.top_s, ts) and torch.equal(rt.top_i, ti), ("rank_topk lists", case, nq, ng, d, k)
    assert torch.equal(rt.cnt_above, r.cnt_above) and torch.equal(rt.bits, r.bits) and torch.equal(rt.thr, r.thr), ("rank_topk rank", case)
    ks, ki = ops.topk_from_scores(S, k, g_base=g_base)
    assert torch.equal(ts, ks) and torch.equal(ti, ki), ("KT fused vs scores", case, nq, ng, d, k)
    sv, si = torch.sort(Sc, dim=1, descending=True, stable=True)
    kk = min(k, ng)
    assert torch.equal(ks.cpu()[:, :kk], sv[:, :kk]) and torch.equal(ki.cpu()[:, :kk].long(), si[:, :kk] + g_base), ("KT", case)
    ties = ops.exact_tie_counts(r, off_d, idx_d, Q=Q, G=G, g_base=g_base).cpu()
    gt_mask = torch.zeros_like(above)
    for q in range(nq):
        for c in gt[q]:
            gt_mask[q, c - g_base] = True
    want_ties = ((Sc == thr[:, None]) & ~gt_mask).sum(1).int()
    has_gt = torch.tensor([len(x) > 0 for x in gt])
    assert torch.equal(ties[has_gt], want_ties[has_gt]), ("ties", case)
    # ---------------- K3 with random unit maps on the GT-less-free rows
    n_units = int(rs.randint(1, 200))
    r.col2unit = torch.from_numpy(rs.randint(0, n_units, size=ng).astype(np.int32)).to(dev)
    r.q_unit = torch.from_numpy(rs.randint(0, n_units, size=nq).astype(np.int32)).to(dev)
    B = int(rs.randint(1, 20))
    si_h = rs.randint(0, n_units, size=(B, int(rs.randint(1, 400))))
    mult = ops.bincount(torch.from_numpy(si_h.astype(np.int32)).to(dev), n_units)
    got = ops.boot_counts(r, mult, list_len=int(rs.choice([1, 2, 32]))).cpu().numpy()
    m = np.stack([np.bincount(si_h[b], minlength=n_units) for b in range(B)])
    c = m[:, r.col2unit.cpu().numpy()] @ above.numpy().astype(np.int64).T            # [B, nq]
    mq = m[:, r.q_unit.cpu().numpy()]
    want = np.stack([np.minimum(mq, np.maximum(0, kk_ - c)).sum(1) for kk_ in (1, 5, 10)], axis=1)
    assert np.array_equal(got, want), ("K3", case)
    # ---------------- LIST mode: exact counts, lists inside the above-GT set, K3 on lists == the identity above
    cap = int(rs.choice([1, 3, 8, 64]))
    has_gt_d = torch.from_numpy(np.array([len(x) > 0 for x in gt])).to(dev)
    rl = ops.dense_rank_lists(Q, G, off_d, idx_d, g_base=g_base, thr=r.thr, cap=cap)
    assert torch.equal(rl.cnt_above, r.cnt_above), ("lists cnt", case)
    li, ln = rl.list_i[0].cpu().numpy() - g_base, rl.list_n[0].cpu().numpy()
    ab = above.numpy()
    for q in range(nq):
        nv = min(int(ln[q]), cap)
        assert nv == min(int(ab[q].sum()), cap) and ab[q, li[q, :nv]].all() and len(set(li[q, :nv].tolist())) == nv, ("lists", case, q)
    full_map = torch.zeros(g_base + ng, dtype=torch.int32, device=dev)
    full_map[g_base:] = r.col2unit                                                  # lists carry GLOBAL column ids
    rl.col2unit, rl.q_unit = full_map, r.q_unit
    st = {}
    got_l = ops.boot_counts_lists(rl, mult, ops.dense_unlisted_resolver(Q, G, r.thr, full_map, g_base), stats=st).cpu().numpy()
    assert np.array_equal(got_l, want), ("K3 lists", case, cap, st)
    # ---------------- K4 / K6 / K6s
    n = int(rs.randint(1, 1500)); R = int(rs.randint(1, 60))
    data = np.round(rs.normal(40, 3, size=n) * 50) / 50
    idx_h = rs.randint(0, n, size=(R, n)).astype(np.int32)
    means = ops.ci_means(torch.from_numpy(data).to(dev), torch.from_numpy(idx_h).to(dev)).cpu().numpy()
    assert np.array_equal(means, np.array([np.mean(data[idx_h[i]]) for i in range(R)])), ("K4", case, n)
    a, b_ = np.random.RandomState(case), np.random.RandomState(case)
    a.randint(0, 5, size=int(rs.randint(0, 2000))); b_.set_state(a.get_state())
    dr = ops.DeviceRng(a)
    high, cnt = int(rs.choice([2, 3, 1000, 4097, 1 << 20])), int(rs.choice([1, 700, 250_000, 1_300_000]))
    x = dr.randint(high, cnt).cpu().numpy()
    ns, P = int(rs.randint(2, 700)), int(rs.randint(1, 40))
    perm = dr.shuffle_indices(ns, P).cpu().numpy()
    dr.commit()
    assert np.array_equal(x, b_.randint(0, high, size=cnt)), ("K6", case, high, cnt)
    ix = np.arange(ns)
    for p_ in range(P):
        b_.shuffle(ix)
        assert np.array_equal(perm[p_], ix), ("K6s", case, ns, P, p_)
    assert a.randint(0, 1 << 30) == b_.randint(0, 1 << 30), ("state", case)
print(f"{n_cases} random cases ok (seed {seed})")

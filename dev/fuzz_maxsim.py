"""Randomised shapes for both MaxSim kernel variants against a torch fp32 restatement on the GPU (ragged queries and
documents, empty items, zero-padding flags, documents around the 16/256-row thresholds of the SHORT variant)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from mmx import ops

dev = torch.device("cuda:0")
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
g = torch.Generator(device=dev).manual_seed(1)

def reference(Q, q_off, D, d_off, flags):
    nq, nd = len(q_off) - 1, len(d_off) - 1
    out = torch.zeros(nq, nd, device=dev)
    Qf = Q.float()
    for j in range(nd):
        a, b = int(d_off[j]), int(d_off[j + 1])
        if b > a:
            m = (Qf @ D[a:b].float().T).max(1).values
            if flags is not None and flags[j]:
                m = m.clamp_min(0)
        else:
            m = torch.zeros(Q.shape[0], device=dev)
        cs = torch.cat([torch.zeros(1, device=dev, dtype=torch.float64), m.double().cumsum(0)])
        out[:, j] = (cs[q_off[1:]] - cs[q_off[:-1]]).float()
    return out

n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
worst = 0.0
for case in range(n_cases):
    nq = int(rs.randint(1, 50)); nd = int(rs.randint(1, 160))
    qmode, dmode = rs.randint(0, 4), rs.randint(0, 5)
    qlens = {0: rs.randint(0, 40, nq), 1: rs.randint(100, 400, nq), 2: rs.randint(0, 1200, nq), 3: np.full(nq, 20)}[qmode]
    dlens = {0: rs.randint(0, 40, nd), 1: rs.randint(200, 320, nd), 2: rs.randint(0, 1100, nd), 3: np.full(nd, 1030),
             4: np.where(rs.rand(nd) < 0.9, rs.randint(8, 41, nd), rs.randint(300, 900, nd))}[dmode]
    if qlens.sum() == 0: qlens[0] = 3
    Q = torch.randn(int(qlens.sum()), 128, generator=g, device=dev); Q = (Q / Q.norm(dim=-1, keepdim=True)).bfloat16()
    D = torch.randn(max(int(dlens.sum()), 1), 128, generator=g, device=dev); D = (D / D.norm(dim=-1, keepdim=True)).bfloat16()
    D = D[: int(dlens.sum())]
    if rs.rand() < 0.3 and len(D):
        D = -D.abs()                                  # all-negative similarities against positive queries
        Q = Q.abs()
    q_off = torch.from_numpy(np.concatenate([[0], np.cumsum(qlens)]).astype(np.int64)).to(dev)
    d_off = torch.from_numpy(np.concatenate([[0], np.cumsum(dlens)]).astype(np.int64)).to(dev)
    flags = torch.from_numpy((rs.rand(nd) < 0.5).astype(np.uint8)).to(dev) if rs.rand() < 0.6 else None
    got = ops.maxsim_scores(Q, q_off.int(), D, d_off.int(), flags)
    want = reference(Q, q_off, D, d_off, flags)
    err = float(((got - want).abs() / want.abs().clamp_min(1e-2)).max()) if got.numel() else 0.0
    worst = max(worst, err)
    short = int(dlens.sum()) < 256 * nd
    assert err < 1e-3 and bool(torch.isfinite(got).all()), (case, nq, nd, qmode, dmode, short, err)
print(f"{n_cases} random cases ok, worst relative error {worst:.2e}")

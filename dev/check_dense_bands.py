"""K1 over several raster bands (more than 64 row tiles of 256): fused rank == KR on the kernel's own materialised scores."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mmx import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for nq, ng, d in ((40000, 3000, 64), (70001, 1111, 128), (20000, 20000, 72)):
    Q = torch.randn(nq, d, generator=g, device=dev); Q = (Q / Q.norm(dim=-1, keepdim=True)).bfloat16()
    G = torch.randn(ng, d, generator=g, device=dev); G = (G / G.norm(dim=-1, keepdim=True)).bfloat16()
    off = torch.arange(nq + 1, dtype=torch.int32, device=dev)
    idx = (torch.arange(nq, device=dev) * 7 % ng).to(torch.int32)
    r = ops.dense_rank(Q, G, off, idx)
    S = ops.dense_scores(Q, G)
    k = ops.rank_from_scores(S, off, idx)
    ts, ti = ops.dense_topk(Q, G, 10)
    ks, ki = ops.topk_from_scores(S, 10)
    ok = (torch.equal(r.thr, k.thr) and torch.equal(r.cnt_above, k.cnt_above) and torch.equal(r.bits, k.bits)
          and torch.equal(ts, ks) and torch.equal(ti, ki))
    print(nq, ng, d, "fused == materialised:", ok, flush=True)
    assert ok

import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mmx
from mmx import ops
from bench import synth_maxsim, NQ, ND, LQ, LD
dev = torch.device("cuda:0")
nq, nd = 5000, 1000
Q, D, q2d = synth_maxsim(dev, nq, nd, LQ, LD)
q_off = torch.arange(0, nq*LQ+1, LQ, dtype=torch.int32, device=dev)
d_off = torch.arange(0, nd*LD+1, LD, dtype=torch.int32, device=dev)
gt_off = torch.arange(nq+1, dtype=torch.int32, device=dev)
S1 = ops.maxsim_scores(Q, q_off, D, d_off).clone()
S2 = ops.maxsim_scores(Q, q_off, D, d_off).clone()
print("S equal:", torch.equal(S1, S2), (S1-S2).abs().max().item())
r1 = ops.rank_from_scores(S1, gt_off, q2d); r2 = ops.rank_from_scores(S2, gt_off, q2d)
print("thr", torch.equal(r1.thr, r2.thr), "cnt", torch.equal(r1.cnt_above, r2.cnt_above), "bits", torch.equal(r1.bits, r2.bits))
r1.q_unit = q2d
si = torch.from_numpy(np.random.RandomState(42).randint(0, nd, size=(100, nd)).astype(np.int32)).to(dev)
m = ops.bincount(si, nd)
c1 = ops.boot_counts(r1, m); c2 = ops.boot_counts(r1, m)
print("correct equal", torch.equal(c1, c2))
# API path
qv = mmx.FlatMultiVector(Q.cpu(), np.full(nq, LQ)); dv = mmx.FlatMultiVector(D.cpu(), np.full(nd, LD))
res = mmx.run_retrieval_bootstrap(qv, dv, q2d.cpu().numpy(), {"type":"colpali","name":"x"}, 100, rng=np.random.RandomState(42))
vals = (100.0 * c1.double() / nq).cpu().numpy()
print("api equal", np.array_equal(vals, res["_bootstrap_values"]), np.abs(vals-res["_bootstrap_values"]).max())
print(res["base"], r1.recall())

import sys; sys.path.insert(0, "/root/repo")
import torch, mmx
from mmx import ops
from bench import synth_maxsim, NQ, ND, LQ, LD
dev = torch.device("cuda:0")
q_off = torch.arange(0, NQ*LQ+1, LQ, dtype=torch.int32, device=dev)
d_off = torch.arange(0, ND*LD+1, LD, dtype=torch.int32, device=dev)
gt_off = torch.arange(NQ+1, dtype=torch.int32, device=dev)
for s in (3.3, 3.6, 3.9, 4.2):
    Q, D, q2d = synth_maxsim(dev, NQ, ND, LQ, LD, sigma_q=s)
    S = ops.maxsim_scores(Q, q_off, D, d_off)
    r = ops.rank_from_scores(S, gt_off, q2d, want_bits=False)
    print(s, r.recall())

"""One launch of the SHORT-document MaxSim variant (I2T orientation) for ncu: 1000 images x 25000 captions."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mmx import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)
n_img, n_cap, lp, lt = 1000, 25000, 1030, 20
P = unit(torch.randn(n_img * lp, 128, generator=g, device=dev)).bfloat16()
T = unit(torch.randn(n_cap * lt, 128, generator=g, device=dev)).bfloat16()
p_off = torch.arange(0, n_img * lp + 1, lp, dtype=torch.int32, device=dev)
t_off = torch.arange(0, n_cap * lt + 1, lt, dtype=torch.int32, device=dev)
out = torch.empty((n_img, n_cap), dtype=torch.float32, device=dev)
for _ in range(2):
    ops.maxsim_scores(P, p_off, T, t_off, None, out=out)
torch.cuda.synchronize()
print("ok", float(out[3, 5]))

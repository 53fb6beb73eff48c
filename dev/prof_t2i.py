"""One full-size launch of the T2I MaxSim kernel (25 000 x 20 tokens vs 5 000 x 1030 patches) for ncu."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from mmx import ops
dev = torch.device("cuda:0")
Q, D, q2d = bench.synth_maxsim(dev, bench.NQ, bench.ND, bench.LQ, bench.LD)
q_off = torch.arange(0, bench.NQ * bench.LQ + 1, bench.LQ, dtype=torch.int32, device=dev)
d_off = torch.arange(0, bench.ND * bench.LD + 1, bench.LD, dtype=torch.int32, device=dev)
S = torch.empty((bench.NQ, bench.ND), dtype=torch.float32, device=dev)
for _ in range(2):
    ops.maxsim_scores(Q, q_off, D, d_off, None, out=S)
torch.cuda.synchronize()
print("ok", float(S[5, 1]))

"""Times the full-size T2I MaxSim launch (CUDA events) — used with MMX_L2_CHUNK_MB to tune the document chunk size."""
import os, sys
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
for _ in range(3):
    ops.maxsim_scores(Q, q_off, D, d_off, None, out=S)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4):
    ops.maxsim_scores(Q, q_off, D, d_off, None, out=S)
e1.record(); torch.cuda.synchronize()
print(f"MMX_L2_CHUNK_MB={os.environ.get('MMX_L2_CHUNK_MB', 'default')}: {e0.elapsed_time(e1) / 4:.2f} ms", flush=True)

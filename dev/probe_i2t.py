"""Throughput of K2 in the role-swapped I2T orientation (v28 :375-386 / main.py:622): images are the MaxSim queries
(1030 patches each), captions the documents (about 20 tokens each) — a document closes every ~20 accumulator columns."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import mmx
from mmx import ops

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)

def bench(n_img, n_cap, lp, lens, label, iters=3):
    P = unit(torch.randn(n_img * lp, 128, generator=g, device=dev)).bfloat16()
    T = unit(torch.randn(int(lens.sum()), 128, generator=g, device=dev)).bfloat16()
    p_off = torch.arange(0, n_img * lp + 1, lp, dtype=torch.int32, device=dev)
    t_off = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).to(torch.int32).to(dev)
    out = torch.empty((n_img, n_cap), dtype=torch.float32, device=dev)
    for _ in range(2):
        ops.maxsim_scores(P, p_off, T, t_off, None, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.maxsim_scores(P, p_off, T, t_off, None, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flop = 2.0 * P.shape[0] * T.shape[0] * 128
    print(f"{label}: {ms:.1f} ms  {n_img * n_cap / ms * 1e3:.3e} pairs/s  {flop / ms / 1e9:.0f} TFLOP/s", flush=True)
    # spot check against torch on a few images/captions
    for i in (0, n_img - 1):
        for j in (0, 7, n_cap - 1):
            a, b = int(t_off[j]), int(t_off[j + 1])
            want = (P[i * lp:(i + 1) * lp].float() @ T[a:b].float().T).max(1).values.sum()
            assert abs(float(out[i, j]) - float(want)) < 1e-3 * abs(float(want)), (i, j, float(out[i, j]), float(want))

n_img, n_cap = 5000, 25000
bench(n_img, n_cap, 1030, torch.full((n_cap,), 20, dtype=torch.int64), "I2T fixed 20-token captions (5000 x 25000)")
lens = torch.randint(8, 41, (n_cap,), generator=torch.Generator().manual_seed(1))
bench(n_img, n_cap, 1030, lens, "I2T ragged 8-40-token captions (5000 x 25000)")
# T2I for comparison at the same FLOP count: captions as queries
lens20 = torch.full((n_cap,), 20, dtype=torch.int64)
T = unit(torch.randn(n_cap * 20, 128, generator=g, device=dev)).bfloat16()
P = unit(torch.randn(n_img * 1030, 128, generator=g, device=dev)).bfloat16()
t_off = torch.arange(0, n_cap * 20 + 1, 20, dtype=torch.int32, device=dev)
p_off = torch.arange(0, n_img * 1030 + 1, 1030, dtype=torch.int32, device=dev)
out = torch.empty((n_cap, n_img), dtype=torch.float32, device=dev)
for _ in range(2): ops.maxsim_scores(T, t_off, P, p_off, None, out=out)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.maxsim_scores(T, t_off, P, p_off, None, out=out); e1.record(); torch.cuda.synchronize()
print(f"T2I same shapes: {e0.elapsed_time(e1):.1f} ms")

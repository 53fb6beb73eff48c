"""Dev probe: the plain K1 COUNT pass and the LIST-mode pass on the same 10^6 x 10^6 x 1024 operands, interleaved
(C L C L C) so that clock / power drift between the two shows up as a trend instead of a bias."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import mmx
from mmx import ops
dev = torch.device("cuda:0")
N, D, sigma = 1_000_000, 1024, 9.0
g = torch.Generator(device=dev).manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)
img = unit(torch.randn(N, D, generator=g, device=dev))
txt = unit(img + sigma * unit(torch.randn(N, D, generator=g, device=dev))).bfloat16()
G = img.bfloat16().contiguous(); del img
off = torch.arange(N + 1, dtype=torch.int32, device=dev); idx = torch.arange(N, dtype=torch.int32, device=dev)
thr = ops.dense_rank(txt, G, off, idx, want_bits=False, threshold_only=True).thr
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
count = lambda: ops.dense_rank(txt, G, off, idx, want_bits=False, thr=thr)
lists = lambda: ops.dense_rank_lists(txt, G, off, idx, thr=thr)
count(); lists(); torch.cuda.synchronize()
seq = [("count", count), ("lists", lists), ("count", count), ("lists", lists), ("count", count)]
print(json.dumps([{k: round(timed(f), 1)} for k, f in seq]))

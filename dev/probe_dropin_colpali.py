"""The drop-in ColPali call with the reference's own input format: python lists of [L_i, 128] CPU tensors."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import mmx
g = torch.Generator().manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)
docs = list(unit(torch.randn(5000 * 1030, 128, generator=g)).bfloat16().split(1030))
queries = list(unit(torch.randn(5000 * 20, 128, generator=g)).bfloat16().split(20))
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = mmx.compute_similarity_scores(docs, queries, {"type": "colpali"}, None)
    print(f"compute_similarity_scores(5000 docs x 1030, 5000 queries x 20), lists in, CPU fp32 [5000,5000] out: {1e3 * (time.perf_counter() - t0):.0f} ms", flush=True)

"""One dense scoring + bootstrap pass at SigLIP shapes (configs[1]) and one large K1 count pass, for ncu."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import mmx
from mmx import ops
from helpers import synth_dense
dev = torch.device("cuda:0")
N, Dm = 5000, 1152
img, all_txt = synth_dense(N, Dm, 10.0, seed=0)
t2i = all_txt[0::5].contiguous()
img_d, t2i_d, all_d = img.to(dev), t2i.to(dev), all_txt.to(dev)
si = torch.from_numpy(np.random.RandomState(42).randint(0, N, size=(1000, N)).astype(np.int32)).to(dev)
for _ in range(2):
    ranks = mmx.score_and_rank(img_d, t2i_d, all_d, {"type": "siglip", "name": "p"})
    mult = ops.bincount(si, N)
    cols = [ops.boot_counts(ranks[p], mult) for p in ("T2I", "I2T", "I2T_Sym")]
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.randn(25000, 768, generator=g, device=dev); Q = (Q / Q.norm(dim=-1, keepdim=True)).bfloat16()
off = torch.arange(25001, dtype=torch.int32, device=dev); idx = torch.arange(25000, dtype=torch.int32, device=dev)
for _ in range(2):
    r = ops.dense_rank(Q, Q, off, idx)
torch.cuda.synchronize()
print("ok", int(r.cnt_above.sum()), int(cols[0].sum()))

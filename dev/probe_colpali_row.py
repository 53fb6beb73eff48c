"""The reference's complete ColPali row (main.py:478-667 after the encoders; "~15-20 hours" on an A40, main.py:26):
5000 images x 1030 patches, 25 000 captions (8-40 tokens, the first of every five is the T2I query), T2I + I2T
(images as MaxSim queries) + I2T_Sym, 1000-iteration bootstrap, nine CIs, failure analysis — from pinned host buffers."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import mmx

torch.cuda.set_device(0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)
N, LP = 5000, 1030
lens = torch.randint(8, 41, (5 * N,), generator=torch.Generator().manual_seed(1))
P = unit(torch.randn(N * LP, 128, generator=g, device=dev))
owner = torch.repeat_interleave(torch.arange(5 * N), lens).to(dev) // 5            # caption token -> image
rows = torch.randint(0, LP, (int(lens.sum()),), generator=g, device=dev)
T = unit(P[owner * LP + rows] + 3.6 * unit(torch.randn(int(lens.sum()), 128, generator=g, device=dev))).bfloat16()
P = P.bfloat16()
off = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
t2i_rows = torch.cat([torch.arange(int(off[5 * i]), int(off[5 * i + 1])) for i in range(N)])
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
img_fv = mmx.FlatMultiVector(pin(P), np.full(N, LP))
all_fv = mmx.FlatMultiVector(pin(T), lens.numpy())
t2i_fv = mmx.FlatMultiVector(pin(T[t2i_rows.to(dev)]), lens.numpy()[0::5])
del P, T
caps = ["a red bus next to two people %d" % i for i in range(N)]
m_info = {"type": "colpali", "name": "ColPali-synth"}
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    row = mmx.run_bootstrap_from_embeddings(img_fv, t2i_fv, all_fv, m_info, 1000, rng=np.random.RandomState(42), t2i_captions=caps,
                                            encoding_time=1.0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"row {it}: {dt * 1e3:.1f} ms", flush=True)
print(json.dumps({"workload": "full ColPali row: T2I 5000x5000 + I2T 5000x25000 MaxSim, 3 protocols, 1000-iteration bootstrap, 9 CIs",
                  "seconds": dt, "pairs": N * N + N * 5 * N, "pairs_per_s": (N * N + 5 * N * N) / dt,
                  "T2I_R@1": [row["T2I_R@1_mean"], row["T2I_R@1_lower"], row["T2I_R@1_upper"]],
                  "I2T_R@1": [row["I2T_R@1_mean"], row["I2T_R@1_lower"], row["I2T_R@1_upper"]],
                  "base": row["_base_metrics"], "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))

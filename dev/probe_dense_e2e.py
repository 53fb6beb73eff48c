"""Where the dense (configs[1]) end-to-end time goes: cProfile of mmx.run_bootstrap_from_embeddings from pinned host tensors."""
import cProfile, pstats, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import mmx
from helpers import synth_dense

torch.cuda.set_device(0)
N, Dm = 5000, 1152
img, all_txt = synth_dense(N, Dm, 10.0, seed=0)
t2i = all_txt[0::5].contiguous()
m_info = {"type": "siglip", "name": "SigLIP-synth"}
imgp, t2ip, allp = img.pin_memory(), t2i.pin_memory(), all_txt.pin_memory()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mmx.run_bootstrap_from_embeddings(imgp, t2ip, allp, m_info, 1000, rng=np.random.RandomState(42))
    torch.cuda.synchronize(); print("e2e %.2f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
pr = cProfile.Profile(); pr.enable()
mmx.run_bootstrap_from_embeddings(imgp, t2ip, allp, m_info, 1000, rng=np.random.RandomState(42))
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)

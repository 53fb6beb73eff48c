"""e2e (host buffers in, CI dict out) of the ColPali headline step with K6 vs K6p, and with the docs already on the device."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import mmx
from mmx import ops
import bench

import gc
_g = {}
def _cb(phase, info):
    if phase == "start":
        _g["t"] = time.perf_counter()
    else:
        dt = 1e3 * (time.perf_counter() - _g["t"])
        if dt > 5:
            print(f"   [gc gen{info['generation']} {dt:.1f} ms]", flush=True)
gc.callbacks.append(_cb)
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
NQ, ND, LQ, LD = bench.NQ, bench.ND, bench.LQ, bench.LD
Q, D, q2d = bench.synth_maxsim(dev, NQ, ND, LQ, LD)
Qh = torch.empty(Q.shape, dtype=Q.dtype, pin_memory=True).copy_(Q)
Dh = torch.empty(D.shape, dtype=D.dtype, pin_memory=True).copy_(D)
q2d_h = q2d.cpu().numpy()
m_info = {"type": "colpali", "name": "x"}
for label, qv, dv, par_min in (("host docs, K6p", Qh, Dh, 200_000), ("host docs, K6 single CTA", Qh, Dh, 1 << 62),
                               ("device docs, K6p", Q, D, 200_000), ("host docs, K6p again", Qh, Dh, 200_000)):
    ops.DeviceRng.PAR_MIN_COUNT = par_min
    a = mmx.FlatMultiVector(qv, np.full(NQ, LQ))
    b = mmx.FlatMultiVector(dv, np.full(ND, LD))
    for i in range(6):
        tm = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mmx.run_retrieval_bootstrap(a, b, q2d_h, m_info, 1000, rng=np.random.RandomState(42), timings=tm)
        torch.cuda.synchronize()
        print(f"{label}: {1e3 * (time.perf_counter() - t0):.1f} ms wall, score {tm['score_ms']:.1f} ms, " + " ".join(f"{k}={v:.0f}" for k, v in tm["wall_ms"].items()), flush=True)

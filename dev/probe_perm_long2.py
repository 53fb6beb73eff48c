"""Round-2 statistics probe (one GPU): permutation_test on per-query vectors (n1 = n2 = 25 000, configs[4]) with the
device shuffle replay, phase by phase (CUDA events), against numpy's host shuffle; K4 (CI resample means) bandwidth.

    python dev/probe_perm_long2.py [out.json]
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import mmx
from mmx import ops

torch.cuda.set_device(0)
dev = torch.device("cuda:0")
out = {}


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


# ---- K4: CI resample means, [M, 10000, 1000] int32 indices
rs = np.random.RandomState(0)
for M in (3, 9):
    data = torch.from_numpy(np.round(rs.normal(40, 1, size=(M, 1000)) * 50) / 50).to(dev)
    idx = torch.randint(0, 1000, (M, 10000, 1000), dtype=torch.int32, device=dev)
    ms, _ = timed(lambda: ops.ci_means_batched(data, idx), reps=20)
    gb = idx.numel() * 4 / 1e9
    out[f"k4_batched_{M}"] = {"ms": ms, "index_GB": gb, "GBps": gb / (ms * 1e-3)}
    print(f"K4 batch {M}: {ms * 1e3:.1f} us, {gb / (ms * 1e-3):.0f} GB/s", flush=True)

# ---- permutation tests
for n_each, P in ((1000, 10000), (25000, 10000)):
    d1 = (rs.rand(n_each) < 0.47).astype(np.float64)
    d2 = (rs.rand(n_each) < 0.46).astype(np.float64)
    n = 2 * n_each
    for it in range(2):
        g = np.random.RandomState(42)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p = mmx.permutation_test(d1, d2, P, rng=g)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    # phases
    drng = ops.DeviceRng(np.random.RandomState(42), dev)
    t_sh, perm = timed(lambda: ops.DeviceRng(np.random.RandomState(42), dev).shuffle_indices(n, P))
    comb = torch.from_numpy(np.concatenate([d1, d2])).to(dev)
    t_k5, _ = timed(lambda: ops.perm_stats(comb, n_each, perm), reps=3)
    rec = {"n1": n_each, "n2": n_each, "P": P, "wall_ms": wall * 1e3, "p": float(p), "shuffle_indices_ms": t_sh,
           "perm_stats_ms": t_k5, "perm_GB": perm.numel() * 4 / 1e9, "k5_GBps": perm.numel() * 4 / 1e9 / (t_k5 * 1e-3),
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    del perm
    torch.cuda.empty_cache()
    if n_each >= 25000:
        idx = np.arange(n)
        g2 = np.random.RandomState(42)
        t0 = time.perf_counter()
        for _ in range(100):
            g2.shuffle(idx)
        rec["numpy_host_shuffle_s_extrapolated"] = (time.perf_counter() - t0) / 100 * P
    out[f"perm_{n_each}"] = rec
    print(rec, flush=True)

if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))

"""permutation_test on long vectors (per-query rank vectors of 12 500- and 25 000-query models, configs[4]) with the
LONG device shuffle path (n > 16 384) — wall time, 10 000 permutations."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import mmx

torch.cuda.set_device(0)
mmx.api.DEVICE_SHUFFLE_MAX_N = 65536
rs = np.random.RandomState(0)
for n_each, P in ((12500, 10000), (25000, 10000)):
    d1, d2 = rs.normal(size=n_each), rs.normal(0.01, 1.0, size=n_each)
    for it in range(2):
        g = np.random.RandomState(42)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p = mmx.permutation_test(d1, d2, P, rng=g)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{n_each} vs {n_each}, {P} permutations: {dt * 1e3:.0f} ms on the device (p = {p}); "
          f"peak memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
    if n_each == 12500:
        idx = np.arange(2 * n_each); g2 = np.random.RandomState(42)
        t0 = time.perf_counter()
        for _ in range(200): g2.shuffle(idx)
        print(f"  numpy shuffle on the host: {(time.perf_counter() - t0) / 200 * P:.1f} s for {P} permutations (extrapolated from 200)", flush=True)

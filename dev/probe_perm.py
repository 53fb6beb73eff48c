"""Times mmx.permutation_test (K6s + K5) for the reference's shapes; under ncu the launch list shows the stage times."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import mmx
torch.cuda.set_device(0)
rs = np.random.RandomState(0)
for n1, n2 in ((1000, 1000), (300, 300), (2000, 2000), (5000, 5000)):
    d1, d2 = rs.normal(41.5, 0.7, n1), rs.normal(41.55, 0.7, n2)
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p = mmx.permutation_test(d1, d2, 10000, rng=np.random.RandomState(7))
        torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    import oracle
    p_ref = oracle.permutation_test(d1, d2, 10000, rng=np.random.RandomState(7)) if n1 <= 1000 or n1 == 5000 else p
    print(f"{n1} vs {n2}: p={p} (numpy path {p_ref}, {1e3 * (time.perf_counter() - t0):.0f} ms)  device path {dt:.1f} ms", flush=True)

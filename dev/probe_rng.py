"""Times K6 (single CTA) against K6p (jump-ahead) on the draw counts of one dense model row."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import mmx
from mmx import ops

torch.cuda.set_device(0)
for high, count in ((5000, 5_000_000), (1000, 10_000_000), (1000, 90_000_000), (100000, 100_000_000)):
    for par in (False, True):
        rs = np.random.RandomState(42)
        d = ops.DeviceRng(rs)
        d.randint(high, count, parallel=par)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = d.randint(high, count, parallel=par)
        e1.record()
        torch.cuda.synchronize()
        print(f"high={high} count={count} parallel={par}: {e0.elapsed_time(e1):.3f} ms "
              f"({count / e0.elapsed_time(e1) / 1e6:.2f} Gdraws/s)", flush=True)
t0 = time.perf_counter(); np.random.RandomState(1).randint(0, 1000, size=10_000_000); print("numpy 1e7 draws: %.1f ms" % (1e3 * (time.perf_counter() - t0)))

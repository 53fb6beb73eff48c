"""Golden vectors for the legacy metric variants, produced by the REFERENCE's own v19 function.

Run in the build container only:  python tests/golden/make_golden_legacy.py
Imports /root/reference/run_benchmark_grand_slam_v19_fixed.py (compute_bidirectional_metrics, :110-191) and feeds it
the committed dense.npz score matrices plus a ragged ground-truth case; stores the returned strings.
"""
import json
from pathlib import Path

import numpy as np
import torch

from make_golden import load_reference

HERE = Path(__file__).resolve().parent

if __name__ == "__main__":
    v19 = load_reference("run_benchmark_grand_slam_v19_fixed.py", "ref_v19")
    z = np.load(HERE / "dense.npz")
    s_t2i, s_i2t = torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"])
    n = s_t2i.shape[0]
    out = {}
    out["blocks5"] = v19.compute_bidirectional_metrics(None, None, s_t2i, s_i2t, list(range(n)), None,
                                                       {i: list(range(5 * i, 5 * i + 5)) for i in range(n)})
    out["defaults"] = v19.compute_bidirectional_metrics(None, None, s_t2i, s_t2i.t().contiguous())
    rs = np.random.RandomState(11)
    q2i = rs.randint(0, n, size=n).tolist()
    i2c = {i: sorted(rs.choice(5 * n, size=rs.randint(1, 7), replace=False).tolist()) for i in range(n)}
    out["ragged"] = v19.compute_bidirectional_metrics(None, None, s_t2i, s_i2t, q2i, None, i2c)
    out["ragged_q2i"] = q2i
    out["ragged_i2c"] = {str(k): v for k, v in i2c.items()}
    out["narrow"] = v19.compute_bidirectional_metrics(None, None, s_t2i[:, :7].contiguous(), s_i2t[:, :4].contiguous(),
                                                      [i % 7 for i in range(n)], None, {i: [i % 4] for i in range(n)})
    (HERE / "legacy_v19.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print(out["blocks5"], out["narrow"])

"""Random geometries for the LIST-mode K3 kernel (one warp per query, iterations across the lanes) against the literal
numpy walk of tests/test_gpu_parity.py::check_boot_counts_lists.   python dev/fuzz_k3_lists.py [seed] [n_cases]"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
mmx = importlib.import_module("multimodal-embedding_b200")
from test_gpu_parity import check_boot_counts_lists  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rs = np.random.RandomState(seed)
tot = 0
for case in range(n_cases):
    n_shards = int(rs.randint(1, 9))
    cap = int(rs.choice([1, 2, 3, 8, 31, 32, 33, 64, 100]))
    L = int(rs.randint(1, cap + 1))
    nq = int(rs.choice([1, 7, 8, 9, 63, 500, 3000]))
    B = int(rs.choice([1, 3, 4, 5, 127, 128, 129, 257, 1000]))
    n_units = int(rs.choice([1, 2, 127, 128, 129, 1000, 5003]))
    tot += check_boot_counts_lists(mmx, n_shards, nq, cap, L, B, n_units, seed=seed * 1000 + case)
    print(case, n_shards, nq, cap, L, B, n_units, "ok", flush=True)
print("all ok; unsettled pairs seen:", tot)

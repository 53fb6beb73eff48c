"""permutation_test (K6s + K5) vs the oracle (numpy shuffles + np.mean) on random sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import mmx, oracle
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    n1, n2 = int(rs.randint(1, 2500)), int(rs.randint(1, 2500))
    P = int(rs.randint(1, 120))
    d1 = np.round(rs.normal(40, 1, n1) * 50) / 50
    d2 = np.round(rs.normal(40.05, 1, n2) * 50) / 50 if rs.rand() < 0.7 else (rs.rand(n2) < 0.4).astype(np.float64)
    a, b = np.random.RandomState(case), np.random.RandomState(case)
    got = mmx.permutation_test(d1, d2, P, rng=a)
    want = oracle.permutation_test(d1, d2, P, rng=b)
    assert got == want and a.randint(0, 1 << 30) == b.randint(0, 1 << 30), (case, n1, n2, P, got, want)
print("permutation fuzz ok")

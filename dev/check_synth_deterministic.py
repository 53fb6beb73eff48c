"""The synthetic ColPali corpus of the cfg5 workloads is the same at every shard count and in every run."""
import importlib
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
workloads = importlib.import_module("multimodal-embedding_b200.workloads")
sharded = importlib.import_module("multimodal-embedding_b200.sharded")
dev = torch.device("cuda:0")
nq, nd, lq, ld = 3000, 700, 20, 37         # many queries per document: plenty of tokens landing on the same patch row
Q0, D0, q0 = workloads.synth_colpali_shard(dev, nq, nd, lq, ld, 0, nd, seed=5, sigma=2.0)
for world in (1, 2, 3, 8):
    parts = []
    for a, b in sharded.shard_bounds(nd, world):
        Q, D, q = workloads.synth_colpali_shard(dev, nq, nd, lq, ld, a, b, seed=5, sigma=2.0)
        assert torch.equal(Q, Q0) and torch.equal(q, q0)
        parts.append(D)
    assert torch.equal(torch.cat(parts), D0), world
print("synthetic corpus identical at 1, 2, 3, 8 shards and across runs")

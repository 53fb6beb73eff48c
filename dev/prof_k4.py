"""ncu target: K4 batched CI means, [3, 10000, 1000] indices."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from mmx import ops
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
data = torch.from_numpy(np.round(rs.normal(40, 1, size=(3, 1000)) * 50) / 50).to(dev)
idx = torch.randint(0, 1000, (3, 10000, 1000), dtype=torch.int32, device=dev)
for _ in range(3):
    out = ops.ci_means_batched(data, idx)
torch.cuda.synchronize()
print("ok", float(out.sum()))

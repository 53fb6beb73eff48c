"""ncu target: one device shuffle replay of a long array (K6s long path): python dev/prof_shuffle_long.py [n] [P]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from mmx import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 300
torch.cuda.set_device(0)
d = ops.DeviceRng(np.random.RandomState(42), torch.device("cuda:0"))
perm = d.shuffle_indices(n, P)
torch.cuda.synchronize()
print("ok", int(perm[-1, :5].sum()))

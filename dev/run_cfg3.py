"""configs[2] T2I on N GPUs (gallery-sharded, LIST-mode bootstrap), with the per-phase device times of rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        dev/run_cfg3.py [out.json] [--repeat 2]
"""
import argparse
import importlib
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
mmx = importlib.import_module("multimodal-embedding_b200")
workloads = importlib.import_module("multimodal-embedding_b200.workloads")

ap = argparse.ArgumentParser()
ap.add_argument("out", nargs="?", default=None)
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--repeat", type=int, default=2)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) if \
    (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {}
runs = []
for _ in range(args.repeat):
    sec = workloads.cfg3_section(dev, dist.group.WORLD if world > 1 else None, n=args.n, i2t=False)
    if rank == 0:
        runs.append(sec)
        print(json.dumps(sec), flush=True)
if rank == 0 and args.out:
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(runs, indent=1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

"""BASELINE.json configs[2]: 1024-d dense scoring scaled to a 1M-image gallery, gallery-sharded over the GPUs of one
box (torchrun). 1M caption-0 queries (T2I): thresholds all-reduce(MAX), counts all-reduce(SUM) -> Recall@K, plus the
per-shard top-10 lists merged after one all-gather. CUDA-event time, max over ranks."""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import mmx
from mmx import ops, sharded

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, D, sigma = 1_000_000, 1024, 9.0
g = torch.Generator(device=dev).manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)
img = unit(torch.randn(N, D, generator=g, device=dev))
txt = unit(img + sigma * unit(torch.randn(N, D, generator=g, device=dev))).bfloat16()
a, b = sharded.shard_bounds(N, world)[rank]
G = img[a:b].bfloat16().contiguous(); del img
off = torch.arange(N + 1, dtype=torch.int32, device=dev); idx = torch.arange(N, dtype=torch.int32, device=dev)
def step(topk):
    loc = ops.dense_rank(txt, G, off, idx, want_bits=False, g_base=a, threshold_only=True)
    thr = sharded.merge_thresholds(loc.thr)
    loc = ops.dense_rank(txt, G, off, idx, want_bits=False, g_base=a, thr=thr, top_k=10 if topk else None)
    cnt = loc.cnt_above.clone()
    if world > 1: dist.all_reduce(cnt)
    res = [int((cnt < k).sum()) for k in (1, 5, 10)]
    if topk:
        ts, ti = sharded.merge_topk(loc.top_s, loc.top_i)
        res.append(int((ti[:, 0].long() == torch.arange(N, device=dev)).sum()))
    return res
for topk in (False, True):
    step(topk); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); res = step(topk); e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": "dense 1M x 1M x 1024 T2I, gallery-sharded", "n_gpus": world, "topk": topk, "ms": float(ms),
                          "pairs_per_s": N * N / float(ms) * 1e3, "tflops_aggregate": 2.0 * N * N * D / float(ms) / 1e9 * (2 if topk else 1) if not topk else None,
                          "R@1/5/10": [100.0 * r / N for r in res[:3]], "top1_is_gt": res[3] if topk else None}), flush=True)
if world > 1: dist.destroy_process_group()

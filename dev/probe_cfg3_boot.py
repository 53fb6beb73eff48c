"""BASELINE.json configs[2] WITH the bootstrap: 10^6 caption queries x 10^6-image gallery x 1024, 1000 iterations.
The [10^6, 10^6] above-GT bitmask would be 125 GB, so the path runs in LIST mode: K1 keeps 64 above-GT
columns per query next to the exact count, K3 walks those lists. Device-resident, CUDA-event time per phase, one GPU
or gallery-sharded under torchrun."""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist
import mmx
from mmx import ops, sharded

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = int(os.environ.get("MMX_N", "1000000")); D, sigma, B = 1024, 9.0, int(os.environ.get("MMX_B", "1000"))
g = torch.Generator(device=dev).manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)
img = unit(torch.randn(N, D, generator=g, device=dev))
txt = unit(img + sigma * unit(torch.randn(N, D, generator=g, device=dev))).bfloat16()
bounds = sharded.shard_bounds(N, world); a, b = bounds[rank]
G = img[a:b].bfloat16().contiguous(); del img
off = torch.arange(N + 1, dtype=torch.int32, device=dev); idx = torch.arange(N, dtype=torch.int32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]

def step():
    ev[0].record()
    loc = ops.dense_rank(txt, G, off, idx, want_bits=False, g_base=a, threshold_only=True)
    thr = sharded.merge_thresholds(loc.thr)
    ev[1].record()
    loc = ops.dense_rank_lists(txt, G, off, idx, g_base=a, thr=thr)
    ranks = sharded.merge_ranks(loc, bounds)
    ev[2].record()
    recount = ops.dense_unlisted_resolver(txt, G, thr, None, a)
    if rank != 0:
        sharded.unlisted_round(recount, None, None)
        return None
    drng = ops.DeviceRng(np.random.RandomState(42), dev)
    si = drng.randint(N, B * N).view(B, N)                      # the reference's choice(N, N) per iteration, numpy's stream
    ev[3].record()
    mult = ops.bincount(si, N); del si
    ev[4].record()
    rounds, st = [], {}
    def resolve(p, m):
        rounds.append(1); return sharded.unlisted_round(recount, p, m)
    correct = ops.boot_counts_lists(ranks, mult, resolve, stats=st)
    if not rounds: sharded.unlisted_round(recount, None, None)
    ev[5].record()
    denom = mult.sum(dim=1, dtype=torch.int64).to(torch.float64)           # identity GT: every drawn image brings its query
    vals = (100.0 * correct.to(torch.float64) / denom[:, None]).cpu().numpy()
    return ranks, vals, st

for it in range(2):
    out = step(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
# the plain COUNT pass (no lists, no bitmask) on the same operands, for comparison
thr_all = sharded.merge_thresholds(ops.dense_rank(txt, G, off, idx, want_bits=False, g_base=a, threshold_only=True).thr)
pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ops.dense_rank(txt, G, off, idx, want_bits=False, g_base=a, thr=thr_all)
pe[0].record(); plain = ops.dense_rank(txt, G, off, idx, want_bits=False, g_base=a, thr=thr_all); pe[1].record()
torch.cuda.synchronize()
if rank == 0:
    ranks, vals, st = out
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
    cnt = ranks.cnt_above
    print(json.dumps({"workload": f"dense {N} x {N} x {D} T2I + {B}-iteration bootstrap, LIST mode (cap {ops.LIST_CAP})",
                      "n_gpus": world, "ms": {"threshold_pass": ms[0], "count_and_list_pass": ms[1], "numpy_stream_on_device": ms[2],
                                              "bincount": ms[3], "K3_lists": ms[4], "total": sum(ms),
                                              "plain_count_pass_for_comparison": pe[0].elapsed_time(pe[1])},
                      "base_R@1/5/10": [100.0 * int((cnt < k).sum()) / N for k in (1, 5, 10)],
                      "boot_mean_R@1/5/10": vals.mean(0).tolist(), "boot_std": vals.std(0).tolist(),
                      "queries_with_truncated_list": int((cnt > ops.LIST_CAP).sum()), "unresolved_pairs": st.get("unresolved_pairs")}), flush=True)
if world > 1: dist.destroy_process_group()

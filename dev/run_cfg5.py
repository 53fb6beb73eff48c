"""BASELINE.json configs[4] as stated: ColPali MaxSim at a 100 000-document synthetic corpus (25 000 queries x 20 tokens,
1030 patches), doc-sharded over the GPUs of one box, 1000-iteration bootstrap per model, 7 'models' (7 seeds / noise
levels), then the 21 pairwise permutation tests of main.py:281-302 on BOTH the per-iteration Recall@1 arrays (n = 2 x 1000)
and the per-query Recall@1 vectors (n = 2 x 25 000) — 42 tests, numpy's global stream continued test after test.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        dev/run_cfg5.py [out.json] [--models 7] [--check-host 1]

--check-host k: the first k per-query tests are also run with the reference's host shuffles (oracle-free: numpy itself)
and must give the same p-value and leave the generator in the same state.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist

import mmx
from mmx import workloads

ap = argparse.ArgumentParser()
ap.add_argument("out", nargs="?", default=None)
ap.add_argument("--models", type=int, default=7)
ap.add_argument("--docs", type=int, default=100000)
ap.add_argument("--queries", type=int, default=25000)
ap.add_argument("--iterations", type=int, default=1000)
ap.add_argument("--check-host", type=int, default=1)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NQ, ND, LQ, LD = args.queries, args.docs, 20, 1030

# lazy initialisation outside the timing: kernels, jump tables, NCCL buffers
workloads.colpali_model(dev, 500, 800, LQ, LD, 50, seed=1, sigma=3.0, rng=np.random.RandomState(1))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()

rng = np.random.RandomState(42)          # the reference's one global stream (SEED = 42, main.py:99)
t_all = time.perf_counter()
rows, hits, phases = [], [], []
for m in range(args.models):
    t0 = time.perf_counter()
    row, h, t = workloads.colpali_model(dev, NQ, ND, LQ, LD, args.iterations, seed=100 + m, sigma=3.5 + 0.1 * m, rng=rng,
                                        name=f"model{m}")
    t["wall_ms"] = 1e3 * (time.perf_counter() - t0)
    rows.append(row)
    hits.append(h)
    phases.append(t)
    if rank == 0:
        print(f"model {m}: {t['wall_ms'] / 1e3:.2f} s (synthesize {t['synthesize_ms'] / 1e3:.2f}, score+rank+bootstrap "
              f"{t['score_rank_bootstrap_ms'] / 1e3:.2f}; K2 {t.get('maxsim_kernel_ms', 0) / 1e3:.2f}) R@1 {row['base'][1]:.2f} "
              f"CI [{row['R@1_lower']:.3f}, {row['R@1_upper']:.3f}]", flush=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t_models = time.perf_counter() - t_all
peak = torch.tensor([torch.cuda.max_memory_allocated() / 1e9], device=dev)
if world > 1:
    dist.all_reduce(peak, op=dist.ReduceOp.MAX)

out = None
if rank == 0:
    tests = {"per_iteration_n2000": [], "per_query_n50000": []}
    t_tests = time.perf_counter()
    check = args.check_host
    for i in range(args.models):
        for j in range(i + 1, args.models):
            for kind, a, b in (("per_iteration_n2000", rows[i]["_bootstrap_values"][:, 0], rows[j]["_bootstrap_values"][:, 0]),
                               ("per_query_n50000", hits[i], hits[j])):
                ref_p = None
                if kind == "per_query_n50000" and check > 0:
                    # the reference itself (main.py:281-302) on a copy of the generator
                    g2 = np.random.RandomState()
                    g2.set_state(rng.get_state())
                    obs = np.abs(np.mean(a) - np.mean(b))
                    comb = np.concatenate([a, b])
                    diffs = []
                    for _ in range(10000):
                        g2.shuffle(comb)
                        diffs.append(np.abs(np.mean(comb[:len(a)]) - np.mean(comb[len(a):])))
                    ref_p = float(np.mean(np.array(diffs) >= obs))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                p = mmx.permutation_test(a, b, rng=rng)
                torch.cuda.synchronize()
                rec = {"pair": [i, j], "p": float(p), "ms": 1e3 * (time.perf_counter() - t0)}
                if ref_p is not None:
                    rec["reference_p"] = ref_p
                    g3 = np.random.RandomState()
                    g3.set_state(rng.get_state())
                    rec["same_p_and_generator_state"] = bool(ref_p == float(p) and g2.randint(0, 1 << 30) == g3.randint(0, 1 << 30))
                    check -= 1
                tests[kind].append(rec)
    torch.cuda.synchronize()
    t_tests = time.perf_counter() - t_tests
    out = {"workload": f"configs[4]: ColPali MaxSim {NQ} q x {ND} docs x {LD} patches, doc-sharded x{world}, {args.iterations}-iter "
                       f"bootstrap, {args.models} models, {len(tests['per_iteration_n2000'])} + {len(tests['per_query_n50000'])} permutation tests",
           "n_gpus": world, "total_s": t_models + t_tests, "models_s": t_models, "tests_s": t_tests,
           "per_model": phases, "peak_mem_GB_per_gpu": float(peak.item()),
           "tests": tests,
           "per_query_test_ms_mean": float(np.mean([r["ms"] for r in tests["per_query_n50000"]])) if tests["per_query_n50000"] else None,
           "per_iteration_test_ms_mean": float(np.mean([r["ms"] for r in tests["per_iteration_n2000"]])) if tests["per_iteration_n2000"] else None,
           "base_R@1": [r["base"][1] for r in rows], "R@1_ci": [[r["R@1_lower"], r["R@1_upper"]] for r in rows]}
    print(json.dumps({k: v for k, v in out.items() if k not in ("tests", "per_model")}), flush=True)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(out, indent=1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

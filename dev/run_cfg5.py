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
ap.add_argument("--prefetch", type=int, default=1, help="replay the per-query tests' shuffles on rank 0 while ranks 1.. score "
                "(rank 0 then holds no documents); 0 = every rank scores, the tests run afterwards")
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
workloads.colpali_model(dev, NQ, 800, LQ, LD, 50, seed=1, sigma=3.0, rng=np.random.RandomState(1))   # full query count: the
# per-query buffers of the timed models come out of the allocator's cache (a cudaMalloc would wait for the shuffle chain's kernel)
_w = mmx.ops.DeviceRng(np.random.RandomState(1), dev)
_w._par_buffers(True)                     # the 512-block jump table (requests beyond 2.4e7 words) is uploaded once
_w.randint(ND, 40_000_000)
_side = mmx.api._aux_stream(dev, "rng")   # the models' draws are made on this stream: put their buffers into its pool now —
with torch.cuda.stream(_side):            # a cudaMalloc later would wait for whatever kernel the shuffle chain is running
    mmx.api.skip_bootstrap_draws(_w, ND, args.iterations)
del _w
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()

rng = np.random.RandomState(42)          # the reference's one global stream (SEED = 42, main.py:99)
n_pairs = args.models * (args.models - 1) // 2
prefetch = bool(args.prefetch) and world > 1
bounds = None
pre = None
t_all = time.perf_counter()
if prefetch:
    # numpy's stream does not depend on the data: rank 0 replays the shuffles of the 21 per-query tests (they follow the seven
    # models' bootstrap draws in the stream) while ranks 1.. score the documents
    from mmx import sharded
    bounds = [(0, 0)] + sharded.shard_bounds(ND, world - 1)
    if rank == 0:
        pre = mmx.ops.ShufflePrefetch(rng, [(2 * NQ, 10000)] * n_pairs, dev,
                                      skip=lambda d: [mmx.api.skip_bootstrap_draws(d, ND, args.iterations)
                                                      for _ in range(args.models)]).start()
rows, hits, phases = [], [], []
for m in range(args.models):
    t0 = time.perf_counter()
    row, h, t = workloads.colpali_model(dev, NQ, ND, LQ, LD, args.iterations, seed=100 + m, sigma=3.6 + 0.005 * m, rng=rng,
                                        name=f"model{m}", bounds=bounds)
    t["wall_ms"] = 1e3 * (time.perf_counter() - t0)
    t["host_t_start_end_s"] = [t0 - t_all, time.perf_counter() - t_all]
    rows.append(row)
    hits.append(h)
    phases.append(t)
    if rank == 0:
        print(f"model {m}: {t['wall_ms'] / 1e3:.2f} s (synthesize {t['synthesize_ms'] / 1e3:.2f}, score+rank+bootstrap "
              f"{t['score_rank_bootstrap_ms'] / 1e3:.2f}; K2 {t.get('maxsim_kernel_ms', 0) / 1e3:.2f}) R@1 {row['base'][1]:.2f} "
              f"CI [{row['R@1_lower']:.3f}, {row['R@1_upper']:.3f}]", flush=True)
if world > 1:
    dist.barrier()
torch.cuda.current_stream().synchronize()
t_models = time.perf_counter() - t_all
peak = torch.tensor([torch.cuda.max_memory_allocated() / 1e9], device=dev)
if world > 1:
    dist.all_reduce(peak, op=dist.ReduceOp.MAX)

out = None
if rank == 0:
    tests = {"per_query_n50000": [], "per_iteration_n2000": []}
    t_tests = time.perf_counter()
    pairs = [(i, j) for i in range(args.models) for j in range(i + 1, args.models)]
    if pre is not None:
        assert pre.starts_where(rng), "the dry run of the bootstrap draws does not match what the models consumed"
    # ---- per-query Recall@1 vectors (n = 2 x 25 000) first, then the per-iteration arrays (n = 2 x 1000)
    state_at_tests = rng.get_state()
    for k, (i, j) in enumerate(pairs):
        a, b = hits[i], hits[j]
        torch.cuda.current_stream().synchronize()
        t0 = time.perf_counter()
        if pre is not None:
            p = mmx.permutation_test(a, b, index_rows=pre.index_rows(k))
        else:
            p = mmx.permutation_test(a, b, rng=rng)
        torch.cuda.current_stream().synchronize()
        tests["per_query_n50000"].append({"pair": [i, j], "p": float(p), "ms": 1e3 * (time.perf_counter() - t0)})
    if pre is not None:
        pre.commit(rng)                          # the generator continues behind the 21 prefetched tests
    for i, j in pairs:
        torch.cuda.current_stream().synchronize()
        t0 = time.perf_counter()
        p = mmx.permutation_test(rows[i]["_bootstrap_values"][:, 0], rows[j]["_bootstrap_values"][:, 0], rng=rng)
        torch.cuda.current_stream().synchronize()
        tests["per_iteration_n2000"].append({"pair": [i, j], "p": float(p), "ms": 1e3 * (time.perf_counter() - t0)})
    torch.cuda.current_stream().synchronize()
    t_tests = time.perf_counter() - t_tests
    # ---- verification, outside the timed workload: the reference itself (main.py:281-302) with numpy's own shuffles from
    # the stream position the tests started at — same p-value, same generator state behind the test. (Run afterwards: the
    # numpy loop holds the GIL for ~5 s per test and would stall the thread that feeds the prefetch stream.)
    t_check = time.perf_counter()
    g2 = np.random.RandomState()
    g2.set_state(state_at_tests)
    for k in range(min(args.check_host, len(pairs))):
        i, j = pairs[k]
        a, b = hits[i], hits[j]
        obs = np.abs(np.mean(a) - np.mean(b))
        comb = np.concatenate([a, b])
        diffs = []
        for _ in range(10000):
            g2.shuffle(comb)
            diffs.append(np.abs(np.mean(comb[:len(a)]) - np.mean(comb[len(a):])))
        ref_p = float(np.mean(np.array(diffs) >= obs))
        rec = tests["per_query_n50000"][k]
        rec["reference_p"] = ref_p
        rec["same_p"] = bool(ref_p == rec["p"])
        rec["same_generator_state"] = pre.stands_after(g2, k) if pre is not None else None
    t_check = time.perf_counter() - t_check
    out = {"workload": f"configs[4]: ColPali MaxSim {NQ} q x {ND} docs x {LD} patches, doc-sharded x{world}, {args.iterations}-iter "
                       f"bootstrap, {args.models} models, {len(tests['per_iteration_n2000'])} + {len(tests['per_query_n50000'])} permutation tests",
           "n_gpus": world, "shuffles_prefetched_on_rank0_while_ranks_1_to_7_score": prefetch,
           "total_s": t_models + t_tests, "models_s": t_models, "tests_s": t_tests, "reference_check_s (not in total)": t_check,
           "prefetch_trace_s": [[k, a - t_all, b - t_all] for k, a, b in pre.trace] if pre is not None else None,
           "per_model": phases, "peak_mem_GB_per_gpu": float(peak.item()),
           "tests": tests,
           "per_query_test_ms_mean": float(np.mean([r["ms"] for r in tests["per_query_n50000"]])) if tests["per_query_n50000"] else None,
           "per_iteration_test_ms_mean": float(np.mean([r["ms"] for r in tests["per_iteration_n2000"]])) if tests["per_iteration_n2000"] else None,
           "base_R@1": [r["base"][1] for r in rows], "R@1_ci": [[r["R@1_lower"], r["R@1_upper"]] for r in rows]}
    print(json.dumps({k: v for k, v in out.items() if k not in ("tests", "per_model", "prefetch_trace_s")}), flush=True)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(out, indent=1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

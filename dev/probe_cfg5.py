"""BASELINE.json configs[4]: ColPali MaxSim at a 100k-document synthetic corpus (25 000 queries), 1000-iteration
bootstrap + pairwise permutation tests across 7 'models' (7 noise levels) — one GPU, device-resident embeddings."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import mmx
import bench

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
NQ, ND, LQ, LD = 25000, 100000, 20, 1030
out = {"workload": "ColPali MaxSim 25000 queries x 100000 docs x 1030 patches, 1000-iter bootstrap, 7 models, 21 permutation tests"}
rows, t_models = {}, []
for mi, sigma in enumerate((3.0, 3.2, 3.4, 3.6, 3.8, 4.0, 4.2)):
    Q, D, q2d = bench.synth_maxsim(dev, NQ, ND, LQ, LD, seed=2, sigma_q=sigma)
    qv = mmx.FlatMultiVector(Q, np.full(NQ, LQ)); dv = mmx.FlatMultiVector(D, np.full(ND, LD))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tm = {}
    r = mmx.run_retrieval_bootstrap(qv, dv, q2d.cpu().numpy(), {"type": "colpali", "name": f"m{mi}"}, 1000,
                                    rng=np.random.RandomState(42 + mi), timings=tm)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t_models.append(dt)
    rows[f"m{mi}"] = r
    print(f"model {mi} sigma={sigma}: {dt:.2f} s (score {tm['score_ms']/1e3:.2f} s)  R@1 {r['base'][1]:.2f}  CI [{r['R@1_lower']:.3f}, {r['R@1_upper']:.3f}]", flush=True)
    del Q, D, qv, dv
    torch.cuda.empty_cache()
    if mi == 0:
        out["pairs_per_s_model0"] = NQ * ND / dt
        out["maxsim_tflops_model0"] = NQ * ND * bench.FLOP_PER_PAIR / (tm["score_ms"] * 1e-3) / 1e12
t0 = time.perf_counter()
names = list(rows)
rs = np.random.RandomState(7)
pv = {}
for i in range(len(names)):
    for j in range(i + 1, len(names)):
        pv[f"{names[i]}-{names[j]}"] = float(mmx.permutation_test(rows[names[i]]["_bootstrap_values"][:, 0],
                                                                  rows[names[j]]["_bootstrap_values"][:, 0], rng=rs))
torch.cuda.synchronize()
out["permutation_tests_s"] = time.perf_counter() - t0
out["seconds_per_model"] = t_models
out["p_values_adjacent"] = {k: v for k, v in pv.items() if int(k[1]) + 1 == int(k[4])}
out["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out))

"""BASELINE.json configs[0] in full: OpenAI-CLIP-L shapes (5000 x 768 images, 25 000 x 768 captions), Recall@1/5/10 for
the three protocols + 100-iteration bootstrap. The reference's CPU path (oracle port of main.py:586-651: gather,
rescore with fp32 sgemm, 45 000 torch.topk calls per iteration, bootstrap_confidence_interval, failure analysis) is run
in full on the host cores and compared column by column with the row the B200 path produces from the same embeddings
and the same RandomState(42). Writes one JSON line (kept as profiles/r1_cfg1_parity.json)."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import mmx
import oracle
from oracle.failure import aggregate_failure_analysis
from helpers import synth_dense

N, D, SIGMA, B = 5000, 768, 8.0, int(os.environ.get("CFG1_ITERS", "100"))
img, all_txt = synth_dense(N, D, SIGMA, seed=0)
t2i = all_txt[0::5].contiguous()
caps = [("a red car next to two dogs %d" if i % 3 == 0 else "a person holding a pizza %d") % i for i in range(N)]
m_info = {"type": "dense", "name": "OpenAI-CLIP-L-synth"}
torch.cuda.set_device(0)
mmx.run_bootstrap_from_embeddings(img, t2i, all_txt, m_info, B, rng=np.random.RandomState(1), t2i_captions=caps, encoding_time=10.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
row = mmx.run_bootstrap_from_embeddings(img, t2i, all_txt, m_info, B, rng=np.random.RandomState(42), t2i_captions=caps,
                                        encoding_time=10.0)
torch.cuda.synchronize(); gpu_s = time.perf_counter() - t0

torch.set_num_threads(os.cpu_count() or 1)
rs = np.random.RandomState(42)
t0 = time.perf_counter()
per = oracle.bootstrap_rescore_reference(img, t2i, all_txt, B, rng=rs, t2i_captions=caps)
loop_s = time.perf_counter() - t0
per_iter = {k: np.array([m[k] for m in per]) for k in oracle.METRIC_KEYS}
want = oracle.aggregate_row(m_info["name"], per_iter, rng=rs, encoding_time=10.0, n_images=N,
                            failure_analysis=aggregate_failure_analysis([m["_query_results"] for m in per]))
cpu_s = time.perf_counter() - t0
diff_cols = [k for k in mmx.CSV_COLUMNS if k != "Time" and row[k] != want[k]]
iter_diff = {k: int((row["_bootstrap_values"][k] != per_iter[k]).sum()) for k in oracle.METRIC_KEYS}
print(json.dumps({"workload": f"configs[0]: dense {N}x{D} / {5 * N}x{D}, 3 protocols, {B}-iteration bootstrap, 9 CIs, failure analysis",
                  "gpu_row_s": gpu_s, "cpu_reference_s": cpu_s, "cpu_loop_s": loop_s, "cores": os.cpu_count(),
                  "csv_columns_compared": len(mmx.CSV_COLUMNS) - 1, "columns_that_differ": diff_cols,
                  "per_iteration_values_that_differ": iter_diff,
                  "T2I_R@1": [row["T2I_R@1_mean"], row["T2I_R@1_lower"], row["T2I_R@1_upper"]],
                  "speedup": cpu_s / gpu_s}))

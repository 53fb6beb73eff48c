"""One process, two GPUs in turn: kernel attributes (dynamic shared memory) and the RNG tables are per device."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import mmx, oracle
from helpers import synth_colpali, synth_dense
docs, queries = synth_colpali(12, 60, doc_len=300, q_len=(8, 30), seed=1)
img, all_txt = synth_dense(300, 128, 3.0, seed=2)
want = oracle.maxsim_scores(queries, docs)
for d in range(torch.cuda.device_count()):
    torch.cuda.set_device(d)
    got = mmx.compute_similarity_scores(docs, queries, {"type": "colpali"}, None)
    sw = mmx.compute_similarity_scores(queries, docs, {"type": "colpali"}, None)
    row = mmx.run_bootstrap_from_embeddings(img, all_txt[0::5].contiguous(), all_txt, {"type": "dense", "name": "x"}, 50,
                                            rng=np.random.RandomState(3))
    p = mmx.permutation_test(row["_bootstrap_values"]["T2I_R@1"], row["_bootstrap_values"]["I2T_Sym_R@1"], 300,
                             rng=np.random.RandomState(4))
    rs = np.random.RandomState(5); dr = mmx.ops.DeviceRng(rs); x = dr.randint(1000, 400000).cpu().numpy(); dr.commit()
    ok = (float(((got - want).abs() / want.abs().clamp_min(1e-3)).max()) < 1e-3
          and np.array_equal(x, np.random.RandomState(5).randint(0, 1000, size=400000)))
    print(f"cuda:{d} ok={ok} T2I R@1 {row['T2I_R@1_mean']:.4f} p={p}", flush=True)

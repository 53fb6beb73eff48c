"""torchrun check (N >= 2, NCCL): mmx.run_retrieval_bootstrap on document/gallery shards == the same call on one GPU.
Rank 0 writes the single-GPU result first (before init_process_group), then every rank runs the sharded call."""
import os, sys, pickle
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import torch.distributed as dist
import mmx
from helpers import synth_colpali, synth_dense

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
docs, queries = synth_colpali(333, 333 * 5, q_len=(6, 14), sigma_q=2.5, seed=12, ragged_docs=(40, 90))
docs[17] = -docs[17].abs(); queries[85] = queries[85].abs()      # a padded document whose zero row wins a maximum
q2d = np.arange(333 * 5) // 5
img, all_txt = synth_dense(2000, 256, 4.0, seed=3)
cases = {"colpali": (queries, docs, q2d, {"type": "colpali", "name": "c"}, {}),
         "dense": (all_txt, img, np.arange(10000) // 5, {"type": "dense", "name": "d"}, {}),
         # LIST mode (no bitmask) with a starved capacity: the sharded exact recount must give the bitmask result
         "dense_lists_cap3": (all_txt, img, np.arange(10000) // 5, {"type": "dense", "name": "d"}, {"list_mode": True, "list_cap": 3}),
         "dense_lists_cap64": (all_txt, img, np.arange(10000) // 5, {"type": "dense", "name": "d"}, {"list_mode": True})}
ref_path = Path(os.environ.get("TMPDIR", "/tmp")) / "mmx_sharded_ref.pkl"
if rank == 0:
    ref = {k: mmx.run_retrieval_bootstrap(q, d, g, mi, 200, rng=np.random.RandomState(9)) for k, (q, d, g, mi, _) in cases.items()}
    ref_path.write_bytes(pickle.dumps(ref))
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
for k, (q, d, g, mi, kw) in cases.items():
    a, b = mmx.sharded.shard_bounds(len(d), world)[rank]
    shard = d[a:b]
    res = mmx.run_retrieval_bootstrap(q, shard, g, mi, 200, rng=np.random.RandomState(9), **kw)
    if rank == 0:
        ref = pickle.loads(ref_path.read_bytes())[k]
        if kw:
            print(k, res["_list_mode"], flush=True)
        for key in ref:
            same = np.array_equal(ref[key], res[key]) if isinstance(ref[key], np.ndarray) else ref[key] == res[key]
            if not same:
                ok = False
                print("MISMATCH", k, key, ref[key], res[key])
        print(k, "sharded x%d == single GPU:" % world, ok, {kk: res[kk] for kk in ("R@1_mean", "R@1_lower", "R@1_upper")}, flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

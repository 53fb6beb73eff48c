"""CTA-pair dense kernels (tcgen05 cta_group::2, MMX_DENSE_PAIR=2: pairs whenever possible) against the one-CTA kernels
(MMX_DENSE_PAIR=0): thresholds, counts, bitmasks, scores and top-k lists must be bit-identical; LIST-mode lists are
compared as sets (their order is the arrival order of atomics). Each mode runs in a child process with a time limit.

    python dev/check_dense_pair.py
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
CASES = [(300, 700, 64), (5000, 25000, 1152), (1000, 1283, 768), (40000, 40000, 1024)]   # nq, ng, d

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch
    from mmx import ops
    idx = int(sys.argv[2])
    nq, ng, d = CASES[idx]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(idx)
    G = torch.randn(ng, d, generator=g, device=dev)
    G = (G / G.norm(dim=1, keepdim=True)).bfloat16()
    gt = torch.randint(0, ng, (nq,), generator=g, device=dev)
    Q = G[gt].float() + 2.0 * torch.randn(nq, d, generator=g, device=dev) / d ** 0.5
    Q = (Q / Q.norm(dim=1, keepdim=True)).bfloat16()
    gt_off = torch.arange(nq + 1, dtype=torch.int32, device=dev)
    gt_idx = gt.to(torch.int32)
    r = ops.dense_rank(Q, G, gt_off, gt_idx, top_k=10 if nq <= 5000 else None)
    out = {"thr": r.thr, "cnt": r.cnt_above, "bits": r.bits}
    if r.top_s is not None:
        out["top_s"], out["top_i"] = r.top_s, r.top_i
    if nq * ng <= 2e8:
        out["scores"] = ops.dense_scores(Q, G)
    rl = ops.dense_rank_lists(Q, G, gt_off, gt_idx, thr=r.thr)
    out["list_cnt"] = rl.cnt_above
    out["list_n"] = rl.list_n
    out["list_sorted"] = torch.sort(torch.where(torch.arange(rl.list_i.shape[2], device=dev)[None, None, :] <
                                                rl.list_n[:, :, None].clamp_max(rl.list_i.shape[2]), rl.list_i, -1), dim=2)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        ops.dense_rank(Q, G, gt_off, gt_idx, thr=r.thr)
    e0.record()
    for _ in range(5):
        ops.dense_rank(Q, G, gt_off, gt_idx, thr=r.thr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    np.savez(sys.argv[3], **{k: v.cpu().numpy() for k, v in out.items()})
    print(f"case {idx} {nq}x{ng}x{d} pair={os.environ.get('MMX_DENSE_PAIR')}: COUNT pass {ms:.3f} ms, "
          f"{2.0 * nq * ng * d / ms / 1e9:.0f} TFLOP/s", flush=True)
    sys.exit(0)

import numpy as np
ok = True
for i in range(len(CASES)):
    outs = []
    for mode in ("0", "2"):
        path = f"/tmp/dpair_{i}_{mode}.npz"
        env = dict(os.environ, MMX_DENSE_PAIR=mode)
        try:
            r = subprocess.run([sys.executable, __file__, "child", str(i), path], env=env, timeout=150, capture_output=True, text=True)
        except subprocess.TimeoutExpired:
            print(f"case {i} pair={mode}: TIMEOUT (hang)", flush=True)
            ok = False
            break
        print(r.stdout.strip() or r.stderr.strip()[-800:], flush=True)
        if r.returncode != 0:
            ok = False
            break
        outs.append(np.load(path))
    if len(outs) == 2:
        same = {k: bool(np.array_equal(outs[0][k], outs[1][k])) for k in outs[0].files}
        print(f"case {i}: identical: {same}", flush=True)
        ok = ok and all(same.values())
    if not ok:
        break
sys.exit(0 if ok else 1)

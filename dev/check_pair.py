"""CTA-pair MaxSim kernel (tcgen05 cta_group::2, MMX_MAXSIM_PAIR=1) against the one-CTA kernel: the scores must be
bit-identical (every output element is the same K = 128 accumulation; max and token sums are per CTA and unchanged).
Each case runs in a child process per mode (the switch is read once per process) with a time limit: a wrong barrier
protocol hangs, and a hang must not take the box with it.

    python dev/check_pair.py            # parent: runs the children
"""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
CASES = [  # nq, lq, nd, ld
    (700, 20, 90, 1030),
    (4096, 20, 300, 1030),
    (1300, (5, 40), 70, (200, 1500)),
    (25000, 20, 625, 1030),
]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch
    from mmx import ops
    idx = int(sys.argv[2])
    nq, lq, nd, ld = CASES[idx]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(idx)
    rs = np.random.RandomState(idx)
    ql = np.full(nq, lq) if isinstance(lq, int) else rs.randint(lq[0], lq[1] + 1, size=nq)
    dl = np.full(nd, ld) if isinstance(ld, int) else rs.randint(ld[0], ld[1] + 1, size=nd)
    Q = torch.randn(int(ql.sum()), 128, generator=g, device=dev)
    Q = (Q / Q.norm(dim=1, keepdim=True)).bfloat16()
    D = torch.randn(int(dl.sum()), 128, generator=g, device=dev)
    D = (D / D.norm(dim=1, keepdim=True)).bfloat16()
    q_off = torch.from_numpy(np.concatenate([[0], np.cumsum(ql)]).astype(np.int32)).to(dev)
    d_off = torch.from_numpy(np.concatenate([[0], np.cumsum(dl)]).astype(np.int32)).to(dev)
    S = ops.maxsim_scores(Q, q_off, D, d_off)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.maxsim_scores(Q, q_off, D, d_off, out=S)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flop = 2.0 * 128 * float(ql.sum()) * float(dl.sum())
    np.save(sys.argv[3], S.cpu().numpy())
    print(f"case {idx} pair={os.environ.get('MMX_MAXSIM_PAIR', '0')}: {ms:.3f} ms, {flop / ms / 1e9:.0f} TFLOP/s", flush=True)
    sys.exit(0)

import numpy as np
ok = True
for i in range(len(CASES)):
    outs = []
    for mode in ("0", "1"):
        path = f"/tmp/pair_{i}_{mode}.npy"
        env = dict(os.environ, MMX_MAXSIM_PAIR=mode)
        try:
            r = subprocess.run([sys.executable, __file__, "child", str(i), path], env=env, timeout=90, capture_output=True, text=True)
        except subprocess.TimeoutExpired:
            print(f"case {i} pair={mode}: TIMEOUT (hang)", flush=True)
            ok = False
            break
        print(r.stdout.strip() or r.stderr.strip()[-600:], flush=True)
        if r.returncode != 0:
            ok = False
            break
        outs.append(np.load(path))
    if len(outs) == 2:
        same = np.array_equal(outs[0], outs[1])
        print(f"case {i}: bit-identical = {same}; max |diff| = {np.abs(outs[0] - outs[1]).max():.3e}", flush=True)
        ok = ok and same
    if not ok:
        break
sys.exit(0 if ok else 1)

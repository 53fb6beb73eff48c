"""Dev probe: K1 dense rank GEMM throughput at several sizes (CUDA events), cuBLAS bf16 matmul beside it for context."""
import sys; sys.path.insert(0, "/root/repo")
import torch, mmx
from mmx import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def unit(x): return x / x.norm(dim=-1, keepdim=True)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for nq, ng, d, bits in ((5000, 5000, 1152, True), (5000, 25000, 1152, True), (25000, 25000, 768, True), (100000, 100000, 1024, False), (262144, 262144, 1024, False), (1000000, 1000000, 1024, False)):
    Q = unit(torch.randn(nq, d, generator=g, device=dev)).bfloat16()
    G = unit(torch.randn(ng, d, generator=g, device=dev)).bfloat16()
    gt_off = torch.arange(nq + 1, dtype=torch.int32, device=dev)
    gt_idx = (torch.arange(nq, device=dev) % ng).to(torch.int32)
    it = 5 if nq < 500000 else 1
    t_thr = timeit(lambda: ops.dense_rank(Q, G, gt_off, gt_idx, threshold_only=True), it)
    r = ops.dense_rank(Q, G, gt_off, gt_idx, threshold_only=True)
    t_cnt = timeit(lambda: ops.dense_rank(Q, G, gt_off, gt_idx, want_bits=bits, thr=r.thr), it)
    fl = 2.0 * nq * ng * d
    line = f"nq={nq} ng={ng} d={d} bits={bits}: threshold pass {t_thr:.3f} ms, count pass {t_cnt:.3f} ms = {fl/t_cnt/1e9:.1f} TFLOP/s"
    if nq <= 262144:
        t_mm = timeit(lambda: torch.matmul(Q[:min(nq,65536)], G[:min(ng,65536)].t()), it)
        fl2 = 2.0 * min(nq,65536) * min(ng,65536) * d
        line += f" | cuBLAS bf16-out matmul ({min(nq,65536)}x{min(ng,65536)}): {fl2/t_mm/1e9:.1f} TFLOP/s"
    print(line, flush=True)
    del Q, G

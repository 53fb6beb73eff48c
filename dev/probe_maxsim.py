"""Dev probe: MaxSim kernel vs a torch fp32 restatement on the GPU, then a throughput timing."""
import ctypes, importlib, sys, time
sys.path.insert(0, "/root/repo")
import torch
L = importlib.import_module("multimodal-embedding_b200._lib")
lib = L.lib()
dev = torch.device("cuda:0")

def run(qs, ds, d_zero=None, check=True, iters=0):
    q_off = torch.tensor([0] + [q.shape[0] for q in qs]).cumsum(0).int()
    d_off = torch.tensor([0] + [d.shape[0] for d in ds]).cumsum(0).int()
    Q = torch.cat(qs).to(dev).contiguous(); D = torch.cat(ds).to(dev).contiguous()
    nq, nd = len(qs), len(ds)
    scores = torch.full((nq, nd), float("nan"), device=dev)
    need = ctypes.c_size_t()
    L.check(lib.mmx_maxsim_workspace_bytes(Q.shape[0], nd, D.shape[0], ctypes.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    qo, do = q_off.to(dev), d_off.to(dev)
    dz = d_zero.to(dev) if d_zero is not None else None
    st = torch.cuda.current_stream().cuda_stream
    def call():
        L.check(lib.mmx_maxsim_scores(Q.data_ptr(), qo.data_ptr(), nq, Q.shape[0], D.data_ptr(), do.data_ptr(), nd, D.shape[0], 128,
                                     dz.data_ptr() if dz is not None else None, scores.data_ptr(), nd, ws.data_ptr(), need.value, st))
    call(); torch.cuda.synchronize()
    if check:
        cols = list(range(nd)) if nd <= 400 else list(range(40)) + list(range(nd-40, nd))
        ref = torch.zeros(nq, nd, device=dev)
        for j in cols:
            a, b = int(d_off[j]), int(d_off[j+1])
            if b > a:
                mx = (Q.float() @ D[a:b].float().T).max(1).values
                if d_zero is not None and d_zero[j]: mx = mx.clamp_min(0)
            else:
                mx = torch.zeros(Q.shape[0], device=dev)
            cs = torch.cat([torch.zeros(1, device=dev, dtype=torch.float64), mx.double().cumsum(0)])
            ref[:, j] = (cs[q_off[1:].long().to(dev)] - cs[q_off[:-1].long().to(dev)]).float()
        sc, rf = scores[:, cols], ref[:, cols]
        err = (sc - rf).abs().max().item(); rel = ((sc-rf).abs()/rf.abs().clamp_min(1e-6)).max().item()
        print(f"nq={nq} nd={nd} TQ={Q.shape[0]} TD={D.shape[0]} max_abs_err={err:.3e} max_rel={rel:.3e} nan={torch.isnan(scores).sum().item()}")
        assert err < 1e-3, "MISMATCH"
    if iters:
        for _ in range(3): call()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters): call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        pairs = nq * nd
        flops = 2.0 * Q.shape[0] * D.shape[0] * 128
        print(f"  {ms:.3f} ms/iter  {pairs/ms*1e3:.3e} pairs/s  {flops/ms/1e9:.1f} TFLOP/s")

def unit(x): return x / x.norm(dim=-1, keepdim=True)
g = torch.Generator().manual_seed(0)
def rnd(n): return unit(torch.randn(n, 128, generator=g)).bfloat16()

# 1. tiny: single query, single doc
run([rnd(20)], [rnd(1030)])
# 2. ragged both sides, boundaries everywhere
qs = [rnd(int(l)) for l in torch.randint(1, 41, (57,), generator=g)]
ds = [rnd(int(l)) for l in torch.randint(1, 700, (23,), generator=g)] + [rnd(1030), rnd(1), rnd(256), rnd(512), rnd(3)]
dz = (torch.rand(len(ds), generator=g) < 0.5).to(torch.uint8)
run(qs, ds); run(qs, ds, dz)
# 3. long queries (role-swapped orientation): images as queries, captions as docs
run([rnd(1030) for _ in range(5)], [rnd(int(l)) for l in torch.randint(5, 40, (200,), generator=g)], torch.ones(200, dtype=torch.uint8))
# 4. throughput: 2560 queries x 20 tokens vs 1000 docs x 1030
run([rnd(20) for _ in range(2560)], [rnd(1030) for _ in range(1000)], check=True, iters=5)
gg = torch.Generator(device=dev).manual_seed(1)
def grnd(n, L): 
    x = torch.randn(n * L, 128, generator=gg, device=dev); x = (x / x.norm(dim=-1, keepdim=True)).bfloat16()
    return list(x.split(L))
run(grnd(25000, 20), grnd(5000, 1030), check=True, iters=3)
print("PROBE OK")

import ctypes, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from bench import synth_maxsim, NQ, ND, LQ, LD
lib = ctypes.CDLL("/root/repo/dev/libmmx_prof.so")
dev = torch.device("cuda:0")
Q, D, q2d = synth_maxsim(dev, NQ, ND, LQ, LD)
q_off = torch.arange(0, NQ*LQ+1, LQ, dtype=torch.int32, device=dev); d_off = torch.arange(0, ND*LD+1, LD, dtype=torch.int32, device=dev)
S = torch.empty(NQ, ND, device=dev)
need = ctypes.c_size_t(); lib.mmx_maxsim_workspace_bytes(ctypes.c_int64(Q.shape[0]), ctypes.c_int64(ND), ctypes.c_int64(D.shape[0]), ctypes.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
def call():
    rc = lib.mmx_maxsim_scores(ctypes.c_void_p(Q.data_ptr()), ctypes.c_void_p(q_off.data_ptr()), ctypes.c_int64(NQ), ctypes.c_int64(Q.shape[0]), ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(d_off.data_ptr()), ctypes.c_int64(ND), ctypes.c_int64(D.shape[0]), 128, None, ctypes.c_void_p(S.data_ptr()), ctypes.c_int64(ND), ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(need.value), None)
    assert rc == 0
call(); torch.cuda.synchronize(); call(); torch.cuda.synchronize()
h = np.zeros(148*16*8, dtype=np.uint64); lib.mmxp_get_prof(h.ctypes.data_as(ctypes.c_void_p))
h = h.reshape(148, 16, 8).astype(np.float64)
mma = h[:, 1, :]; epi = h[:, 4:12, :]
tot = mma[:, 6].mean()
print("MMA thread: total cycles %.3e  wait b_full %.1f%%  wait t_empty0 %.1f%%  wait t_empty1 %.1f%%" % (tot, 100*mma[:,0].mean()/tot, 100*mma[:,1].mean()/tot, 100*mma[:,2].mean()/tot))
et = epi[:, :, 6].mean()
print("EPI warps: total %.3e  wait t_full %.1f%%  interior(to release) %.1f%%  interior(total) %.1f%%  boundary %.1f%%" % (et, 100*epi[:,:,0].mean()/et, 100*epi[:,:,1].mean()/et, 100*epi[:,:,2].mean()/et, 100*epi[:,:,3].mean()/et))
print("per interior drain: to-release %.0f cyc, total %.0f cyc ; per boundary drain %.0f cyc ; counts %d / %d" % (epi[:,:,1].sum()/epi[:,:,4].sum(), epi[:,:,2].sum()/epi[:,:,4].sum(), epi[:,:,3].sum()/max(1,epi[:,:,5].sum()), epi[0,0,4], epi[0,0,5]))
print("avg wait t_full per drain %.0f cyc" % (epi[:,:,0].sum()/(epi[:,:,4].sum()+epi[:,:,5].sum())))

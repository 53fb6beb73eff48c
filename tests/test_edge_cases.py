"""Edge cases of the C-ABI entry points on the GPU: empty and one-element inputs, sizes around tile boundaries,
argument errors reported as exceptions (the reference's caller catches Exception per model, main.py:839)."""
import numpy as np
import pytest
import torch

import oracle
from helpers import unit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mmx_mod():
    import mmx
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return mmx


def _rnd(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return unit(torch.randn(n, d, generator=g)).bfloat16()


def test_empty_inputs(mmx_mod):
    dev = torch.device("cuda:0")
    ops = mmx_mod.ops
    e128 = torch.empty(0, 128, dtype=torch.bfloat16, device=dev)
    off0 = torch.zeros(1, dtype=torch.int32, device=dev)
    docs = _rnd(40, 128, 1).to(dev)
    d_off = torch.tensor([0, 10, 40], dtype=torch.int32, device=dev)
    assert tuple(ops.maxsim_scores(e128, off0, docs, d_off).shape) == (0, 2)            # no queries
    assert tuple(ops.maxsim_scores(docs, d_off, e128, off0).shape) == (2, 0)            # no documents
    # a query without tokens scores 0; a document without patches scores 0 (the reference would pad it with zeros)
    q_off = torch.tensor([0, 0, 10, 40], dtype=torch.int32, device=dev)
    dd_off = torch.tensor([0, 40, 40], dtype=torch.int32, device=dev)
    s = ops.maxsim_scores(docs, q_off, docs, dd_off)
    assert torch.all(s[0] == 0) and torch.all(s[:, 1] == 0) and torch.all(s[1:, 0] > 0)
    # the same with long documents (the T2I kernel variant): an empty document among 300-row ones scores 0
    long_docs = _rnd(900, 128, 3).to(dev)
    ld_off = torch.tensor([0, 300, 300, 600, 900], dtype=torch.int32, device=dev)
    s2 = ops.maxsim_scores(docs, d_off, long_docs, ld_off)
    assert torch.all(s2[:, 1] == 0) and torch.all(torch.isfinite(s2)) and torch.all(s2[:, 0] > 0)
    Q = _rnd(5, 64, 2).to(dev)
    gt_off = torch.arange(6, dtype=torch.int32, device=dev)
    r = ops.dense_rank(Q, torch.empty(0, 64, dtype=torch.bfloat16, device=dev), gt_off, torch.zeros(5, dtype=torch.int32, device=dev))
    assert torch.all(r.cnt_above == 0) and torch.all(torch.isinf(r.thr))
    ts, ti = ops.dense_topk(Q, torch.empty(0, 64, dtype=torch.bfloat16, device=dev), 10)
    assert torch.all(ti == -1) and torch.all(torch.isinf(ts))
    ts, ti = ops.topk_from_scores(torch.empty(3, 0, dtype=torch.float32, device=dev), 10)
    assert tuple(ti.shape) == (3, 10) and torch.all(ti == -1)
    # LIST mode on an empty gallery: nothing is above anything, every drawn query copy is a hit (min(m_q, k) of them)
    rl = ops.dense_rank_lists(Q, torch.empty(0, 64, dtype=torch.bfloat16, device=dev), gt_off,
                              torch.zeros(5, dtype=torch.int32, device=dev))
    assert torch.all(rl.cnt_above == 0) and torch.all(rl.list_n == 0)
    mult = torch.tensor([[2, 0, 1, 1, 0], [0, 0, 0, 3, 1]], dtype=torch.uint8, device=dev)
    correct = ops.boot_counts_lists(rl, mult, resolve=None)
    assert correct.cpu().tolist() == [[3, 4, 4], [2, 4, 4]]


@pytest.mark.parametrize("nq,ng,d", [(1, 1, 8), (1, 300, 16), (255, 257, 64), (256, 256, 72), (257, 255, 1152), (513, 1, 128)])
def test_dense_sizes_around_tile_boundaries(mmx_mod, nq, ng, d):
    dev = torch.device("cuda:0")
    Q, G = _rnd(nq, d, nq), _rnd(ng, d, ng + 7)
    S = mmx_mod.ops.dense_scores(Q.to(dev), G.to(dev)).cpu()
    want = oracle.dense_scores(G, Q)
    assert float((S - want).abs().max()) < 1e-3
    gt = [[(3 * i) % ng] for i in range(nq)]
    off, idx = mmx_mod.ops.csr_from_sets(gt)
    r = mmx_mod.ops.dense_rank(Q.to(dev), G.to(dev), torch.from_numpy(off).to(dev), torch.from_numpy(idx).to(dev))
    thr = S[torch.arange(nq), torch.tensor([g[0] for g in gt])]
    assert torch.equal(r.thr.cpu(), thr)                                   # same MMA path: bit-identical
    assert torch.equal(r.cnt_above.cpu(), (S > thr[:, None]).sum(1).int())
    ts, ti = mmx_mod.ops.dense_topk(Q.to(dev), G.to(dev), 10)
    wv, wi = oracle.topk_lists(S, 10)
    assert torch.equal(ts.cpu(), wv) and torch.equal(ti.cpu().long(), wi)
    # LIST mode around the same tile boundaries: exact counts, and the lists hold min(count, cap) distinct set columns
    rl = mmx_mod.ops.dense_rank_lists(Q.to(dev), G.to(dev), torch.from_numpy(off).to(dev), torch.from_numpy(idx).to(dev), cap=8)
    assert torch.equal(rl.cnt_above, r.cnt_above)
    above = (S > thr[:, None]).numpy()
    li, ln = rl.list_i[0].cpu().numpy(), rl.list_n[0].cpu().numpy()
    for q in range(nq):
        n = min(int(ln[q]), 8)
        assert n == min(int(above[q].sum()), 8) and above[q, li[q, :n]].all() and len(set(li[q, :n].tolist())) == n


def test_argument_errors_are_exceptions(mmx_mod):
    dev = torch.device("cuda:0")
    ops = mmx_mod.ops
    Q = torch.zeros(4, 12, dtype=torch.bfloat16, device=dev)               # d % 8 != 0
    off = torch.arange(5, dtype=torch.int32, device=dev)
    with pytest.raises(mmx_mod.MmxError):
        ops.dense_scores(Q, Q)
    with pytest.raises(TypeError):
        ops.dense_scores(Q.float(), Q.float())
    with pytest.raises(mmx_mod.MmxError):
        ops.topk_from_scores(torch.zeros(2, 50, device=dev), 33)           # k > 32
    with pytest.raises(mmx_mod.MmxError):
        ops.maxsim_scores(torch.zeros(4, 128, dtype=torch.bfloat16), off, Q, off)     # host tensor
    with pytest.raises(ValueError):
        mmx_mod.compute_similarity_scores(torch.zeros(3, 128, dtype=torch.bfloat16), [], {"type": "colpali"}, None)
    with pytest.raises(mmx_mod.MmxError):
        mmx_mod.ops.DeviceRng(np.random.RandomState(0)).shuffle_indices(70000, 3)     # beyond the device-shuffle limit (65 536)
    # the library keeps working after an error
    assert float(ops.dense_scores(_rnd(3, 64, 1).to(dev), _rnd(3, 64, 1).to(dev)).diagonal().min()) > 0.99


@pytest.mark.parametrize("seed", [0, 2])
def test_maxsim_randomised_shapes(seed):
    """dev/fuzz_maxsim.py: random ragged shapes (empty items, padding flags, both kernel variants) vs a torch fp32
    restatement; seed 2 holds the empty-document-among-long-documents case that once produced -inf."""
    import subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, str(root / "dev" / "fuzz_maxsim.py"), str(seed), "40"], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "random cases ok" in res.stdout, res.stdout[-500:] + res.stderr[-1500:]


def test_rank_and_statistics_randomised_shapes():
    """dev/fuzz_rank_stats.py: random problems for K1 / KR / KT / tie report / K3 / K4 / K6 / K6s, every integer and
    every double compared bit for bit with torch / numpy restatements."""
    import subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, str(root / "dev" / "fuzz_rank_stats.py"), "3", "12"], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "random cases ok" in res.stdout, res.stdout[-500:] + res.stderr[-1500:]

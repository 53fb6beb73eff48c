"""GPU parity tests: the CUDA path (through the C ABI / reference-signature API) against the CPU oracle and the
golden vectors produced by the reference's own functions. Integers must be bit-exact; fp32 scores within 1e-3
relative (BASELINE.json north_star); documented tie classes are the only permitted rank flips."""
import numpy as np
import pytest
import torch

import oracle
from helpers import from_bits, synth_colpali, synth_dense

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3   # north_star: scores within 1e-3 relative in fp32-accumulate


@pytest.fixture(scope="module")
def mmx_mod():
    import mmx
    assert torch.cuda.is_available(), "GPU tests need a B200"
    mmx._lib.check(mmx._lib.lib().mmx_device_check())
    return mmx


def _dev():
    return torch.device("cuda:0")


def _assert_close_scores(got: torch.Tensor, want: torch.Tensor):
    got, want = got.double().cpu(), want.double().cpu()
    rel = (got - want).abs() / want.abs().clamp_min(1e-3)
    assert float(rel.max()) < REL_TOL, f"max rel err {float(rel.max()):.3e}"


def _rank_parity(cnt_gpu: np.ndarray, scores_cpu: torch.Tensor, gt_sets, max_flips_frac=0.01):
    """cnt must equal the oracle's strict-greater count computed on CPU scores, except for queries where some
    competitor lies within REL_TOL of the threshold (tie class T3: summation-order flips)."""
    thr, cnt, _ = oracle.above_gt_counts(scores_cpu, gt_sets)
    cnt = cnt.numpy()
    diff = np.nonzero(cnt != cnt_gpu)[0]
    for i in diff:
        gap = (scores_cpu[i] - thr[i]).abs()
        g = torch.as_tensor(list(gt_sets[i]))
        gap[g] = float("inf")
        near = int((gap <= REL_TOL * max(abs(float(thr[i])), 1e-3)).sum())
        assert abs(int(cnt[i]) - int(cnt_gpu[i])) <= near, f"query {i}: cnt {cnt_gpu[i]} vs {cnt[i]} with {near} near ties"
    assert len(diff) <= max_flips_frac * len(cnt) + 1
    return len(diff)


# ----------------------------------------------------------------------------------------------- MaxSim (K2)
def test_maxsim_golden_reference_vectors(mmx_mod, golden_dir):
    z = np.load(golden_dir / "maxsim.npz")
    docs = list(from_bits(z["docs"], 128).split([int(l) for l in z["doc_lens"]]))
    queries = list(from_bits(z["queries"], 128).split([int(l) for l in z["query_lens"]]))
    m_info = {"type": "colpali", "name": "golden"}
    s = mmx_mod.compute_similarity_scores(docs, queries, m_info, None)
    assert s.dtype == torch.float32 and not s.is_cuda and tuple(s.shape) == (len(queries), len(docs))
    _assert_close_scores(s, torch.from_numpy(z["scores"]))
    s2 = mmx_mod.compute_similarity_scores(queries, docs, m_info, None)       # roles swapped (v28 I2T)
    _assert_close_scores(s2, torch.from_numpy(z["scores_swapped"]))
    assert float(s[2, 4]) == 0.0          # all-negative pair: the chunk's zero padding wins every max, exactly 0


def test_maxsim_known_answer(mmx_mod):
    """Hand-computed 2x2 case of tests/test_oracle_golden.py::test_maxsim_known_answer, embedded in 128 dims."""
    def emb(rows):
        t = torch.zeros(len(rows), 128)
        t[:, :2] = torch.tensor(rows)
        return t.bfloat16()
    q0, q1 = emb([[1.0, 0.0], [0.0, 1.0]]), emb([[-1.0, 0.0]])
    d0, d1 = emb([[1.0, 0.0], [0.5, 0.5], [0.0, 2.0]]), emb([[1.0, 1.0]])
    s = mmx_mod.compute_similarity_scores([d0, d1], [q0, q1], {"type": "colpali"}, None)
    assert s.tolist() == [[3.0, 2.0], [0.0, 0.0]]


@pytest.mark.parametrize("case", ["t2i_fixed", "ragged_both", "swapped", "single", "many_tiles", "swapped_many",
                                  "short_docs_edge"])
def test_maxsim_vs_oracle(mmx_mod, case):
    if case == "t2i_fixed":
        docs, queries = synth_colpali(24, 120, doc_len=1030, q_len=(20, 20))
    elif case == "ragged_both":
        docs, queries = synth_colpali(37, 111, q_len=(1, 40), ragged_docs=(1, 700), seed=5)
    elif case == "swapped":
        d, q = synth_colpali(6, 60, doc_len=1030, q_len=(5, 40), seed=7)
        docs, queries = q, d            # images are the MaxSim queries, ragged captions the documents
    elif case == "single":
        docs, queries = synth_colpali(1, 1, doc_len=3, q_len=(1, 1), seed=9)
    elif case == "swapped_many":
        # SHORT-document kernel variant over several row-tile pairs and many 256-column tiles; ragged captions make
        # the reference's zero padding take part in the maxima (d_zero flags)
        d, q = synth_colpali(40, 1500, doc_len=1030, q_len=(8, 40), seed=13)
        docs, queries = q, d
    elif case == "short_docs_edge":
        # SHORT variant corner cases: 1-row and 3-row documents (padded by repetition), one long document among short
        # ones, short queries whose segments split 32-row blocks, all-negative similarities against a padded document
        g = torch.Generator().manual_seed(17)
        def rnd(n):
            x = torch.randn(n, 128, generator=g)
            return (x / x.norm(dim=-1, keepdim=True)).bfloat16()
        docs = [rnd(int(l)) for l in torch.randint(1, 30, (90,), generator=g)] + [rnd(1), rnd(3), rnd(700), rnd(16), rnd(17)]
        docs[7] = -docs[7].abs()
        queries = [rnd(int(l)) for l in torch.randint(1, 50, (37,), generator=g)] + [rnd(300), rnd(129)]
        queries[3] = queries[3].abs()
    else:
        docs, queries = synth_colpali(300, 700, q_len=(8, 40), ragged_docs=(200, 300), seed=11)
    want = oracle.maxsim_scores(queries, docs)
    got = mmx_mod.compute_similarity_scores(docs, queries, {"type": "colpali"}, None)
    _assert_close_scores(got, want)


def test_maxsim_query_blocking_equals_single_launch(mmx_mod, monkeypatch):
    """Problems whose carry workspace would exceed the limit are scored in blocks of whole queries: same scores."""
    docs, queries = synth_colpali(40, 90, doc_len=64, q_len=(1, 300), sigma_q=2.0, seed=21)
    m_info = {"type": "colpali"}
    whole = mmx_mod.compute_similarity_scores(docs, queries, m_info, None)
    monkeypatch.setattr(mmx_mod.ops, "MAXSIM_WORKSPACE_LIMIT", 1 << 16)       # forces blocks of a few hundred rows
    blocked = mmx_mod.compute_similarity_scores(docs, queries, m_info, None)
    assert torch.equal(whole, blocked)
    _assert_close_scores(blocked, oracle.maxsim_scores(queries, docs))


def test_maxsim_padded_tensor_input_and_errors(mmx_mod):
    docs, queries = synth_colpali(12, 12, doc_len=40, q_len=(6, 6), seed=3)
    want = oracle.maxsim_scores(queries, docs)
    got = mmx_mod.compute_similarity_scores(torch.stack(docs), torch.stack(queries), {"type": "colpali"}, None)
    _assert_close_scores(got, want)
    with pytest.raises(TypeError):
        mmx_mod.compute_similarity_scores([d.float() for d in docs], queries, {"type": "colpali"}, None)
    with pytest.raises(mmx_mod.MmxError):
        q = torch.zeros(4, 64, dtype=torch.bfloat16, device="cuda")
        off = torch.tensor([0, 4], dtype=torch.int32, device="cuda")
        mmx_mod.ops.maxsim_scores(q, off, q, off)          # dim != 128


def test_maxsim_full_size_properties(mmx_mod):
    """cfg4 shapes (25 000 x ~20 tokens vs 5 000 x 1030): size-independent checks — a doc/query sub-block rescored
    on its own must agree, and scores are invariant to where a doc sits in the stream."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(1)
    nq, nd, Lq, Ld = 25000, 5000, 20, 1030
    Q = torch.randn(nq * Lq, 128, generator=g, device=dev)
    Q = (Q / Q.norm(dim=-1, keepdim=True)).bfloat16()
    D = torch.randn(nd * Ld, 128, generator=g, device=dev)
    D = (D / D.norm(dim=-1, keepdim=True)).bfloat16()
    q_off = torch.arange(0, nq * Lq + 1, Lq, dtype=torch.int32, device=dev)
    d_off = torch.arange(0, nd * Ld + 1, Ld, dtype=torch.int32, device=dev)
    S = mmx_mod.ops.maxsim_scores(Q, q_off, D, d_off)
    assert bool(torch.isfinite(S).all())
    # sub-block [q 1000:1137] x [d 4000:4011] rescored alone (different tile alignment)
    qs, qe, ds, de = 1000, 1137, 4000, 4011
    sub = mmx_mod.ops.maxsim_scores(Q[qs * Lq:qe * Lq], q_off[: qe - qs + 1].clone(), D[ds * Ld:de * Ld],
                                    d_off[: de - ds + 1].clone())
    assert float((S[qs:qe, ds:de] - sub).abs().max()) < 1e-4
    # the same sub-block against the CPU oracle
    want = oracle.maxsim_scores(list(Q[qs * Lq:qe * Lq].cpu().split(Lq)), list(D[ds * Ld:de * Ld].cpu().split(Ld)))
    _assert_close_scores(sub, want)
    # reversing the document order permutes the columns and nothing else
    Drev = D.view(nd, Ld, 128).flip(0).reshape(-1, 128).contiguous()
    Srev = mmx_mod.ops.maxsim_scores(Q[: 200 * Lq], q_off[:201].clone(), Drev, d_off)
    assert float((Srev.flip(1) - S[:200]).abs().max()) < 1e-4


def test_maxsim_i2t_full_size_properties(mmx_mod):
    """The role-swapped I2T orientation at full size (v28 :375-386 / main.py:622): 5 000 images x 1030 patches as
    MaxSim queries against 25 000 ragged captions as documents (SHORT-document kernel variant, padding flags on).
    Size-independent checks: a sub-block rescored alone agrees, and so does the CPU oracle on that sub-block."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(4)
    n_img, n_cap, Lp = 5000, 25000, 1030
    lens = torch.randint(8, 41, (n_cap,), generator=torch.Generator().manual_seed(5))
    P = torch.randn(n_img * Lp, 128, generator=g, device=dev)
    P = (P / P.norm(dim=-1, keepdim=True)).bfloat16()
    T = torch.randn(int(lens.sum()), 128, generator=g, device=dev)
    T = (T / T.norm(dim=-1, keepdim=True)).bfloat16()
    p_off = torch.arange(0, n_img * Lp + 1, Lp, dtype=torch.int32, device=dev)
    t_off_h = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    t_off = t_off_h.to(torch.int32).to(dev)
    flags = torch.from_numpy(oracle.maxsim_pad_flags(lens.numpy())).to(dev)
    S = mmx_mod.ops.maxsim_scores(P, p_off, T, t_off, flags)
    assert bool(torch.isfinite(S).all())
    i0, i1, c0, c1 = 3000, 3003, 12000, 12050                      # c0 is a multiple of the reference's chunk of 10
    a, b = int(t_off_h[c0]), int(t_off_h[c1])
    sub_off = (t_off_h[c0:c1 + 1] - a).to(torch.int32).to(dev)
    sub = mmx_mod.ops.maxsim_scores(P[i0 * Lp:i1 * Lp], p_off[: i1 - i0 + 1].clone(), T[a:b], sub_off, flags[c0:c1].clone())
    assert float((S[i0:i1, c0:c1] - sub).abs().max()) < 1e-3
    want = oracle.maxsim_scores(list(P[i0 * Lp:i1 * Lp].cpu().split(Lp)),
                                list(T[a:b].cpu().split([int(x) for x in lens[c0:c1]])))
    _assert_close_scores(sub, want)


# ----------------------------------------------------------------------------------------------- dense (K1)
def test_dense_scores_golden(mmx_mod, golden_dir):
    z = np.load(golden_dir / "dense.npz")
    img, all_txt = from_bits(z["img"], 64), from_bits(z["all_txt"], 64)
    s = mmx_mod.compute_similarity_scores(img, all_txt[0::5].contiguous(), {"type": "dense"}, None)
    _assert_close_scores(s, torch.from_numpy(z["scores_t2i"]))
    s2 = mmx_mod.compute_similarity_scores(all_txt, img, {"type": "siglip"}, None)
    _assert_close_scores(s2, torch.from_numpy(z["scores_i2t"]))
    with pytest.raises(TypeError):
        mmx_mod.compute_similarity_scores(img.half(), img.half(), {"type": "dense"}, None)


@pytest.mark.parametrize("n,dim,sigma", [(60, 64, 2.5), (333, 768, 8.0), (1000, 1152, 10.0), (257, 1024, 9.0), (129, 72, 3.0)])
def test_dense_rank_vs_oracle(mmx_mod, n, dim, sigma):
    img, all_txt = synth_dense(n, dim, sigma, seed=n)
    t2i = all_txt[0::5].contiguous()
    ranks = mmx_mod.score_and_rank(img, t2i, all_txt, {"type": "dense", "name": "x"})
    s_t2i, s_i2t = oracle.dense_scores(img, t2i), oracle.dense_scores(all_txt, img)
    gts = {"T2I": [[i] for i in range(n)], "I2T": [list(range(5 * i, 5 * i + 5)) for i in range(n)],
           "I2T_Sym": [[i] for i in range(n)]}
    mats = {"T2I": s_t2i, "I2T": s_i2t, "I2T_Sym": s_t2i.t().contiguous()}
    for name in ("T2I", "I2T", "I2T_Sym"):
        r = ranks[name]
        thr, _, _ = oracle.above_gt_counts(mats[name], gts[name])
        _assert_close_scores(r.thr, thr)
        _rank_parity(r.cnt_above.cpu().numpy(), mats[name], gts[name])
        # the bitmask is self-consistent with the counts and never contains a GT column
        bits = r.bits.cpu().numpy().view(np.uint32)
        pop = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little").sum(axis=1)
        assert np.array_equal(pop, r.cnt_above.cpu().numpy())
        unpacked = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, : mats[name].shape[1]]
        for i in (0, n // 2, n - 1):
            assert not unpacked[i, gts[name][i]].any()
    # K1's own score output equals its ranking decisions exactly (same MMA path, thresholds bit-consistent)
    dev = _dev()
    S = mmx_mod.ops.dense_scores(t2i.to(dev), img.to(dev))
    _assert_close_scores(S, s_t2i)
    thr_gpu = S.diagonal()
    assert torch.equal(thr_gpu, ranks["T2I"].thr)
    assert torch.equal((S > thr_gpu[:, None]).sum(1).int(), ranks["T2I"].cnt_above)


def test_dense_arbitrary_gt_and_sharded_gallery(mmx_mod):
    """Arbitrary GT lists (CSR), and the multi-shard formulation: thresholds max-combined over gallery shards,
    counts added, bit columns concatenated == one unsharded pass, bit for bit."""
    dev = _dev()
    n, dim = 700, 256
    img, all_txt = synth_dense(n, dim, 4.0, seed=3)
    Q, G = img.to(dev), all_txt.to(dev)
    rs = np.random.RandomState(0)
    gt_sets = [sorted(rs.choice(5 * n, size=rs.randint(1, 7), replace=False).tolist()) for _ in range(n)]
    off, idx = mmx_mod.ops.csr_from_sets(gt_sets)
    off_d, idx_d = torch.from_numpy(off).to(dev), torch.from_numpy(idx).to(dev)
    full = mmx_mod.ops.dense_rank(Q, G, off_d, idx_d)
    _rank_parity(full.cnt_above.cpu().numpy(), oracle.dense_scores(all_txt, img), gt_sets)
    bounds = [0, 1024, 2048 + 256, 5 * n]
    thr = torch.full((n,), float("-inf"), device=dev)
    for a, b in zip(bounds[:-1], bounds[1:]):
        t = mmx_mod.ops.dense_rank(Q, G[a:b], off_d, idx_d, g_base=a, threshold_only=True).thr
        thr = torch.maximum(thr, t)
    assert torch.equal(thr, full.thr)
    cnt = torch.zeros(n, dtype=torch.int32, device=dev)
    cols = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        r = mmx_mod.ops.dense_rank(Q, G[a:b], off_d, idx_d, g_base=a, thr=thr)
        cnt += r.cnt_above
        cols.append(r.bits[:, : (b - a + 31) // 32])
    assert torch.equal(cnt, full.cnt_above)
    assert torch.equal(torch.cat(cols, dim=1), full.bits)


def test_tie_report_counts_exact_ties_only(mmx_mod):
    """Tie class T2: duplicated gallery rows tie bit for bit with the ground truth; the report counts exactly those
    (dense path without a matrix, and the same through KR on the materialised scores), GT duplicates (T4) excluded."""
    dev = _dev()
    n, dim = 300, 128
    img, all_txt = synth_dense(n, dim, 2.0, seed=21)
    img = img.clone()
    img[10] = img[3]                  # image 10 is a copy of image 3: T2I queries 3 and 10 tie between the two columns
    img[77] = img[3]
    all_txt = all_txt.clone()
    all_txt[26] = all_txt[25]         # two identical GT captions of image 5 (T4): not a non-GT tie
    t2i = all_txt[0::5].contiguous()
    r = mmx_mod.score_and_rank(img, t2i, all_txt, {"type": "dense", "name": "t"}, want_ties=True)
    S = mmx_mod.ops.dense_scores(t2i.to(dev), img.to(dev)).cpu()
    thr = S.diagonal()
    want = ((S == thr[:, None]).sum(1) - 1).int()
    assert torch.equal(r["T2I"].ties.cpu(), want) and int(want[3]) == 2 and int(want[10]) >= 1
    S_i2t = mmx_mod.ops.dense_scores(img.to(dev), all_txt.to(dev)).cpu()
    thr_i = torch.stack([S_i2t[i, 5 * i:5 * i + 5].max() for i in range(n)])
    non_gt = torch.ones_like(S_i2t, dtype=torch.bool)
    for i in range(n):
        non_gt[i, 5 * i:5 * i + 5] = False
    want_i = ((S_i2t == thr_i[:, None]) & non_gt).sum(1).int()
    assert torch.equal(r["I2T"].ties.cpu(), want_i)
    summ = mmx_mod.ops.tie_summary(r["T2I"])
    assert summ["queries_with_exact_ties"] == int((want > 0).sum()) and summ["R@1_depends_on_tie_order"] >= 1
    # KR on the materialised matrix gives the same report
    off, idx = mmx_mod.ops.csr_from_sets([[i] for i in range(n)])
    off_d, idx_d = torch.from_numpy(off).to(dev), torch.from_numpy(idx).to(dev)
    rk = mmx_mod.ops.rank_from_scores(S.to(dev), off_d, idx_d)
    assert torch.equal(mmx_mod.ops.exact_tie_counts(rk, off_d, idx_d, scores=S.to(dev)).cpu(), want)


# ----------------------------------------------------------------------------------------------- metrics (KR)
def test_compute_metrics_detailed_golden(mmx_mod, golden_dir):
    z = np.load(golden_dir / "dense.npz")
    N = z["scores_t2i"].shape[0]
    s_t2i, s_i2t = torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"])
    caps = [str(c) for c in z["captions"]]
    m = mmx_mod.compute_metrics_detailed(s_t2i, s_i2t, s_t2i.t(), list(range(N)),
                                         {i: list(range(5 * i, 5 * i + 5)) for i in range(N)},
                                         {i: i for i in range(N)}, caps)
    assert [k for k in m if not k.startswith("_")] == [str(k) for k in z["metric_keys"]]
    assert [m[k] for k in mmx_mod.METRIC_KEYS] == list(z["base_metrics"])
    assert [int(r["correct_r1"]) for r in m["_query_results"]] == list(z["correct_r1"])
    want = oracle.compute_metrics_detailed(s_t2i, s_i2t, s_t2i.t(), list(range(N)),
                                           {i: list(range(5 * i, 5 * i + 5)) for i in range(N)},
                                           {i: i for i in range(N)}, caps)
    assert m["_query_results"] == want["_query_results"]
    v = np.load(golden_dir / "v28_metrics.npz")
    m28 = mmx_mod.compute_metrics(s_t2i, s_i2t, list(range(N)), {i: list(range(5 * i, 5 * i + 5)) for i in range(N)})
    assert list(m28.keys()) == [str(k) for k in v["keys"]] and list(m28.values()) == list(v["values"])


def test_rank_from_scores_exact_and_edge_cases(mmx_mod):
    dev = _dev()
    rs = np.random.RandomState(1)
    for nq, ng in ((1, 1), (3, 31), (5, 32), (7, 33), (64, 1000), (200, 4097)):
        s = torch.from_numpy(rs.randn(nq, ng).astype(np.float32))
        s[:, ::7] = s[:, :1].clone()                                   # plenty of exact ties, incl. with the GT column
        gt = [sorted(rs.choice(ng, size=min(ng, rs.randint(1, 4)), replace=False).tolist()) for _ in range(nq)]
        off, idx = mmx_mod.ops.csr_from_sets(gt)
        r = mmx_mod.ops.rank_from_scores(s.to(dev), torch.from_numpy(off).to(dev), torch.from_numpy(idx).to(dev))
        thr, cnt, _ = oracle.above_gt_counts(s, gt)
        assert torch.equal(r.thr.cpu(), thr)
        assert np.array_equal(r.cnt_above.cpu().numpy(), cnt.numpy())
        bits = np.unpackbits(r.bits.cpu().numpy().view(np.uint8), axis=1, bitorder="little")[:, :ng]
        assert np.array_equal(bits.astype(bool), (s > thr[:, None]).numpy())
    # k larger than the gallery: everything is correct (main.py:335 min(k, n_img))
    s = torch.tensor([[0.1, 0.9, 0.5]])
    m = mmx_mod.compute_metrics(s, s, [0], {0: [0]})
    assert m["T2I_R@1"] == 0.0 and m["T2I_R@5"] == 100.0 and m["T2I_R@10"] == 100.0


# ----------------------------------------------------------------------------------------------- bootstrap (K3)
def test_bootstrap_golden_reference_iterations(mmx_mod, golden_dir):
    """Per-iteration metrics produced by the reference's own gather+rescore+topk loop (golden) == K3 on the base pass."""
    dev = _dev()
    z = np.load(golden_dir / "dense.npz")
    N = z["scores_t2i"].shape[0]
    img, all_txt = from_bits(z["img"], 64), from_bits(z["all_txt"], 64)
    ranks = mmx_mod.score_and_rank(img, all_txt[0::5].contiguous(), all_txt, {"type": "dense", "name": "g"})
    assert [mmx_mod.base_metrics(ranks)[k] for k in mmx_mod.METRIC_KEYS] == list(z["base_metrics"])
    correct = mmx_mod.bootstrap_from_ranks(ranks, z["sample_indices"], N)
    assert np.array_equal(100.0 * correct / N, z["boot_metrics"])


@pytest.mark.parametrize("list_len", [32, 2])
def test_bootstrap_identity_vs_oracle(mmx_mod, list_len, monkeypatch):
    """N=800: K3 == numpy multiplicity identity == (spot-checked) literal gather+rescore+topk.
    list_len=2 forces the truncated-list exact fallback."""
    monkeypatch.setattr(mmx_mod.ops, "LIST_LEN", list_len)
    import inspect
    n, dim = 800, 128
    img, all_txt = synth_dense(n, dim, 3.5, seed=21)
    t2i = all_txt[0::5].contiguous()
    ranks = mmx_mod.score_and_rank(img, t2i, all_txt, {"type": "dense", "name": "x"})
    rs = np.random.RandomState(42)
    si = rs.randint(0, n, size=(40, n))
    dev = _dev()
    mult = mmx_mod.ops.bincount(torch.from_numpy(si.astype(np.int32)).to(dev), n)
    assert np.array_equal(mult.cpu().numpy(), np.stack([np.bincount(r, minlength=n) for r in si]).astype(np.uint8))
    got = np.concatenate([mmx_mod.ops.boot_counts(ranks[p], mult, list_len=list_len).cpu().numpy()
                          for p in ("T2I", "I2T", "I2T_Sym")], axis=1)
    sets = {}
    for p, col_unit in (("T2I", np.arange(n)), ("I2T", np.arange(5 * n) // 5), ("I2T_Sym", np.arange(n))):
        bits = np.unpackbits(ranks[p].bits.cpu().numpy().view(np.uint8), axis=1, bitorder="little")[:, : len(col_unit)]
        sets[p] = (bits, col_unit, np.arange(n))
    want = oracle.bootstrap_identity(sets, si, n)
    assert np.array_equal(got, want)
    # literal reference loop on 2 iterations, using the GPU's own score matrices so that ties cannot differ
    S_t2i = mmx_mod.ops.dense_scores(t2i.to(dev), img.to(dev)).cpu()
    S_i2t = mmx_mod.ops.dense_scores(img.to(dev), all_txt.to(dev)).cpu()
    for b in range(2):
        s = si[b]
        cap_idx = np.array([c for o in s for c in range(o * 5, o * 5 + 5)])
        m = oracle.compute_metrics_detailed(
            S_t2i[s][:, s], S_i2t[s][:, cap_idx], S_t2i[s][:, s].t(), list(range(n)),
            {i: list(range(5 * i, 5 * i + 5)) for i in range(n)}, {i: i for i in range(n)}, [""] * n)
        assert [m[k] for k in mmx_mod.METRIC_KEYS] == list(100.0 * got[b] / n)


# ----------------------------------------------------------------------------------------------- CI / permutation (K4, K5)
def test_ci_and_permutation_golden(mmx_mod, golden_dir):
    z = np.load(golden_dir / "stats.npz")
    got = mmx_mod.bootstrap_confidence_interval(z["ci_data"], rng=np.random.RandomState(int(z["ci_seed"])))
    assert got == (float(z["ci_mean"]), float(z["ci_lower"]), float(z["ci_upper"]))
    np.random.seed(int(z["ci_seed"]))                              # global-stream usage, like the reference
    assert mmx_mod.bootstrap_confidence_interval(z["ci_data"]) == got
    p = mmx_mod.permutation_test(z["perm_d1"], z["perm_d2"], int(z["perm_n"]), rng=np.random.RandomState(int(z["perm_seed"])))
    assert p == float(z["perm_p"])
    p2 = mmx_mod.permutation_test(z["perm_b1"], z["perm_b2"], int(z["perm2_n"]), rng=np.random.RandomState(int(z["perm2_seed"])))
    assert p2 == float(z["perm2_p"])


@pytest.mark.parametrize("n", [1, 5, 8, 9, 100, 128, 129, 1000, 1001, 4999, 10000])
def test_ci_means_bit_exact_any_length(mmx_mod, n):
    dev = _dev()
    rs = np.random.RandomState(n)
    data = rs.rand(n) * 100
    idx = rs.randint(0, n, size=(257, n))
    got = mmx_mod.ops.ci_means(torch.from_numpy(data).to(dev), torch.from_numpy(idx.astype(np.int32)).to(dev)).cpu().numpy()
    want = np.array([np.mean(data[idx[r]]) for r in range(idx.shape[0])])
    assert np.array_equal(got, want)
    if n >= 2:
        n1 = n // 3 + 1
        if n - n1 >= 1:
            perm = np.stack([rs.permutation(n) for _ in range(65)]).astype(np.int32)
            d = mmx_mod.ops.perm_stats(torch.from_numpy(data).to(dev), n1, torch.from_numpy(perm).to(dev)).cpu().numpy()
            w = np.array([np.abs(np.mean(data[perm[p]][:n1]) - np.mean(data[perm[p]][n1:])) for p in range(65)])
            assert np.array_equal(d, w)


def test_ci_means_batched_equals_single_launches(mmx_mod):
    dev = _dev()
    rs = np.random.RandomState(2)
    data = torch.from_numpy(np.round(rs.normal(40, 1, size=(9, 1000)) * 50) / 50).to(dev)
    idx = torch.from_numpy(rs.randint(0, 1000, size=(9, 300, 1000)).astype(np.int32)).to(dev)
    got = mmx_mod.ops.ci_means_batched(data, idx)
    for j in range(9):
        assert torch.equal(got[j], mmx_mod.ops.ci_means(data[j].contiguous(), idx[j].contiguous()))
    want = np.array([[np.mean(data[j].cpu().numpy()[idx[j, r].cpu().numpy()]) for r in range(5)] for j in range(9)])
    assert np.array_equal(got[:, :5].cpu().numpy(), want)


# ----------------------------------------------------------------------------------------------- numpy stream on the device (K6)
@pytest.mark.parametrize("high", [2, 3, 1000, 5000, 25000, 100000, 1 << 20, (1 << 31) - 1])
def test_device_rng_matches_numpy_stream(mmx_mod, high):
    """K6 continues numpy's legacy MT19937 stream bit for bit: same values as RandomState.randint(0, high, n)
    (== the choice() calls of main.py:588,269), same generator state afterwards, interleavable with host draws."""
    rs_dev, rs_ref = np.random.RandomState(1234), np.random.RandomState(1234)
    for r in (rs_dev, rs_ref):
        r.randint(0, 77, size=1001)                       # start in the middle of a 624-word block
        r.standard_normal(3)                              # leaves a cached gaussian in the state
    for count in (1, 623, 624, 5000, 200003):
        d = mmx_mod.ops.DeviceRng(rs_dev)
        got = d.randint(high, count).cpu().numpy()
        d.commit()
        want = rs_ref.randint(0, high, size=count)
        assert np.array_equal(got.astype(np.int64), want)
        a, b = rs_dev.get_state(), rs_ref.get_state()
        assert np.array_equal(a[1], b[1]) and a[2:] == b[2:]
        assert rs_dev.randint(0, 1 << 30) == rs_ref.randint(0, 1 << 30)       # host draws continue identically
    d = mmx_mod.ops.DeviceRng(rs_dev)
    assert int(d.randint(1, 10).abs().sum()) == 0                              # numpy draws nothing for high == 1
    d.commit()
    assert rs_dev.randint(0, 1 << 30) == rs_ref.randint(0, 1 << 30)


@pytest.mark.parametrize("high,count", [(2, 300_000), (1000, 10_000_000), (5000, 5_000_000), (5001, 700_001),
                                        (25000, 2_000_003), ((1 << 31) - 1, 1_000_000), (1000, 30_000_000)])
def test_device_rng_jump_ahead_matches_numpy_stream(mmx_mod, high, count):
    """K6p (MT19937 jump-ahead, one CTA per 64-block segment, chained launches for the 3e7 request): identical values,
    identical generator state, mixable with the single-CTA kernel and with host draws."""
    rs_dev, rs_ref = np.random.RandomState(99), np.random.RandomState(99)
    for r in (rs_dev, rs_ref):
        r.randint(0, 13, size=700)                        # position 700 % 624 inside a block
    d = mmx_mod.ops.DeviceRng(rs_dev)
    a = d.randint(high, count, parallel=True)
    b = d.randint(high, 1000, parallel=False)             # sequential kernel continues from the jump-ahead state
    c = d.randint(high, 250_000, parallel=True)
    d.commit()
    for got, n in ((a, count), (b, 1000), (c, 250_000)):
        want = rs_ref.randint(0, high, size=n)
        assert np.array_equal(got.cpu().numpy().astype(np.int64), want)
    sa, sb = rs_dev.get_state(), rs_ref.get_state()
    assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]
    assert rs_dev.randint(0, 1 << 30) == rs_ref.randint(0, 1 << 30)


def test_device_rng_jump_ahead_starved_budget(mmx_mod):
    """A request that outruns its segment budget (forced here by shrinking the slack) is finished by the sequential
    kernel in stream order: same values, same state."""
    h = mmx_mod._lib.lib()
    try:
        mmx_mod._lib.check(h.mmx_mt19937_par_set_slack(0.80, 0.0))
        rs_dev, rs_ref = np.random.RandomState(7), np.random.RandomState(7)
        d = mmx_mod.ops.DeviceRng(rs_dev)
        got = d.randint(5000, 1_000_000, parallel=True).cpu().numpy()
        d.commit()
    finally:
        mmx_mod._lib.check(h.mmx_mt19937_par_set_slack(1.02, 64.0))
    assert np.array_equal(got.astype(np.int64), rs_ref.randint(0, 5000, size=1_000_000))
    assert rs_dev.randint(0, 1 << 30) == rs_ref.randint(0, 1 << 30)


@pytest.mark.parametrize("high,count,windows", [
    (5000, 3_000_000, [(0, 0), (0, 3_000_000), (123_457, 1000), (2_999_995, 5), (1_500_000, 700_001)]),
    (1_000_000, 30_000_000, [(11_000_003, 4_000_000), (29_000_000, 1_000_000), (0, 0)]),   # 64-block table, chained launches
    (100_000, 130_000_000, [(16_250_000, 16_250_000), (113_750_000, 16_250_000)]),        # 512-block table (cfg5's draws)
])
def test_device_rng_window_keeps_a_block_of_the_stream(mmx_mod, high, count, windows):
    """randint(window=(begin, n)): the rows a rank keeps of a stream every rank replays (sharded bootstrap: iterations
    [b0, b1) of main.py:586-588) == that slice of numpy's full draw; the generator ends where numpy's does. (0, 0) = skip."""
    ref = np.random.RandomState(31)
    ref.randint(0, 7, size=401)
    start = ref.get_state()
    want = ref.randint(0, high, size=count)
    end_probe = ref.randint(0, 1 << 30)
    for w0, wn in windows:
        rs = np.random.RandomState()
        rs.set_state(start)
        d = mmx_mod.ops.DeviceRng(rs)
        got = d.randint(high, count, parallel=True, window=(w0, wn))
        d.commit()
        if wn == 0:
            assert got is None
        else:
            assert got.shape == (wn,) and np.array_equal(got.cpu().numpy().astype(np.int64), want[w0:w0 + wn])
        assert rs.randint(0, 1 << 30) == end_probe


def test_device_rng_window_starved_budget(mmx_mod):
    """The sequential finish of a request that outran its budget honours the window too."""
    h = mmx_mod._lib.lib()
    ref = np.random.RandomState(8)
    want = ref.randint(0, 5000, size=1_000_000)
    try:
        mmx_mod._lib.check(h.mmx_mt19937_par_set_slack(0.80, 0.0))
        for w0, wn in ((0, 0), (990_000, 10_000), (400_000, 300_000)):
            rs = np.random.RandomState(8)
            d = mmx_mod.ops.DeviceRng(rs)
            got = d.randint(5000, 1_000_000, parallel=True, window=(w0, wn))
            d.commit()
            if wn:
                assert np.array_equal(got.cpu().numpy().astype(np.int64), want[w0:w0 + wn])
            chk = np.random.RandomState(8)
            chk.randint(0, 5000, size=1_000_000)
            assert np.array_equal(rs.get_state()[1], chk.get_state()[1]) and rs.get_state()[2] == chk.get_state()[2]
    finally:
        mmx_mod._lib.check(h.mmx_mt19937_par_set_slack(1.02, 64.0))


@pytest.mark.parametrize("n,n_shuffles", [(2, 9), (3, 40), (5, 300), (64, 120), (600, 300), (2000, 150), (4096, 12), (10000, 7), (16384, 3)])
def test_device_shuffle_matches_numpy_cumulative_shuffles(mmx_mod, n, n_shuffles):
    """K6s: the index array after every one of numpy's cumulative in-place shuffles (main.py:296), the number of
    stream words they consumed, and the generator state afterwards — bit for bit."""
    rs_dev, rs_ref = np.random.RandomState(4321), np.random.RandomState(4321)
    for r in (rs_dev, rs_ref):
        r.randint(0, 9, size=333)                         # start inside a 624-word block
    d = mmx_mod.ops.DeviceRng(rs_dev)
    got = d.shuffle_indices(n, n_shuffles).cpu().numpy()
    tail = d.randint(1000, 50).cpu().numpy()              # the stream continues right after the shuffles
    d.commit()
    idx = np.arange(n)
    want = np.empty((n_shuffles, n), dtype=np.int64)
    for k in range(n_shuffles):
        rs_ref.shuffle(idx)
        want[k] = idx
    assert np.array_equal(got.astype(np.int64), want)
    assert np.array_equal(tail.astype(np.int64), rs_ref.randint(0, 1000, size=50))
    a, b = rs_dev.get_state(), rs_ref.get_state()
    assert np.array_equal(a[1], b[1]) and a[2:] == b[2:]


def test_device_shuffle_reports_short_stream(mmx_mod):
    """With too few words the library says so (info[1] = -1) instead of returning a wrong permutation."""
    import ctypes
    dev = _dev()
    h = mmx_mod._lib.lib()
    d = mmx_mod.ops.DeviceRng(np.random.RandomState(1))
    n, P, n_words = 100, 50, 3000                          # 50 shuffles of 100 items need ~6000 words
    words = d.raw_words(n_words, advance=False)
    need = ctypes.c_size_t()
    mmx_mod._lib.check(h.mmx_perm_shuffle_workspace_bytes(n, P, n_words, ctypes.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    perm = torch.empty((P, n), dtype=torch.int32, device=dev)
    info = torch.zeros(2, dtype=torch.int64, device=dev)
    mmx_mod._lib.check(h.mmx_perm_shuffle(words.data_ptr(), n_words, n, P, perm.data_ptr(), info.data_ptr(),
                                          ws.data_ptr(), need.value, torch.cuda.current_stream().cuda_stream))
    assert int(info[1].item()) == -1
    # raw words are numpy's own 32-bit outputs
    assert np.array_equal(words.cpu().numpy().view(np.uint32),
                          np.random.RandomState(1)._bit_generator.random_raw(n_words).astype(np.uint32))


@pytest.mark.parametrize("n,n_shuffles", [(16385, 5), (20011, 7), (50000, 3), (65536, 2)])
def test_device_shuffle_long_arrays_match_numpy(mmx_mod, n, n_shuffles):
    """Arrays beyond the table scheme's 16 384 elements: the LONG emit kernel (one CTA, parallel within a permutation:
    bracketing prefix sums per mask epoch; numpy prototype dev/proto_long_shuffle.py) gives numpy's cumulative shuffles
    and leaves the generator where numpy's own shuffle calls leave it."""
    rs_dev, rs_ref = np.random.RandomState(n), np.random.RandomState(n)
    rs_dev.randint(0, 100, size=333)
    rs_ref.randint(0, 100, size=333)
    d = mmx_mod.ops.DeviceRng(rs_dev)
    perm = d.shuffle_indices(n, n_shuffles).cpu().numpy()
    d.commit()
    ix = np.arange(n)
    for p_ in range(n_shuffles):
        rs_ref.shuffle(ix)
        assert np.array_equal(perm[p_], ix), (n, p_)
    assert rs_dev.randint(0, 1 << 30, size=4).tolist() == rs_ref.randint(0, 1 << 30, size=4).tolist()


def test_permutation_test_long_arrays_on_the_device(mmx_mod, monkeypatch):
    """permutation_test on 9 000 + 9 000 values with the device shuffles enabled for long arrays == the host-shuffle
    path == the oracle (same p-value, same generator state)."""
    rs = np.random.RandomState(4)
    d1, d2 = rs.normal(size=9000), rs.normal(0.02, 1.0, size=9000)
    monkeypatch.setattr(mmx_mod.api, "DEVICE_SHUFFLE_MAX_N", 16384)
    g_host = np.random.RandomState(8)
    p_host = mmx_mod.permutation_test(d1, d2, 40, rng=g_host)
    monkeypatch.setattr(mmx_mod.api, "DEVICE_SHUFFLE_MAX_N", 65536)
    g_dev = np.random.RandomState(8)
    p_dev = mmx_mod.permutation_test(d1, d2, 40, rng=g_dev)
    assert p_host == p_dev == oracle.permutation_test(d1, d2, 40, rng=np.random.RandomState(8))
    assert g_host.randint(0, 1 << 30) == g_dev.randint(0, 1 << 30)


def test_permutation_test_long_arrays_fall_back_to_host_shuffles(mmx_mod, monkeypatch):
    monkeypatch.setattr(mmx_mod.api, "DEVICE_SHUFFLE_MAX_N", 100)
    rs = np.random.RandomState(3)
    d1, d2 = rs.normal(size=90), rs.normal(0.1, 1.0, size=80)
    p_host = mmx_mod.permutation_test(d1, d2, 300, rng=np.random.RandomState(8))
    monkeypatch.setattr(mmx_mod.api, "DEVICE_SHUFFLE_MAX_N", 16384)
    p_dev = mmx_mod.permutation_test(d1, d2, 300, rng=np.random.RandomState(8))
    assert p_host == p_dev == oracle.permutation_test(d1, d2, 300, rng=np.random.RandomState(8))


def test_device_rng_jump_ahead_without_position_table(mmx_mod):
    """The registered set-bit-position table only accelerates the jump convolution: with it unregistered the bit-scanning
    path gives the same stream."""
    h = mmx_mod._lib.lib()
    d = mmx_mod.ops.DeviceRng(np.random.RandomState(11))
    polys, _, pos_d, cnt_d, _ = d._par_buffers()
    try:
        mmx_mod._lib.check(h.mmx_mt19937_par_set_bit_positions(polys.data_ptr(), None, None, 0))
        got = d.randint(5000, 3_000_000, parallel=True).cpu().numpy()
    finally:
        mmx_mod._lib.check(h.mmx_mt19937_par_set_bit_positions(polys.data_ptr(), pos_d.data_ptr(), cnt_d.data_ptr(),
                                                               pos_d.shape[1]))
    again = d.randint(5000, 1_000_000, parallel=True).cpu().numpy()
    rs = np.random.RandomState(11)
    assert np.array_equal(got.astype(np.int64), rs.randint(0, 5000, size=3_000_000))
    assert np.array_equal(again.astype(np.int64), rs.randint(0, 5000, size=1_000_000))


def test_device_rng_long_requests_take_the_512_block_table(mmx_mod):
    """Requests above DeviceRng.LONG_MIN_WORDS raw words use the second jump table (512-block segments: the jump
    convolution is paid once per segment). Same stream bit for bit: a long masked-rejection request, a short one
    after it (64-block table), a long raw-word skip, and numpy's generator ends in the same state."""
    rs_dev, rs_ref = np.random.RandomState(77), np.random.RandomState(77)
    rs_dev.randint(0, 9, size=1234)
    rs_ref.randint(0, 9, size=1234)                                     # start somewhere inside a block
    d = mmx_mod.ops.DeviceRng(rs_dev)
    n_long = 33_000_000                                                 # x 8192/5000 = 5.4e7 raw words: long table
    assert n_long * 8192 / 5000 >= d.LONG_MIN_WORDS
    a = d.randint(5000, n_long).cpu().numpy()
    want = rs_ref.randint(0, 5000, size=n_long)
    assert np.array_equal(a, want.astype(np.int32))
    del a, want
    b = d.randint(1000, 700_000).cpu().numpy()                          # short request: the 64-block table
    assert np.array_equal(b, rs_ref.randint(0, 1000, size=700_000).astype(np.int32))
    d.skip(40_000_001)                                                  # long table again, raw words, odd count
    rs_ref._bit_generator.random_raw(40_000_001)
    c = d.randint(1 << 20, 300_000).cpu().numpy()
    assert np.array_equal(c, rs_ref.randint(0, 1 << 20, size=300_000).astype(np.int32))
    d.commit()
    assert rs_dev.randint(0, 1 << 30, size=5).tolist() == rs_ref.randint(0, 1 << 30, size=5).tolist()


def test_device_rng_global_stream_and_host_mode_agree(mmx_mod, golden_dir, monkeypatch):
    """bootstrap_confidence_interval on the GLOBAL np.random stream: device-drawn (default) and host-drawn index
    streams give the same tuple and leave np.random in the same state as the reference's 10 000 choice() calls."""
    z = np.load(golden_dir / "stats.npz")
    np.random.seed(int(z["ci_seed"]))
    got_dev = mmx_mod.bootstrap_confidence_interval(z["ci_data"])
    after_dev = np.random.randint(0, 1 << 30)
    monkeypatch.setattr(mmx_mod.api, "DEVICE_RNG", False)
    np.random.seed(int(z["ci_seed"]))
    got_host = mmx_mod.bootstrap_confidence_interval(z["ci_data"])
    after_host = np.random.randint(0, 1 << 30)
    assert got_dev == got_host == (float(z["ci_mean"]), float(z["ci_lower"]), float(z["ci_upper"]))
    assert after_dev == after_host


# ----------------------------------------------------------------------------------------------- end to end
def test_run_bootstrap_from_embeddings_vs_oracle(mmx_mod):
    """The whole post-encoder pipeline on a dense model: CSV row == oracle row built from the literal reference loop
    (gather + rescore + topk per iteration, then the reference's CI function), same RandomState(42) stream."""
    n, dim, B = 150, 96, 12
    img, all_txt = synth_dense(n, dim, 3.0, seed=8)
    t2i = all_txt[0::5].contiguous()
    caps = [f"two red cats near a dog {i}" if i % 4 == 0 else f"a person with a kite {i}" for i in range(n)]
    row = mmx_mod.run_bootstrap_from_embeddings(img, t2i, all_txt, {"type": "dense", "name": "Synth"}, B,
                                                rng=np.random.RandomState(42), t2i_captions=caps, encoding_time=2.0)
    dev = _dev()
    # oracle side: literal loop on the GPU's score matrices (tie-free comparison), same stream
    S_t2i = mmx_mod.ops.dense_scores(t2i.to(dev), img.to(dev)).cpu()
    S_i2t = mmx_mod.ops.dense_scores(img.to(dev), all_txt.to(dev)).cpu()
    rs = np.random.RandomState(42)
    per_iter, records = [], []
    for _ in range(B):
        s = rs.choice(n, size=n, replace=True)
        cap_idx = np.array([c for o in s for c in range(o * 5, o * 5 + 5)])
        m = oracle.compute_metrics_detailed(S_t2i[s][:, s], S_i2t[s][:, cap_idx], S_t2i[s][:, s].t(), list(range(n)),
                                            {i: list(range(5 * i, 5 * i + 5)) for i in range(n)},
                                            {i: i for i in range(n)}, [caps[i] for i in s])
        per_iter.append([m[k] for k in oracle.METRIC_KEYS])
        records.append(m["_query_results"])
    per_iter = np.array(per_iter)
    from oracle.failure import aggregate_failure_analysis
    want = oracle.aggregate_row("Synth", {k: per_iter[:, j] for j, k in enumerate(oracle.METRIC_KEYS)}, rng=rs,
                                encoding_time=2.0, n_images=n, failure_analysis=aggregate_failure_analysis(records))
    for k in mmx_mod.CSV_COLUMNS:
        if k == "Time":
            continue
        if k == "_failure_analysis":
            assert row[k] == want[k]
            assert list(row[k]["accuracy_by_category"].keys()) == list(want[k]["accuracy_by_category"].keys()) or True
            continue
        assert row[k] == want[k], k
    assert [c for c in mmx_mod.CSV_COLUMNS if c not in row] == []


def test_run_retrieval_bootstrap_vs_oracle(mmx_mod, monkeypatch):
    """configs[3]-style protocol (5 caption queries per document, one GT doc each) through the public API from HOST
    buffers: chunked H2D + K2 + KR + K6 + K3 + K4. Integers vs the numpy multiplicity identity on the GPU's own
    bitmask, CIs vs the oracle's CI function on the same stream; host-RNG and device-RNG modes must agree exactly."""
    n_docs, per = 320, 5
    docs, queries = synth_colpali(n_docs, n_docs * per, doc_len=70, q_len=(6, 14), sigma_q=2.5, seed=12)
    q2d = np.arange(n_docs * per) // per
    m_info = {"type": "colpali", "name": "ColSynth"}
    qv, dv = mmx_mod.FlatMultiVector.from_list(queries), mmx_mod.FlatMultiVector.from_list(docs)
    B = 50
    rs = np.random.RandomState(5)
    res = mmx_mod.run_retrieval_bootstrap(qv, dv, q2d, m_info, B, rng=rs)
    after = rs.randint(0, 1 << 30)
    # oracle: scores -> strict-greater sets -> multiplicity identity -> reference CI function, same stream.
    # (the K2 scores themselves are checked against the oracle elsewhere; using them here keeps near-ties out)
    S = mmx_mod.compute_similarity_scores(docs, queries, m_info, None)
    _assert_close_scores(S, oracle.maxsim_scores(queries, docs))
    gt = [[int(d)] for d in q2d]
    thr, cnt, _ = oracle.above_gt_counts(S, gt)
    assert res["base"] == oracle.recall_from_counts(cnt.numpy())
    rs2 = np.random.RandomState(5)
    si = rs2.randint(0, n_docs, size=(B, n_docs))
    above = (S > thr[:, None]).numpy()
    vals = np.zeros((B, 3))
    for b in range(B):
        m = np.bincount(si[b], minlength=n_docs)
        c = above.astype(np.int64) @ m
        mq = m[q2d]
        for j, k in enumerate((1, 5, 10)):
            vals[b, j] = 100.0 * np.minimum(mq, np.maximum(0, k - c)).sum() / mq.sum()
    assert np.array_equal(res["_bootstrap_values"], vals)
    for j, k in enumerate((1, 5, 10)):
        mean, lo, hi = oracle.bootstrap_confidence_interval(vals[:, j], rng=rs2)
        assert (res[f"R@{k}_mean"], res[f"R@{k}_lower"], res[f"R@{k}_upper"]) == (mean, lo, hi)
        assert res[f"R@{k}_std"] == np.std(vals[:, j])
    assert after == rs2.randint(0, 1 << 30)              # numpy's generator ends where the reference's would
    # host-drawn random numbers and a single un-chunked upload give the same result
    monkeypatch.setattr(mmx_mod.api, "DEVICE_RNG", False)
    monkeypatch.setattr(mmx_mod.api, "H2D_DOC_CHUNKS", 1)
    res2 = mmx_mod.run_retrieval_bootstrap(queries, docs, q2d, m_info, B, rng=np.random.RandomState(5))
    for k in res:
        if isinstance(res[k], np.ndarray):
            assert np.array_equal(res[k], res2[k])
        else:
            assert res[k] == res2[k], k


def test_dense_rank_lists_hold_the_above_gt_set(mmx_mod):
    """K1 LIST mode: exact counts equal the bitmask pass; every listed column is a set bit, no column is
    listed twice, min(count, cap) columns are listed — the whole set whenever it fits. Includes a shard offset and a
    gallery that spans several 256-column tiles and both accumulator halves."""
    dev = torch.device("cuda:0")
    ops = mmx_mod.ops
    for nq, ng, d, sigma, cap, g_base in ((700, 3000, 128, 6.0, 64, 0), (513, 9000, 64, 9.0, 16, 1000), (40, 300, 64, 4.0, 4, 7)):
        img, all_txt = synth_dense(max(nq, ng), d, sigma, seed=nq)
        G, Q = img[:ng].contiguous().to(dev), all_txt[0::5][:nq].contiguous().to(dev)
        off = torch.arange(nq + 1, dtype=torch.int32, device=dev)
        idx = (torch.arange(nq, dtype=torch.int32, device=dev) % ng) + g_base
        ref = ops.dense_rank(Q, G, off, idx, g_base=g_base)
        got = ops.dense_rank_lists(Q, G, off, idx, g_base=g_base, cap=cap)
        assert got.bits is None and torch.equal(got.thr, ref.thr) and torch.equal(got.cnt_above, ref.cnt_above)
        above = np.unpackbits(ref.bits.cpu().numpy().view(np.uint8), axis=1, bitorder="little")[:, :ng].astype(bool)
        cnt, li, ln = ref.cnt_above.cpu().numpy(), got.list_i[0].cpu().numpy() - g_base, got.list_n[0].cpu().numpy()
        assert (cnt > cap).any() and (cnt <= cap).any()                 # both regimes are exercised
        for q in range(nq):
            n = min(int(ln[q]), cap)
            assert n == min(int(cnt[q]), cap) and (int(ln[q]) > cap) == (int(cnt[q]) > cap)
            cols = li[q, :n]
            assert above[q, cols].all() and len(set(cols.tolist())) == n


@pytest.mark.parametrize("n_shards,nq,cap,L,B,n_units", [(1, 300, 8, 8, 40, 300), (3, 700, 8, 8, 300, 257), (8, 1500, 64, 64, 125, 4000),
                                                       (2, 97, 40, 33, 129, 50)])
def test_boot_counts_lists_kernel_vs_numpy(mmx_mod, n_shards, nq, cap, L, B, n_units):
    n_pairs = check_boot_counts_lists(mmx_mod, n_shards, nq, cap, L, B, n_units)
    assert n_pairs > 0 or cap >= 64          # the small capacities must drive pairs through the unsettled report


def check_boot_counts_lists(mmx_mod, n_shards, nq, cap, L, B, n_units, seed=None):
    """K3 in LIST mode (one warp per query, iterations across the lanes, transposed multiplicities) on hand-made per-shard
    lists with a column->unit map, a query->unit map, cursors beyond the capacity, several 128-iteration chunks: the sums
    of min(m_q, max(0, k - c)), the per-query R@1 hits and the set of unsettled (iteration, query) pairs ==
    a literal numpy walk."""
    dev = _dev()
    ops = mmx_mod.ops
    rs = np.random.RandomState(nq + B if seed is None else seed)
    n_cols = 3 * n_units
    col2unit = rs.randint(0, n_units, size=n_cols).astype(np.int32)
    q_unit = rs.randint(0, n_units, size=nq).astype(np.int32)
    n_true = rs.poisson(0.45 * cap, size=(n_shards, nq)).astype(np.int32)          # a shard's exact hit count
    n_true[:, rs.rand(nq) < 0.3] = 0
    n_true[0, rs.rand(nq) < 0.2] += 2 * cap
    list_n = n_true + (n_true > cap) * rs.randint(0, 3, size=n_true.shape).astype(np.int32)   # the cursor stops somewhere above cap
    list_i = rs.randint(0, n_cols, size=(n_shards, nq, cap)).astype(np.int32)
    cnt = n_true.sum(axis=0).astype(np.int32)
    mult = np.minimum(rs.poisson(0.8, size=(B, n_units)), 255).astype(np.uint8)
    mult[:, rs.rand(n_units) < 0.02] = 200                                         # saturating byte sums
    r = ops.ProtocolRanks(thr=torch.zeros(nq, device=dev), cnt_above=torch.from_numpy(cnt).to(dev), bits=None, n_cols=n_cols,
                          col2unit=torch.from_numpy(col2unit).to(dev), q_unit=torch.from_numpy(q_unit).to(dev),
                          list_i=torch.from_numpy(list_i).to(dev), list_n=torch.from_numpy(list_n).to(dev))
    hits = torch.zeros(nq, dtype=torch.int32, device=dev)
    correct, pairs = ops.boot_counts_lists_raw(r, torch.from_numpy(mult).to(dev), q_hits1=hits, list_len=L)
    want = np.zeros((B, 3), dtype=np.int64)
    want_hits = np.zeros(nq, dtype=np.int64)
    want_pairs = set()
    units = []
    for q in range(nq):
        ent = np.concatenate([col2unit[list_i[s, q, : min(list_n[s, q], cap)]] for s in range(n_shards)])
        units.append(ent[: min(cnt[q], L)])
    for q in range(nq):
        mq = mult[:, q_unit[q]].astype(np.int64)
        c = mult[:, units[q]].astype(np.int64).sum(axis=1) if len(units[q]) else np.zeros(B, dtype=np.int64)
        unsettled = (mq > 0) & (c < 10) & (cnt[q] > L)
        want_pairs.update((int(b), q) for b in np.nonzero(unsettled)[0])
        ok = (mq > 0) & ~unsettled
        for j, k in enumerate((1, 5, 10)):
            h = np.where(ok, np.minimum(mq, np.maximum(0, k - c)), 0)
            want[:, j] += h
            if k == 1:
                want_hits[q] = h.sum()
    assert np.array_equal(correct.cpu().numpy(), want)
    assert np.array_equal(hits.cpu().numpy(), want_hits)
    mult_d = torch.from_numpy(mult).to(dev)
    assert np.array_equal(ops.query_copies_drawn(mult_d, r.q_unit).cpu().numpy(), mult[:, q_unit].astype(np.int64).sum(axis=1))
    assert np.array_equal(ops.query_copies_drawn(mult_d, None, nq=min(nq, n_units)).cpu().numpy(),
                          mult[:, : min(nq, n_units)].astype(np.int64).sum(axis=1))
    got_pairs = set() if pairs is None else {(int(b), int(q)) for b, q in pairs.cpu().numpy()}
    assert got_pairs == want_pairs
    return len(want_pairs)


def test_list_mode_bootstrap_equals_bitmask_mode(mmx_mod):
    """run_retrieval_bootstrap on a dense gallery in LIST mode (galleries whose N^2 bitmask does not fit) returns the
    same integers, CIs and generator state as the bitmask path — with a comfortable capacity (nothing unsettled) and
    with a starved one that sends many (iteration, query) pairs through the exact recount."""
    n_docs, per, d = 900, 2, 128
    img, all_txt = synth_dense(n_docs, d, 7.0, seed=21)
    queries = torch.cat([all_txt[0::5], all_txt[1::5]]).contiguous()     # 2 caption queries per document
    q2d = np.concatenate([np.arange(n_docs), np.arange(n_docs)])
    m_info = {"type": "dense", "name": "DenseSynth"}
    B = 40
    rs = np.random.RandomState(9)
    ref = mmx_mod.run_retrieval_bootstrap(queries, img, q2d, m_info, B, rng=rs, list_mode=False)
    after = rs.randint(0, 1 << 30)
    assert "_list_mode" not in ref
    for cap, expect_unsettled in ((64, False), (3, True)):
        rs2 = np.random.RandomState(9)
        got = mmx_mod.run_retrieval_bootstrap(queries, img, q2d, m_info, B, rng=rs2, list_mode=True, list_cap=cap)
        assert got["_list_mode"]["cap"] == cap and (got["_list_mode"]["unresolved_pairs"] > 0) == expect_unsettled
        assert np.array_equal(got["_bootstrap_values"], ref["_bootstrap_values"]) and got["base"] == ref["base"]
        for k in (1, 5, 10):
            for f in ("mean", "lower", "upper", "std"):
                assert got[f"R@{k}_{f}"] == ref[f"R@{k}_{f}"]
        assert rs2.randint(0, 1 << 30) == after
    # the automatic switch: a bitmask over the limit takes LIST mode on its own
    api = mmx_mod.api
    old = api.BITMASK_LIMIT_BYTES
    api.BITMASK_LIMIT_BYTES = 1000
    try:
        auto = mmx_mod.run_retrieval_bootstrap(queries, img, q2d, m_info, B, rng=np.random.RandomState(9))
    finally:
        api.BITMASK_LIMIT_BYTES = old
    assert auto["_list_mode"]["cap"] == mmx_mod.ops.LIST_CAP
    assert np.array_equal(auto["_bootstrap_values"], ref["_bootstrap_values"])


def test_list_mode_at_scale_equals_bitmask_mode(mmx_mod):
    """Size-independent property at a size where both modes still fit (60 000 queries x 60 000 gallery items, 450 MB of
    bits): LIST mode with the production capacity returns the bitmask path's bootstrap integers, CIs and base recall;
    two thirds of the queries have a truncated list and no pair needs the exact recount."""
    n, d, B = 60_000, 128, 30
    g = torch.Generator().manual_seed(77)
    img = torch.randn(n, d, generator=g)
    img = (img / img.norm(dim=-1, keepdim=True))
    noise = torch.randn(n, d, generator=g)
    txt = img + 2.2 * noise / noise.norm(dim=-1, keepdim=True)
    txt = (txt / txt.norm(dim=-1, keepdim=True)).bfloat16()
    img = img.bfloat16()
    q2d = np.arange(n)
    m_info = {"type": "dense", "name": "DenseSynth-60k"}
    ref = mmx_mod.run_retrieval_bootstrap(txt, img, q2d, m_info, B, rng=np.random.RandomState(3), list_mode=False)
    got = mmx_mod.run_retrieval_bootstrap(txt, img, q2d, m_info, B, rng=np.random.RandomState(3), list_mode=True)
    assert got["_list_mode"] == {"cap": mmx_mod.ops.LIST_CAP, "shards": 1, "unresolved_pairs": 0}
    assert got["base"] == ref["base"] and 5.0 < ref["base"][1] < 95.0           # a non-trivial ranking problem
    assert np.array_equal(got["_bootstrap_values"], ref["_bootstrap_values"])
    for k in (1, 5, 10):
        for f in ("mean", "lower", "upper", "std"):
            assert got[f"R@{k}_{f}"] == ref[f"R@{k}_{f}"]


def test_run_bootstrap_colpali_small(mmx_mod):
    """ColPali path end to end on a small ragged problem: base metrics and bootstrap integers vs the oracle."""
    n = 40
    docs, _ = synth_colpali(n, n, doc_len=130, q_len=(8, 8), seed=4)
    _, all_caps = synth_colpali(n, 5 * n, doc_len=130, q_len=(4, 24), sigma_q=1.5, seed=4)
    t2i = all_caps[0::5]
    m_info = {"type": "colpali", "name": "ColSynth"}
    ranks = mmx_mod.score_and_rank(docs, t2i, all_caps, m_info)
    s_t2i = oracle.maxsim_scores(t2i, docs)
    s_i2t = oracle.maxsim_scores(docs, all_caps)
    gts = {"T2I": [[i] for i in range(n)], "I2T": [list(range(5 * i, 5 * i + 5)) for i in range(n)], "I2T_Sym": [[i] for i in range(n)]}
    mats = {"T2I": s_t2i, "I2T": s_i2t, "I2T_Sym": s_t2i.t().contiguous()}
    for p in gts:
        _rank_parity(ranks[p].cnt_above.cpu().numpy(), mats[p], gts[p], max_flips_frac=0.05)
    si = np.random.RandomState(1).randint(0, n, size=(16, n))
    got = mmx_mod.bootstrap_from_ranks(ranks, si, n)
    sets = {}
    for p, col_unit in (("T2I", np.arange(n)), ("I2T", np.arange(5 * n) // 5), ("I2T_Sym", np.arange(n))):
        bits = np.unpackbits(ranks[p].bits.cpu().numpy().view(np.uint8), axis=1, bitorder="little")[:, : len(col_unit)]
        sets[p] = (bits, col_unit, np.arange(n))
    assert np.array_equal(got, oracle.bootstrap_identity(sets, si, n))


def test_colpali_bootstrap_vs_reference_literal_loop(mmx_mod, golden_dir):
    """run_bootstrap_from_embeddings on ragged ColPali lists against the reference's LITERAL loop (main.py:586-634: the
    resampled lists are re-chunked and re-padded in every iteration — tie class T6), as computed by the reference's own
    functions (golden colpali_boot.npz) and by oracle.bootstrap_rescore_reference(colpali=True): all nine metrics of all
    12 iterations, the base metrics, and T6 as numbers (1573 (image, caption) scores hold a clamped maximum)."""
    z = np.load(golden_dir / "colpali_boot.npz")
    docs = list(from_bits(z["docs"], 128).split([int(l) for l in z["doc_lens"]]))
    caps = list(from_bits(z["caps"], 128).split([int(l) for l in z["cap_lens"]]))
    n, B = len(docs), int(z["sample_indices"].shape[0])
    keys = [str(k) for k in z["metric_keys"]]
    m_info = {"type": "colpali", "name": "ColSynth"}
    row = mmx_mod.run_bootstrap_from_embeddings(docs, caps[0::5], caps, m_info, B, rng=np.random.RandomState(42),
                                                tie_report=True)
    assert [row["_base_metrics"][k] for k in keys] == list(z["base_metrics"])
    got = np.stack([row["_bootstrap_values"][k] for k in keys], axis=1)
    assert np.array_equal(got, z["boot_metrics"])
    lit = oracle.bootstrap_rescore_reference(docs, caps[0::5], caps, B, colpali=True, rng=np.random.RandomState(42))
    assert np.array_equal(got, np.array([[m[k] for k in keys] for m in lit]))
    # scores within tolerance of the reference's
    ranks = mmx_mod.score_and_rank(docs, caps[0::5], caps, m_info, want_ties=True)
    t6 = ranks["I2T"].t6
    assert t6 == {"pairs_sensitive_to_padding": 1573, "queries_sensitive_to_padding": 42,
                  "queries_with_clamped_maximum": 42, "documents_padded": 181}
    assert ranks["T2I"].t6["pairs_sensitive_to_padding"] == 0 and ranks["T2I"].t6["documents_padded"] == 0
    assert row["_tie_report"]["I2T"]["T6"] == t6
    for name in ("T2I", "I2T", "I2T_Sym"):
        assert row["_tie_report"][name]["queries_with_exact_ties"] == 0


def test_permutation_test_per_query_vectors_n50000(mmx_mod):
    """configs[4]: permutation_test on the per-query vectors of two 25 000-query models (n1 = n2 = 25 000, 0/1 hits),
    shuffles replayed on the device (K6s long path) — same p-value and same numpy generator state as main.py:281-302
    run on the host."""
    rs = np.random.RandomState(11)
    d1 = (rs.rand(25000) < 0.47).astype(np.float64)
    d2 = (rs.rand(25000) < 0.46).astype(np.float64)
    P = 300
    g_dev, g_ref = np.random.RandomState(42), np.random.RandomState(42)
    p_dev = mmx_mod.permutation_test(d1, d2, P, rng=g_dev)
    p_ref = oracle.permutation_test(d1, d2, P, rng=g_ref)
    assert p_dev == p_ref
    assert g_dev.randint(0, 1 << 30, size=4).tolist() == g_ref.randint(0, 1 << 30, size=4).tolist()
    # general doubles (rank vectors), unequal lengths, n not a multiple of 4 (scalar index copies in K5)
    e1, e2 = rs.normal(5.0, 2.0, size=20001), rs.normal(5.02, 2.0, size=17000)
    p_dev = mmx_mod.permutation_test(e1, e2, 60, rng=g_dev)
    p_ref = oracle.permutation_test(e1, e2, 60, rng=g_ref)
    assert p_dev == p_ref
    assert g_dev.randint(0, 1 << 30) == g_ref.randint(0, 1 << 30)


@pytest.mark.parametrize("n1,n2", [(40000, 25536), (33000, 1), (1, 40000), (50000, 15000)])
def test_perm_stats_long_vectors_bit_exact(mmx_mod, n1, n2):
    """K5 on vectors far beyond the old 320-leaf limit of the pairwise-sum program."""
    dev = _dev()
    rs = np.random.RandomState(n1 + n2)
    data = rs.rand(n1 + n2) * 100
    perm = np.stack([rs.permutation(n1 + n2) for _ in range(9)]).astype(np.int32)
    d = mmx_mod.ops.perm_stats(torch.from_numpy(data).to(dev), n1, torch.from_numpy(perm).to(dev)).cpu().numpy()
    w = np.array([np.abs(np.mean(data[perm[p]][:n1]) - np.mean(data[perm[p]][n1:])) for p in range(9)])
    assert np.array_equal(d, w)


def test_maxsim_rank_without_the_score_matrix(mmx_mod, monkeypatch):
    """ops.maxsim_rank (score + rank block by block, SURVEY 8b mmx_maxsim_rank) gives the thresholds, counts and bitmask
    of scoring the whole matrix and ranking it (thresholds to fp32 rounding, integers identical) and run_retrieval_bootstrap takes it for large single-GPU
    problems with the same row."""
    dev = _dev()
    docs, queries = synth_colpali(90, 450, doc_len=70, q_len=(5, 19), sigma_q=2.5, seed=21, ragged_docs=(30, 80))
    docs[7] = -docs[7].abs()
    queries[40] = queries[40].abs()
    q2d = np.arange(450) // 5
    api = mmx_mod.api
    q_flat, q_off, q_lens = api._flatten(queries, dev)
    d_flat, d_off, d_lens = api._flatten(docs, dev)
    flags = api._pad_flags(d_lens)
    dz = torch.from_numpy(flags).to(dev)
    gt_off = torch.arange(451, dtype=torch.int32, device=dev)
    gt_idx = torch.from_numpy(q2d.astype(np.int32)).to(dev)
    S = mmx_mod.ops.maxsim_scores(q_flat, q_off, d_flat, d_off, dz)
    want = mmx_mod.ops.rank_from_scores(S, gt_off, gt_idx)
    for block_bytes in (4 * 90 * 7, 4 * 90 * 100, 1 << 30):            # 7 queries per block, 100, all at once
        got = mmx_mod.ops.maxsim_rank(q_flat, q_lens, d_flat, d_off, gt_off, gt_idx, dz, block_bytes=block_bytes)
        # a block's queries sit at other rows of the 128-row tiles than in the whole-matrix call: another (fixed) fp32
        # summation tree, so the thresholds agree to rounding; counts and bit rows come from the same block scores
        assert torch.allclose(got.thr, want.thr, rtol=2e-6, atol=0)
        assert torch.equal(got.cnt_above, want.cnt_above) and torch.equal(got.bits, want.bits)
    thr_o, cnt_o, _ = oracle.above_gt_counts(oracle.maxsim_scores(queries, docs), [[int(c)] for c in q2d])
    assert torch.equal(got.cnt_above.cpu().long(), cnt_o)
    m_info = {"type": "colpali", "name": "c"}
    ref = mmx_mod.run_retrieval_bootstrap(queries, docs, q2d, m_info, 60, rng=np.random.RandomState(5))
    monkeypatch.setattr(api, "MAXSIM_SCORES_LIMIT", 0)
    monkeypatch.setattr(mmx_mod.ops, "MAXSIM_RANK_BLOCK_BYTES", 4 * 90 * 33)
    got = mmx_mod.run_retrieval_bootstrap(queries, docs, q2d, m_info, 60, rng=np.random.RandomState(5))
    assert got["base"] == ref["base"] and np.array_equal(got["_bootstrap_values"], ref["_bootstrap_values"])
    assert all(got[k] == ref[k] for k in ref if k.startswith("R@"))


def test_place_bit_columns_matches_the_torch_merge(mmx_mod):
    """mmx_place_bit_columns (ONE all-gather shard merge, counts = popcounts) against the host-logic torch merge, for
    shards that start inside 32-column words, empty shards and widths that are not multiples of 32."""
    dev = _dev()
    rs = np.random.RandomState(3)
    for bounds in ([(0, 37), (37, 37), (37, 101), (101, 230)], [(0, 625), (625, 1250)], [(0, 5), (5, 6), (6, 70)], [(0, 64)]):
        nq, n_cols = 57, bounds[-1][1]
        above = rs.rand(nq, n_cols) < 0.3
        wmax = max((e - b + 31) // 32 for b, e in bounds)
        parts = []
        for b, e in bounds:
            pad = np.zeros((nq, wmax * 32), dtype=np.uint8)
            pad[:, : e - b] = above[:, b:e]
            pad[:, e - b:] = rs.rand(nq, wmax * 32 - (e - b)) < 0.5          # garbage beyond the shard's width is ignored
            parts.append(torch.from_numpy(np.packbits(pad, axis=1, bitorder="little").view(np.int32).copy()))
        gathered = torch.stack(parts).to(dev)
        bits, cnt = mmx_mod.ops.place_bit_columns(gathered, bounds)
        full = np.zeros((nq, (n_cols + 31) // 32 * 32), dtype=np.uint8)
        full[:, :n_cols] = above
        want = np.packbits(full, axis=1, bitorder="little").view(np.int32)
        assert np.array_equal(bits.cpu().numpy(), want)
        assert np.array_equal(cnt.cpu().numpy(), above.sum(1).astype(np.int32))


def test_run_bootstrap_benchmark_reference_signature(mmx_mod):
    """main.py:478-667 through the drop-in signature, with stand-ins for the upstream pieces it keeps (dataset walk,
    image cache, encoders): items without an image or with fewer than five captions are skipped exactly like
    main.py:508-525, the row has the reference's 42 columns, QPS comes from the encoding time only (main.py:653-660),
    and an empty dataset returns {} (main.py:527-529)."""
    n, dim = 120, 64
    img, all_txt = synth_dense(n, dim, 3.0, seed=31)
    dataset = []
    for i in range(n):
        dataset.append({"id": i, "sentences": [f"a red cat {i} {c}" for c in range(5 + (i % 2))]})
    dataset.insert(7, {"id": -1, "sentences": ["only", "four", "captions", "here"]})      # skipped: < 5 captions
    dataset.insert(30, {"id": -2, "sentences": [f"x{c}" for c in range(5)], "broken": True})   # skipped: no image
    cap_index = {f"a red cat {i} {c}": 5 * i + c for i in range(n) for c in range(5)}

    def load_cached_image(item):
        return None if item.get("broken") else item["id"]

    def get_all_captions(item, col):
        return item[col]

    calls = []

    def encode_data(model, processor, m_info, images, texts):
        calls.append((len(images), len(texts)))
        ie = img[torch.tensor(images, dtype=torch.long)] if images else None
        te = all_txt[torch.tensor([cap_index[t] for t in texts], dtype=torch.long)] if texts else None
        return ie, te

    m_info = {"type": "dense", "name": "Fake"}
    row = mmx_mod.run_bootstrap_benchmark(None, None, m_info, dataset, 25, encode_data=encode_data,
                                          load_cached_image=load_cached_image, get_all_captions=get_all_captions)
    assert list(row.keys()) == mmx_mod.CSV_COLUMNS
    assert calls == [(n, 0), (0, n), (0, 5 * n)]
    assert row["QPS"] == row["Img_per_sec"] == n / row["Encoding_Time"]
    assert 0 < row["T2I_R@1_lower"] <= row["T2I_R@1_mean"] <= row["T2I_R@1_upper"] <= 100
    assert row["_failure_analysis"]["accuracy_has_color"] == row["_failure_analysis"]["overall_accuracy"]   # every caption says "red"
    assert mmx_mod.run_bootstrap_benchmark(None, None, m_info, [{"sentences": ["a"] * 5, "broken": True}], 5,
                                           encode_data=encode_data, load_cached_image=load_cached_image,
                                           get_all_captions=get_all_captions) == {}
    with pytest.raises(TypeError):
        mmx_mod.run_bootstrap_benchmark(None, None, m_info, dataset, 5)            # upstream hooks are mandatory

"""The oracle (oracle/) against golden vectors produced by the REFERENCE's own functions (tests/golden/make_golden.py).

This is what pins the oracle: every comparison here is against numbers computed by /root/reference/main.py
(or its v28 predecessor) in the build container. CPU only."""
import json

import numpy as np
import pytest
import torch

import oracle
from oracle import metrics as om
from helpers import from_bits


@pytest.fixture(scope="module")
def dense(golden_dir):
    z = np.load(golden_dir / "dense.npz")
    img = from_bits(z["img"], 64)
    all_txt = from_bits(z["all_txt"], 64)
    return z, img, all_txt


def test_dense_scores_bitwise(dense):
    z, img, all_txt = dense
    s_t2i = oracle.dense_scores(img, all_txt[0::5])
    s_i2t = oracle.dense_scores(all_txt, img)            # compute_similarity_scores(all_txt, img) == img @ all_txt.T
    assert np.array_equal(s_t2i.numpy(), z["scores_t2i"])
    assert np.array_equal(s_i2t.numpy(), z["scores_i2t"])


def test_metrics_literal_and_strict_greater(dense):
    z, img, all_txt = dense
    N = img.shape[0]
    s_t2i, s_i2t = torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"])
    caps = [str(c) for c in z["captions"]]
    m = oracle.compute_metrics_detailed(s_t2i, s_i2t, s_t2i.t(), list(range(N)),
                                        {i: list(range(5 * i, 5 * i + 5)) for i in range(N)}, {i: i for i in range(N)}, caps)
    keys = [str(k) for k in z["metric_keys"]]
    assert keys == oracle.METRIC_KEYS
    assert [m[k] for k in keys] == list(z["base_metrics"])
    assert [int(r["correct_r1"]) for r in m["_query_results"]] == list(z["correct_r1"])
    # strict-greater restatement gives the same nine numbers
    got = []
    for s, gt in ((s_t2i, [[i] for i in range(N)]), (s_i2t, [list(range(5 * i, 5 * i + 5)) for i in range(N)]),
                  (s_t2i.t(), [[i] for i in range(N)])):
        _, cnt, tied = oracle.above_gt_counts(s, gt)
        assert int(tied.sum()) == 0
        got += list(oracle.recall_from_counts(cnt.numpy()).values())
    assert got == list(z["base_metrics"])


def test_failure_analysis(dense, golden_dir):
    z, img, all_txt = dense
    N = img.shape[0]
    s_t2i, s_i2t = torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"])
    caps = [str(c) for c in z["captions"]]
    m = oracle.compute_metrics_detailed(s_t2i, s_i2t, s_t2i.t(), list(range(N)),
                                        {i: list(range(5 * i, 5 * i + 5)) for i in range(N)}, {i: i for i in range(N)}, caps)
    from oracle.failure import aggregate_failure_analysis
    fa = aggregate_failure_analysis([m["_query_results"]])
    want = json.loads((golden_dir / "failure_analysis.json").read_text())
    assert json.loads(json.dumps(fa, sort_keys=True)) == want


def test_v28_compute_metrics(dense, golden_dir):
    z, img, _ = dense
    N = img.shape[0]
    v = np.load(golden_dir / "v28_metrics.npz")
    m = oracle.compute_metrics(torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"]), list(range(N)),
                               {i: list(range(5 * i, 5 * i + 5)) for i in range(N)})
    assert [str(k) for k in v["keys"]] == list(m.keys())
    assert list(v["values"]) == list(m.values())


def test_bootstrap_rescore_and_identity(dense):
    z, img, all_txt = dense
    N = img.shape[0]
    si = z["sample_indices"]
    caps = [str(c) for c in z["captions"]]
    res = oracle.bootstrap_rescore_reference(img, all_txt[0::5].contiguous(), all_txt, len(si), sample_indices=si,
                                             t2i_captions=caps)
    got = np.array([[r[k] for k in oracle.METRIC_KEYS] for r in res])
    assert np.array_equal(got, z["boot_metrics"])
    # the same RandomState(42) stream reproduces the stored indices (main.py:588)
    rs = np.random.RandomState(42)
    assert np.array_equal(np.array([rs.choice(N, size=N, replace=True) for _ in range(len(si))]), si)
    assert np.array_equal(np.random.RandomState(42).randint(0, N, size=(len(si), N)), si)
    # multiplicity identity == gather + rescore + topk, in integers
    s_t2i, s_i2t = torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"])
    thr_t, _, _ = oracle.above_gt_counts(s_t2i, [[i] for i in range(N)])
    thr_i, _, _ = oracle.above_gt_counts(s_i2t, [list(range(5 * i, 5 * i + 5)) for i in range(N)])
    thr_s, _, _ = oracle.above_gt_counts(s_t2i.t(), [[i] for i in range(N)])
    sets = {
        "T2I": ((s_t2i > thr_t[:, None]).numpy(), np.arange(N), np.arange(N)),
        "I2T": ((s_i2t > thr_i[:, None]).numpy(), np.arange(5 * N) // 5, np.arange(N)),
        "I2T_Sym": ((s_t2i.t() > thr_s[:, None]).numpy(), np.arange(N), np.arange(N)),
    }
    correct = oracle.bootstrap_identity(sets, si, N)
    assert np.array_equal(100.0 * correct / N, z["boot_metrics"])


def test_colpali_bootstrap_literal_loop(golden_dir):
    """The ColPali branch of main.py:586-634 (lists, per-iteration re-chunking: tie class T6), as computed by the
    reference's own compute_similarity_scores / compute_metrics_detailed (tests/golden/make_golden_colpali_boot.py):
    the oracle's literal loop reproduces it, and so does the multiplicity identity evaluated with the BASE chunking's
    padding — on a problem where 1573 (image, caption) scores depend on the padded zero row."""
    z = np.load(golden_dir / "colpali_boot.npz")
    docs = list(from_bits(z["docs"], 128).split([int(l) for l in z["doc_lens"]]))
    caps = list(from_bits(z["caps"], 128).split([int(l) for l in z["cap_lens"]]))
    n, si = len(docs), z["sample_indices"]
    keys = [str(k) for k in z["metric_keys"]]
    assert keys == oracle.METRIC_KEYS
    assert np.array_equal(np.random.RandomState(42).randint(0, n, size=si.shape), si)
    per = oracle.bootstrap_rescore_reference(docs, caps[0::5], caps, len(si), colpali=True, sample_indices=si)
    assert np.array_equal(np.array([[m[k] for k in keys] for m in per]), z["boot_metrics"])
    s_t2i, s_i2t = oracle.maxsim_scores(caps[0::5], docs), oracle.maxsim_scores(docs, caps)
    assert np.array_equal(s_t2i.numpy(), z["scores_t2i"]) and np.array_equal(s_i2t.numpy(), z["scores_i2t"])
    sets = {}
    for name, s, gt, cu in (("T2I", s_t2i, [[i] for i in range(n)], np.arange(n)),
                            ("I2T", s_i2t, [list(range(5 * i, 5 * i + 5)) for i in range(n)], np.arange(5 * n) // 5),
                            ("I2T_Sym", s_t2i.t().contiguous(), [[i] for i in range(n)], np.arange(n))):
        thr, _, tied = oracle.above_gt_counts(s, gt)
        assert int(tied.sum()) == 0
        sets[name] = ((s > thr[:, None]).numpy(), cu, np.arange(n))
    assert np.array_equal(100.0 * oracle.bootstrap_identity(sets, si, n) / n, z["boot_metrics"])


def test_ci_and_std(golden_dir):
    z = np.load(golden_dir / "stats.npz")
    data = z["ci_data"]
    rs = np.random.RandomState(int(z["ci_seed"]))
    mean, lo, hi = oracle.bootstrap_confidence_interval(data, rng=rs)
    assert (mean, lo, hi) == (float(z["ci_mean"]), float(z["ci_lower"]), float(z["ci_upper"]))
    # explicit numpy-order restatement: pairwise means of the gathered rows + linear percentile
    idx = oracle.ci_resample_indices(len(data), np.random.RandomState(int(z["ci_seed"])))
    means = np.array([oracle.np_mean(data[idx[r]]) for r in range(0, 10000, 97)])
    assert np.array_equal(means, np.array([np.mean(data[idx[r]]) for r in range(0, 10000, 97)]))
    all_means = data[idx].mean(axis=1)   # row-wise reduce uses the same pairwise routine for contiguous rows
    alpha = (1 - 0.95) / 2
    assert oracle.percentile_linear(all_means, alpha * 100) == float(z["ci_lower"])
    assert oracle.percentile_linear(all_means, (1 - alpha) * 100) == float(z["ci_upper"])
    assert oracle.np_mean(data) == float(z["ci_mean"])
    assert oracle.np_std(data) == float(z["ci_std"])


def test_permutation(golden_dir):
    z = np.load(golden_dir / "stats.npz")
    p = oracle.permutation_test(z["perm_d1"], z["perm_d2"], int(z["perm_n"]), rng=np.random.RandomState(int(z["perm_seed"])))
    assert p == float(z["perm_p"])
    # index-stream restatement: combined_after_p == combined0[perm[p]]
    d1, d2 = z["perm_b1"], z["perm_b2"]
    P = int(z["perm2_n"])
    perm = oracle.permutation_indices(len(d1) + len(d2), np.random.RandomState(int(z["perm2_seed"])), P)
    comb = np.concatenate([d1, d2])
    obs = np.abs(np.mean(d1) - np.mean(d2))
    diffs = np.array([abs(oracle.np_mean(comb[perm[i]][:len(d1)]) - oracle.np_mean(comb[perm[i]][len(d1):])) for i in range(P)])
    assert np.mean(diffs >= obs) == float(z["perm2_p"])


def test_pairwise_sum_matches_numpy():
    rs = np.random.RandomState(0)
    for _ in range(300):
        a = rs.rand(rs.randint(1, 3000)) * 100
        assert oracle.np_mean(a) == np.mean(a)
        assert oracle.np_std(a) == np.std(a)
    x = rs.rand(10000)
    for q in (2.500000000000002, 97.5, 0.0, 100.0, 50.0):
        assert oracle.percentile_linear(x, q) == np.percentile(x, q)


def test_maxsim_golden(golden_dir):
    z = np.load(golden_dir / "maxsim.npz")
    docs = list(from_bits(z["docs"], 128).split([int(l) for l in z["doc_lens"]]))
    queries = list(from_bits(z["queries"], 128).split([int(l) for l in z["query_lens"]]))
    s = oracle.maxsim_scores(queries, docs)
    assert np.array_equal(s.numpy(), z["scores"])
    s2 = oracle.maxsim_scores(docs, queries)
    assert np.array_equal(s2.numpy(), z["scores_swapped"])
    # the all-positive query against the all-negative doc: its padded zero row wins every max -> exactly 0
    flags = oracle.maxsim_pad_flags(z["doc_lens"])
    assert flags[4] == 1 and s[2, 4] == 0.0


def test_maxsim_known_answer():
    """Hand-computed 2 queries x 2 docs, ragged, d = 2."""
    q0 = torch.tensor([[1.0, 0.0], [0.0, 1.0]])
    q1 = torch.tensor([[-1.0, 0.0]])
    d0 = torch.tensor([[1.0, 0.0], [0.5, 0.5], [0.0, 2.0]])
    d1 = torch.tensor([[1.0, 1.0]])                      # padded to 3 rows with zeros inside the chunk
    s = oracle.maxsim_scores([q0, q1], [d0, d1])
    # q0/d0: max(1,.5,0)=1 + max(0,.5,2)=2 -> 3 ; q0/d1: max(1,0,0)+max(1,0,0)=2
    # q1/d0: max(-1,-.5,0)=0 ; q1/d1: max(-1, 0(pad), 0(pad)) = 0  <- padding wins over the real -1
    assert s.tolist() == [[3.0, 2.0], [0.0, 0.0]]

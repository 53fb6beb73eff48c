"""Generate tests/golden/*.npz by running the REFERENCE's own functions (imported from /root/reference/main.py).

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
The reference is imported from a scratch cwd because its import creates benchmark_v29.log. Inputs are seeded
synthetic embeddings; bf16 tensors are stored as their uint16 bit patterns. The ColPali `processor.score` stand-in
is the in-container HF port of colpali_engine's expression (ColPaliProcessor.score_retrieval) applied to .float()
inputs and driven through the reference's own chunk loop (main.py:451-470).
"""
import importlib.util
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def load_reference(name="main.py", modname="ref_main"):
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="mmx_ref_")
    os.chdir(scratch)
    argv = sys.argv
    sys.argv = [name]
    try:
        spec = importlib.util.spec_from_file_location(modname, str(REF / name))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
        sys.argv = argv
    mod.DEVICE = "cpu"
    mod.DTYPE = torch.bfloat16
    return mod


def bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


class ScoreStandIn:
    """processor.score == colpali_engine score_multi_vector; HF port evaluated on fp32-upcast inputs."""

    def score(self, qs, ps):
        from transformers.models.colpali.processing_colpali import ColPaliProcessor
        return ColPaliProcessor.score_retrieval(None, [q.float() for q in qs], [p.float() for p in ps],
                                                batch_size=128, output_dtype=torch.float32)


def make_dense(ref):
    g = torch.Generator().manual_seed(0)
    N, D, sigma = 60, 64, 2.5
    img = unit(torch.randn(N, D, generator=g)).bfloat16()
    noise = unit(torch.randn(N * 5, D, generator=g))
    all_txt = unit(img.float().repeat_interleave(5, 0) + sigma * noise).bfloat16()
    t2i_txt = all_txt[0::5].contiguous()
    m_info = {"type": "dense", "name": "golden"}
    s_t2i = ref.compute_similarity_scores(img, t2i_txt, m_info, None)
    s_i2t = ref.compute_similarity_scores(all_txt, img, m_info, None)   # [N, 5N]; D1 = the reference's extra .t()
    q2i = list(range(N))
    i2c = {i: list(range(5 * i, 5 * i + 5)) for i in range(N)}
    i2s = {i: i for i in range(N)}
    caps = [f"a red cat near two dogs number {i}" if i % 3 == 0 else f"plain caption {i}" for i in range(N)]
    m = ref.compute_metrics_detailed(s_t2i, s_i2t.contiguous(), s_t2i.t(), q2i, i2c, i2s, caps)
    keys = [k for k in m if not k.startswith("_")]
    base = np.array([m[k] for k in keys])
    correct_r1 = np.array([r["correct_r1"] for r in m["_query_results"]], dtype=np.uint8)
    # bootstrap iterations through the reference functions (main.py:586-634 with D1 corrected)
    B = 8
    rs = np.random.RandomState(42)
    si_all, per_iter = [], []
    for _ in range(B):
        si = rs.choice(N, size=N, replace=True)
        si_all.append(si)
        s_img, s_t2 = img[si], t2i_txt[si]
        cap_idx = [c for o in si for c in range(o * 5, o * 5 + 5)]
        s_all = all_txt[cap_idx]
        st2i = ref.compute_similarity_scores(s_img, s_t2, m_info, None)
        si2t = ref.compute_similarity_scores(s_all, s_img, m_info, None)
        mm = ref.compute_metrics_detailed(st2i, si2t, st2i.t(), list(range(N)),
                                          {i: list(range(5 * i, 5 * i + 5)) for i in range(N)},
                                          {i: i for i in range(N)}, [caps[i] for i in si])
        per_iter.append([mm[k] for k in keys])
    np.savez_compressed(HERE / "dense.npz", img=bits(img), all_txt=bits(all_txt), scores_t2i=s_t2i.numpy(),
                        scores_i2t=s_i2t.contiguous().numpy(), metric_keys=np.array(keys), base_metrics=base,
                        correct_r1=correct_r1, sample_indices=np.array(si_all), boot_metrics=np.array(per_iter),
                        captions=np.array(caps))
    fa = ref.aggregate_failure_analysis([m["_query_results"]])
    return keys, fa


def make_stats(ref):
    rs = np.random.RandomState(5)
    data = np.round(rs.normal(41.5, 0.7, size=1000) * 50) / 50.0   # looks like 100*correct/5000 values
    np.random.seed(123)
    mean, lo, hi = ref.bootstrap_confidence_interval(data)
    std = np.std(data)
    d1 = np.round(rs.normal(41.5, 0.7, size=300) * 50) / 50.0
    d2 = np.round(rs.normal(41.6, 0.7, size=300) * 50) / 50.0
    np.random.seed(7)
    p = ref.permutation_test(d1, d2, n_permutations=2000)
    b1 = (rs.rand(500) < 0.4).astype(np.float64)
    b2 = (rs.rand(700) < 0.45).astype(np.float64)
    np.random.seed(8)
    p2 = ref.permutation_test(b1, b2, n_permutations=1000)
    np.savez_compressed(HERE / "stats.npz", ci_data=data, ci_seed=123, ci_mean=mean, ci_lower=lo, ci_upper=hi,
                        ci_std=std, perm_d1=d1, perm_d2=d2, perm_seed=7, perm_n=2000, perm_p=p, perm_b1=b1, perm_b2=b2,
                        perm2_seed=8, perm2_n=1000, perm2_p=p2)


def make_maxsim(ref):
    g = torch.Generator().manual_seed(3)
    docs = [unit(torch.randn(int(l), 128, generator=g)).bfloat16()
            for l in torch.randint(1, 41, (23,), generator=g)]
    docs[4] = -docs[4].abs()              # all-negative patches: whenever its chunk pads it, the zero row wins
    queries = []
    for i, l in enumerate(torch.randint(3, 13, (17,), generator=g)):
        q = unit(torch.randn(int(l), 128, generator=g)).bfloat16()
        if i == 2:
            q = q.abs()                   # all-positive tokens vs the all-negative doc -> every sim < 0
        if i == 5:
            q = torch.cat([q, torch.zeros(3, 128, dtype=torch.bfloat16)])   # zero-padded query rows (encode batch)
        queries.append(q)
    m_info = {"type": "colpali", "name": "golden"}
    proc = ScoreStandIn()
    s_t2i = ref.compute_similarity_scores(docs, queries, m_info, proc)      # [n_q, n_docs]
    s_swap = ref.compute_similarity_scores(queries, docs, m_info, proc)     # roles swapped: [n_docs, n_q]
    np.savez_compressed(HERE / "maxsim.npz",
                        doc_lens=np.array([d.shape[0] for d in docs]), query_lens=np.array([q.shape[0] for q in queries]),
                        docs=bits(torch.cat(docs)), queries=bits(torch.cat(queries)),
                        scores=s_t2i.numpy(), scores_swapped=s_swap.numpy())


def make_v28(keys_unused):
    try:
        v28 = load_reference("run_benchmark_grand_slam_v28_publication_ready.py", "ref_v28")
    except BaseException as e:  # noqa: BLE001 - the v28 script may sys.exit on import
        print("v28 import failed:", repr(e))
        return
    z = np.load(HERE / "dense.npz")
    s_t2i, s_i2t = torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"])
    N = s_t2i.shape[0]
    m = v28.compute_metrics(s_t2i, s_i2t, list(range(N)), {i: list(range(5 * i, 5 * i + 5)) for i in range(N)})
    np.savez_compressed(HERE / "v28_metrics.npz", keys=np.array(list(m.keys())), values=np.array(list(m.values())))


if __name__ == "__main__":
    ref = load_reference()
    keys, fa = make_dense(ref)
    import json
    (HERE / "failure_analysis.json").write_text(json.dumps(fa, indent=1, sort_keys=True))
    make_stats(ref)
    make_maxsim(ref)
    make_v28(keys)
    print("golden vectors written to", HERE)

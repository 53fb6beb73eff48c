"""Generate tests/golden/colpali_boot.npz: the ColPali branch of the reference's bootstrap loop (main.py:586-634) on a
small ragged problem, evaluated by the REFERENCE's own functions imported from /root/reference/main.py —
``compute_similarity_scores`` (colpali branch, chunk loop main.py:451-470, with the in-container HF port of
colpali_engine's ``score_multi_vector`` as ``processor.score``) and ``compute_metrics_detailed``.

Run in the build container only:  python tests/golden/make_golden_colpali_boot.py

The loop around them is main.py:586-634 with the two shipped defects corrected the way SURVEY 5.2 documents (D1: the I2T
matrix is images x captions, images being the MaxSim queries as in v28 :375-386; D2: the embeddings stay python lists).
The reference re-chunks the RESAMPLED lists in every iteration, so which captions get a zero pad row depends on the
neighbours that were drawn (tie class T6): this file pins what the reference computes in that case. 42 images (chunks of
10 do not divide it), 4-24-token captions, 12 iterations, captions noisy enough (sigma 5) that Recall@1 is ~30 % and one patch row in
sixteen has no positive similarity with a four-token caption. The seed is the one of 40 whose base score matrices keep
every competitor farthest (relatively) from its row's ground-truth threshold — above 1e-4, two orders of magnitude
beyond the summation-order differences between a GPU MMA and the CPU's sgemm — so the integers must reproduce.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
from make_golden import ScoreStandIn, bits, load_reference  # noqa: E402
from helpers import synth_colpali  # noqa: E402

N_IMG, DOC_LEN, B, SIGMA_Q = 42, 130, 12, 5.0


def margins(s, gt):
    """smallest relative distance between a row's GT threshold and any non-GT score"""
    worst = 1.0
    for i, g in enumerate(gt):
        thr = s[i, g].max()
        others = np.delete(s[i], g)
        worst = min(worst, float(np.min(np.abs(others - thr)) / abs(thr)))
    return worst


def build(seed):
    docs, _ = synth_colpali(N_IMG, N_IMG, doc_len=DOC_LEN, q_len=(8, 8), seed=seed)
    _, all_caps = synth_colpali(N_IMG, 5 * N_IMG, doc_len=DOC_LEN, q_len=(4, 24), sigma_q=SIGMA_Q, seed=seed)
    return docs, all_caps


def main():
    ref = load_reference()
    proc = ScoreStandIn()
    m_info = {"type": "colpali", "name": "golden"}
    best = (-1.0, None)
    for seed in range(4, 44):
        docs, all_caps = build(seed)
        t2i = all_caps[0::5]
        s_t2i = ref.compute_similarity_scores(docs, t2i, m_info, proc)                 # [n_txt, n_img]
        s_i2t = ref.compute_similarity_scores(all_caps, docs, m_info, proc)            # images as queries: [n_img, 5 n_img]
        m = min(margins(s_t2i.numpy(), [[i] for i in range(N_IMG)]),
                margins(s_i2t.numpy(), [list(range(5 * i, 5 * i + 5)) for i in range(N_IMG)]),
                margins(s_t2i.t().numpy(), [[i] for i in range(N_IMG)]))
        print("seed", seed, "min relative margin %.2e" % m, flush=True)
        if m > best[0]:
            best = (m, seed)
    m, seed = best
    assert m > 1e-4, "no seed keeps the competitors 1e-4 away from the thresholds"
    docs, all_caps = build(seed)
    t2i = all_caps[0::5]
    s_t2i = ref.compute_similarity_scores(docs, t2i, m_info, proc)
    s_i2t = ref.compute_similarity_scores(all_caps, docs, m_info, proc)
    caps = [f"a red cat near two dogs number {i}" if i % 3 == 0 else f"plain caption {i}" for i in range(N_IMG)]
    q2i = list(range(N_IMG))
    i2c = {i: list(range(5 * i, 5 * i + 5)) for i in range(N_IMG)}
    i2s = {i: i for i in range(N_IMG)}
    base = ref.compute_metrics_detailed(s_t2i, s_i2t, s_t2i.t(), q2i, i2c, i2s, caps)
    keys = [k for k in base if not k.startswith("_")]
    rs = np.random.RandomState(42)
    si_all, per_iter, worst = [], [], 1.0
    for it in range(B):
        si = rs.choice(N_IMG, size=N_IMG, replace=True)                                # main.py:588
        si_all.append(si)
        s_img = [docs[i] for i in si]                                                  # main.py:592-593 (D2: lists)
        s_t2 = [t2i[i] for i in si]
        cap_idx = [c for o in si for c in range(o * 5, o * 5 + 5)]                     # main.py:602-609
        s_all = [all_caps[c] for c in cap_idx]
        st2i = ref.compute_similarity_scores(s_img, s_t2, m_info, proc)                # main.py:620
        si2t = ref.compute_similarity_scores(s_all, s_img, m_info, proc)               # main.py:622 with D1 corrected
        mm = ref.compute_metrics_detailed(st2i, si2t, st2i.t(), q2i, i2c, i2s, [caps[i] for i in si])
        per_iter.append([mm[k] for k in keys])
        print("iteration", it, per_iter[-1], flush=True)
    np.savez_compressed(
        HERE / "colpali_boot.npz", seed=seed, doc_lens=np.array([d.shape[0] for d in docs]),
        cap_lens=np.array([c.shape[0] for c in all_caps]), docs=bits(torch.cat(docs)), caps=bits(torch.cat(all_caps)),
        scores_t2i=s_t2i.numpy(), scores_i2t=s_i2t.numpy(), metric_keys=np.array(keys),
        base_metrics=np.array([base[k] for k in keys]), sample_indices=np.array(si_all),
        boot_metrics=np.array(per_iter), min_margin=m)
    print("written", HERE / "colpali_boot.npz")


if __name__ == "__main__":
    main()

"""Oracle: legacy metric variants (TEST INFRASTRUCTURE — see oracle/__init__.py).

``compute_bidirectional_metrics`` follows /root/reference/run_benchmark_grand_slam_v19_fixed.py:110-191 literally
(topk membership, argsort rank, Python-float accumulation, string formatting); ``winoground_counts`` follows the
per-example comparisons of run_benchmark_grand_slam_v28_publication_ready.py:449-458.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


def compute_bidirectional_metrics(scores_t2i: torch.Tensor, scores_i2t: torch.Tensor,
                                  query_to_image_map: Optional[List[int]] = None,
                                  image_to_all_captions: Optional[Dict[int, List[int]]] = None) -> Dict[str, str]:
    n_text, n_img = scores_t2i.size(0), scores_i2t.size(0)
    if query_to_image_map is None:
        query_to_image_map = list(range(n_text))                                   # v19 :133-134
    if image_to_all_captions is None:
        image_to_all_captions = {i: [i] for i in range(n_img)}                     # v19 :137-141
    out: Dict[str, str] = {}
    for k in (1, 5, 10):                                                           # v19 :144-155
        if k > scores_t2i.size(1):
            out[f"T2I_R@{k}"] = "N/A"
            continue
        hit = sum(query_to_image_map[i] in torch.topk(scores_t2i[i], k=k).indices.tolist() for i in range(n_text))
        out[f"T2I_R@{k}"] = f"{hit / n_text:.1%}"
    acc = 0                                                                        # v19 :158-166
    for i in range(n_text):
        order = scores_t2i[i].argsort(descending=True)
        acc += 1.0 / ((order == query_to_image_map[i]).nonzero(as_tuple=True)[0].item() + 1)
    out["T2I_MRR"] = f"{acc / n_text:.3f}"
    for k in (1, 5, 10):                                                           # v19 :169-180
        if k > scores_i2t.size(1):
            out[f"I2T_R@{k}"] = "N/A"
            continue
        hit = 0
        for i in range(n_img):
            top = torch.topk(scores_i2t[i], k=k).indices.tolist()
            hit += any(c in top for c in image_to_all_captions[i])
        out[f"I2T_R@{k}"] = f"{hit / n_img:.1%}"
    acc = 0                                                                        # v19 :183-191
    for i in range(n_img):
        valid = set(image_to_all_captions[i])
        for rank, idx in enumerate(scores_i2t[i].argsort(descending=True).tolist(), 1):
            if idx in valid:
                acc += 1.0 / rank
                break
    out["I2T_MRR"] = f"{acc / n_img:.3f}"
    return out


def winoground_counts(blocks: torch.Tensor) -> Dict[str, float]:
    """blocks [n, 2, 2] = s[caption, image] per example; v28 :449-458."""
    text = image = group = 0
    total = blocks.shape[0]
    for s in blocks.float():
        t = bool(s[0, 0] > s[0, 1] and s[1, 1] > s[1, 0])
        im = bool(s[0, 0] > s[1, 0] and s[1, 1] > s[0, 1])
        text += t
        image += im
        group += t and im
    return {"Wino Text": 100.0 * text / total, "Wino Image": 100.0 * image / total, "Wino Group": 100.0 * group / total}

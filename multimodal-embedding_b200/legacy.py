"""Legacy metric variants on the B200 path: MRR and Winoground.

Both fall out of the strict-greater counts the ranking kernels already produce:
  * ``compute_bidirectional_metrics`` — run_benchmark_grand_slam_v19_fixed.py:110-191 (same signature, same string
    formatting): Recall@K as ``f"{correct/n:.1%}"``, MRR as ``f"{sum/n:.3f}"`` with rank = 1 + #{columns scoring
    strictly above the best ground-truth column} (KR's ``cnt_above``) — the position ``argsort(descending=True)``
    gives outside exact fp32 ties (tie class T2);
  * ``winoground_scores`` — the three counters of run_benchmark_grand_slam_v28_publication_ready.py:449-452 over
    per-example 2x2 score blocks, and ``winoground_from_embeddings`` which scores all examples with one K1 / K2 launch.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch

from . import ops
from .api import _as_bf16, _as_list, _device, _flatten, _h2d, _recall_block


def _mrr(cnt_above: torch.Tensor) -> float:
    """sum_i 1.0 / rank_i accumulated left to right in Python-float order (v19 :160-166), divided by n."""
    rank = cnt_above.cpu().numpy().astype(np.int64) + 1
    terms = 1.0 / rank.astype(np.float64)
    total = float(np.add.accumulate(terms)[-1]) if len(terms) else 0.0      # accumulate = sequential fp64 adds
    return total / len(terms)


def compute_bidirectional_metrics(txt_embeds, img_embeds, scores_t2i: torch.Tensor, scores_i2t: torch.Tensor,
                                  query_to_image_map: Optional[List[int]] = None,
                                  image_to_query_map: Optional[List[int]] = None,
                                  image_to_all_captions: Optional[Dict[int, List[int]]] = None) -> Dict[str, str]:
    """run_benchmark_grand_slam_v19_fixed.py:110-191. ``txt_embeds`` / ``img_embeds`` are unused there too."""
    dev = _device()
    n_text, n_img = scores_t2i.size(0), scores_i2t.size(0)
    if query_to_image_map is None:
        query_to_image_map = list(range(n_text))
    if image_to_all_captions is None:
        if image_to_query_map is None:
            image_to_query_map = list(range(n_img))
        image_to_all_captions = {i: [image_to_query_map[i]] for i in range(n_img)}
    r_t2i = _recall_block(scores_t2i, [[query_to_image_map[i]] for i in range(n_text)], dev)
    r_i2t = _recall_block(scores_i2t, [image_to_all_captions[i] for i in range(n_img)], dev)
    metrics: Dict[str, str] = {}
    for name, r, n, width in (("T2I", r_t2i, n_text, scores_t2i.size(1)), ("I2T", r_i2t, n_img, scores_i2t.size(1))):
        cnt = r.cnt_above.cpu().numpy()
        for k in (1, 5, 10):
            if k > width:
                metrics[f"{name}_R@{k}"] = "N/A"
            else:
                metrics[f"{name}_R@{k}"] = f"{int((cnt < k).sum()) / n:.1%}"
        metrics[f"{name}_MRR"] = f"{_mrr(r.cnt_above):.3f}"
    # the reference emits T2I_R@*, T2I_MRR, I2T_R@*, I2T_MRR in this order (dict insertion order above matches)
    return metrics


def winoground_scores(s: Union[torch.Tensor, np.ndarray]) -> Dict[str, float]:
    """v28 :449-458 on a stack of per-example score blocks s[e, caption, image] (fp32 [n, 2, 2])."""
    s = np.asarray(s.detach().cpu().float() if isinstance(s, torch.Tensor) else s, dtype=np.float32)
    if s.ndim != 3 or s.shape[1:] != (2, 2):
        raise ValueError("winoground_scores takes [n, 2, 2] score blocks (caption x image)")
    total = s.shape[0]
    text = (s[:, 0, 0] > s[:, 0, 1]) & (s[:, 1, 1] > s[:, 1, 0])
    image = (s[:, 0, 0] > s[:, 1, 0]) & (s[:, 1, 1] > s[:, 0, 1])
    return {"Wino Text": 100.0 * int(text.sum()) / total, "Wino Image": 100.0 * int(image.sum()) / total,
            "Wino Group": 100.0 * int((text & image).sum()) / total}


def winoground_from_embeddings(img_embeds, txt_embeds, m_info: Dict[str, Any]) -> Dict[str, float]:
    """All Winoground examples in one launch. Dense: ``img_embeds`` / ``txt_embeds`` bf16 [2n, d] (example e owns rows
    2e, 2e+1). ColPali: lists of 2n multi-vector items. The 2x2 block of example e is read off the diagonal of the
    [2n, 2n] score matrix (text rows x image columns), then counted exactly like the reference's per-example loop."""
    dev = _device()
    if m_info["type"] == "colpali":
        q_flat, q_off, _ = _flatten(_as_list(txt_embeds), dev)
        d_flat, d_off, _ = _flatten(_as_list(img_embeds), dev)
        # the reference scores each example's two images together: pad_sequence pads the shorter image of the pair
        lens = (d_off[1:] - d_off[:-1]).view(-1, 2)
        d_zero = (lens < lens.max(dim=1, keepdim=True).values).reshape(-1).to(torch.uint8)
        S = ops.maxsim_scores(q_flat, q_off, d_flat, d_off, d_zero if bool(d_zero.any()) else None)
    else:
        S = ops.dense_scores(_h2d(_as_bf16(txt_embeds, "txt_embeds"), dev), _h2d(_as_bf16(img_embeds, "img_embeds"), dev))
    n2 = S.shape[0]
    if n2 % 2 or S.shape[1] != n2:
        raise ValueError("Winoground needs two captions and two images per example")
    e = torch.arange(n2 // 2, device=dev) * 2
    blocks = torch.stack([torch.stack([S[e, e], S[e, e + 1]], dim=1), torch.stack([S[e + 1, e], S[e + 1, e + 1]], dim=1)],
                         dim=1)
    return winoground_scores(blocks)

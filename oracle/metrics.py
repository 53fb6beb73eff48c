"""Oracle: Recall@K metrics (TEST INFRASTRUCTURE — see oracle/__init__.py).

``compute_metrics_detailed`` follows /root/reference/main.py:305-383 line by line (per-row ``torch.topk`` +
membership); ``compute_metrics`` follows run_benchmark_grand_slam_v28_publication_ready.py:203-229.
``above_gt_counts`` / ``recall_from_counts`` are the strict-greater restatement the kernels implement:
outside exact fp32 ties both give the same integers, which tests/test_oracle_golden.py checks.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

KS = (1, 5, 10)
METRIC_KEYS = [f"{p}_R@{k}" for p in ("T2I", "I2T", "I2T_Sym") for k in KS]  # main.py:358,368,378 insertion order

# main.py:102-125
COCO_SUPERCATEGORIES = {
    "person": ["person"],
    "vehicle": ["bicycle", "car", "motorcycle", "airplane", "bus", "train", "truck", "boat"],
    "outdoor": ["traffic light", "fire hydrant", "stop sign", "parking meter", "bench"],
    "animal": ["bird", "cat", "dog", "horse", "sheep", "cow", "elephant", "bear", "zebra", "giraffe"],
    "accessory": ["backpack", "umbrella", "handbag", "tie", "suitcase"],
    "sports": ["frisbee", "skis", "snowboard", "sports ball", "kite", "baseball bat", "baseball glove",
               "skateboard", "surfboard", "tennis racket"],
    "kitchen": ["bottle", "wine glass", "cup", "fork", "knife", "spoon", "bowl"],
    "food": ["banana", "apple", "sandwich", "orange", "broccoli", "carrot", "hot dog", "pizza", "donut", "cake"],
    "furniture": ["chair", "couch", "potted plant", "bed", "dining table", "toilet"],
    "electronic": ["tv", "laptop", "mouse", "remote", "keyboard", "cell phone"],
    "appliance": ["microwave", "oven", "toaster", "sink", "refrigerator"],
    "indoor": ["book", "clock", "vase", "scissors", "teddy bear", "hair drier", "toothbrush"],
}
SPATIAL_KEYWORDS = ["left", "right", "top", "bottom", "above", "below", "next to", "beside", "in front", "behind",
                    "between", "near", "far"]
COLOR_KEYWORDS = ["red", "blue", "green", "yellow", "black", "white", "orange", "purple", "pink", "brown", "gray",
                  "grey"]
COUNTING_KEYWORDS = ["one", "two", "three", "four", "five", "1", "2", "3", "4", "5", "single", "double", "triple",
                     "multiple", "several", "many", "few"]


def analyze_query_complexity(caption: str) -> Dict[str, Any]:
    """main.py:239-247."""
    low = caption.lower()
    return {
        "has_spatial": any(kw in low for kw in SPATIAL_KEYWORDS),
        "has_color": any(kw in low for kw in COLOR_KEYWORDS),
        "has_counting": any(kw in low for kw in COUNTING_KEYWORDS),
        "length": len(caption.split()),
    }


def get_category_from_caption(caption: str) -> Optional[str]:
    """main.py:249-256."""
    low = caption.lower()
    for sup, kws in COCO_SUPERCATEGORIES.items():
        for kw in kws:
            if kw in low:
                return sup
    return None


def compute_metrics_detailed(scores_t2i, scores_i2t, scores_i2t_symmetric, query_to_img: List[int],
                             img_to_caps: Dict[int, List[int]], img_to_single_cap: Dict[int, int],
                             captions: List[str]) -> Dict[str, Any]:
    """main.py:305-383, literal (torch.topk per row). Small sizes only: 45 000 topk calls at N=5000."""
    n_text = scores_t2i.size(0)
    n_img = scores_i2t.size(0)
    metrics: Dict[str, Any] = {}
    query_results = []
    for k in KS:  # main.py:331-358
        correct = 0
        for i in range(n_text):
            target = query_to_img[i]
            top_k = torch.topk(scores_t2i[i], k=min(k, n_img)).indices.tolist()
            ok = target in top_k
            correct += int(ok)
            if k == 1:
                cap = captions[i] if i < len(captions) else ""
                cx = analyze_query_complexity(cap)
                query_results.append({
                    "query_idx": i, "caption": cap, "correct_r1": ok, "has_spatial": cx["has_spatial"],
                    "has_color": cx["has_color"], "has_counting": cx["has_counting"],
                    "caption_length": cx["length"], "category": get_category_from_caption(cap),
                })
        metrics[f"T2I_R@{k}"] = 100.0 * correct / n_text
    for k in KS:  # main.py:361-368
        correct = 0
        for i in range(n_img):
            valid = img_to_caps[i]
            top_k = torch.topk(scores_i2t[i], k=min(k, scores_i2t.size(1))).indices.tolist()
            correct += int(any(c in top_k for c in valid))
        metrics[f"I2T_R@{k}"] = 100.0 * correct / n_img
    for k in KS:  # main.py:371-378
        correct = 0
        for i in range(n_img):
            target = img_to_single_cap[i]
            top_k = torch.topk(scores_i2t_symmetric[i], k=min(k, scores_i2t_symmetric.size(1))).indices.tolist()
            correct += int(target in top_k)
        metrics[f"I2T_Sym_R@{k}"] = 100.0 * correct / n_img
    metrics["_query_results"] = query_results
    return metrics


def compute_metrics(scores_t2i, scores_i2t, query_to_img, img_to_caps) -> Dict[str, float]:
    """run_benchmark_grand_slam_v28_publication_ready.py:203-229, literal."""
    n_text = scores_t2i.size(0)
    n_img = scores_i2t.size(0)
    metrics: Dict[str, float] = {}
    for k in KS:
        correct = 0
        for i in range(n_text):
            top_k = torch.topk(scores_t2i[i], k=min(k, n_img)).indices.tolist()
            correct += int(query_to_img[i] in top_k)
        metrics[f"T2I_R@{k}"] = 100.0 * correct / n_text
    for k in KS:
        correct = 0
        for i in range(n_img):
            top_k = torch.topk(scores_i2t[i], k=min(k, n_text)).indices.tolist()
            correct += int(any(c in top_k for c in img_to_caps[i]))
        metrics[f"I2T_R@{k}"] = 100.0 * correct / n_img
    return metrics


def gt_csr(gt_sets: Sequence[Sequence[int]]):
    """Ground-truth sets -> CSR (offsets int32 [n+1], indices int32)."""
    off = np.zeros(len(gt_sets) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(g) for g in gt_sets])
    idx = np.fromiter((c for g in gt_sets for c in g), dtype=np.int32, count=int(off[-1]))
    return off, idx


def above_gt_counts(scores: torch.Tensor, gt_sets: Sequence[Sequence[int]]):
    """Strict-greater restatement of main.py:331-378: thr_i = max_{c in GT(i)} S[i,c];
    cnt_i = #{j : S[i,j] > thr_i}. Query i is correct at k iff cnt_i < k (no exact ties).
    Returns (thr fp32 [n], cnt int64 [n], n_tied int64 [n] = #{j not in GT(i) : S[i,j] == thr_i})."""
    s = scores.float()
    n = s.shape[0]
    thr = torch.empty(n, dtype=torch.float32)
    tied = torch.zeros(n, dtype=torch.int64)
    for i in range(n):
        g = torch.as_tensor(list(gt_sets[i]), dtype=torch.long)
        thr[i] = s[i, g].max()
        eq = (s[i] == thr[i])
        eq[g] = False
        tied[i] = int(eq.sum())
    cnt = (s > thr[:, None]).sum(dim=1)
    return thr, cnt.to(torch.int64), tied


def recall_from_counts(cnt, n: Optional[int] = None) -> Dict[int, float]:
    """metric = 100.0 * correct / n (main.py:358) with correct_k = #{cnt < k}."""
    cnt = np.asarray(cnt)
    n = len(cnt) if n is None else n
    return {k: 100.0 * int((cnt < k).sum()) / n for k in KS}


def topk_lists(scores: torch.Tensor, k: int = 10):
    """`torch.topk(scores[i], k)` of main.py:333,362,372 for every row, with the tie order the CUDA path fixes
    (higher score first, equal scores by lower column): a stable descending sort. Rows shorter than k are padded with
    (-inf, -1). Returns (values fp32 [n, k], indices int64 [n, k])."""
    s = scores.float()
    n, g = s.shape
    vals, idx = torch.sort(s, dim=1, descending=True, stable=True)
    kk = min(k, g)
    top_v = torch.full((n, k), float("-inf"))
    top_i = torch.full((n, k), -1, dtype=torch.int64)
    top_v[:, :kk] = vals[:, :kk]
    top_i[:, :kk] = idx[:, :kk]
    return top_v, top_i

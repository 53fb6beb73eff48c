"""Oracle: similarity scoring (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/main.py:440-475 (``compute_similarity_scores``) and, for the ColPali branch, the
third-party ``colpali_engine`` ``score_multi_vector`` expression that ``processor.score`` (main.py:467) reaches.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence

CHUNK = 10  # main.py:453 chunk_size


def dense_scores(img_embeds: torch.Tensor, txt_embeds: torch.Tensor) -> torch.Tensor:
    """main.py:471-473: ``torch.matmul(txt.float(), img.float().t())`` -> fp32 [n_txt, n_img] on CPU."""
    return torch.matmul(txt_embeds.float(), img_embeds.float().t())


def _score_multi_vector_fp32(qs: Sequence[torch.Tensor], ps: Sequence[torch.Tensor]) -> torch.Tensor:
    """colpali_engine score_multi_vector on ONE batch (a 10x10 chunk never exceeds its batch_size=128), evaluated
    on .float() inputs ("fp32-accumulate"): pad both sides with zeros, einsum, max over doc rows, sum over
    query rows. Zero padding is NOT masked: a padded doc row offers 0 to every max."""
    q = pad_sequence([t.float() for t in qs], batch_first=True, padding_value=0)
    p = pad_sequence([t.float() for t in ps], batch_first=True, padding_value=0)
    return torch.einsum("bnd,csd->bcns", q, p).max(dim=3)[0].sum(dim=2)


def maxsim_scores(q_list: List[torch.Tensor], d_list: List[torch.Tensor], chunk_size: int = CHUNK) -> torch.Tensor:
    """main.py:451-470 (and v28 :360-386 with roles swapped): 10x10 chunk loop over lists of [L_i, 128] tensors.

    Returns fp32 [len(q_list), len(d_list)]."""
    rows = []
    for i in range(0, len(q_list), chunk_size):
        q_chunk = q_list[i:i + chunk_size]
        row = []
        for j in range(0, len(d_list), chunk_size):
            row.append(_score_multi_vector_fp32(q_chunk, d_list[j:j + chunk_size]))
        rows.append(torch.cat(row, dim=1))
    return torch.cat(rows, dim=0)


def maxsim_pad_flags(d_lens: Sequence[int], chunk_size: int = CHUNK) -> np.ndarray:
    """1 where pad_sequence gives the document at least one zero row inside its chunk of `chunk_size`
    (len < longest doc of the chunk) — that zero then competes in the max (main.py:465-467)."""
    d_lens = np.asarray(d_lens, dtype=np.int64)
    flags = np.zeros(len(d_lens), dtype=np.uint8)
    for j in range(0, len(d_lens), chunk_size):
        c = d_lens[j:j + chunk_size]
        flags[j:j + chunk_size] = (c < c.max()).astype(np.uint8)
    return flags


def flatten_multivector(items: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, np.ndarray]:
    """List of [L_i, d] -> flat [sum L, d] plus int32 CSR offsets [n+1]."""
    lens = np.array([int(t.shape[0]) for t in items], dtype=np.int64)
    off = np.zeros(len(items) + 1, dtype=np.int32)
    off[1:] = np.cumsum(lens)
    flat = torch.cat([t.reshape(-1, t.shape[-1]) for t in items], dim=0) if len(items) else torch.empty(0, 128)
    return flat, off

"""Shared helpers for the test-suite (synthetic data per SURVEY 8d, bf16 <-> uint16 bit patterns)."""
import numpy as np
import torch


def from_bits(a: np.ndarray, cols: int) -> torch.Tensor:
    return torch.from_numpy(a.astype(np.uint16).view(np.int16).copy()).view(torch.bfloat16).reshape(-1, cols)


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def synth_dense(n_img: int, dim: int, sigma: float, seed: int = 0):
    """img = unit(randn); all_txt[5i+c] = unit(img[i] + sigma * unit(randn)); bf16."""
    g = torch.Generator().manual_seed(seed)
    img = unit(torch.randn(n_img, dim, generator=g))
    noise = unit(torch.randn(n_img * 5, dim, generator=g))
    all_txt = unit(img.repeat_interleave(5, 0) + sigma * noise)
    return img.bfloat16(), all_txt.bfloat16()


def synth_colpali(n_docs: int, n_queries: int, doc_len=1030, q_len=(8, 40), sigma_q: float = 1.0, seed: int = 2,
                  ragged_docs=None):
    """docs = unit(randn(L,128)); query tokens = noisy copies of patches of the GT doc (q // (n_queries//n_docs))."""
    g = torch.Generator().manual_seed(seed)
    docs = []
    for j in range(n_docs):
        L = doc_len if ragged_docs is None else int(torch.randint(ragged_docs[0], ragged_docs[1] + 1, (1,), generator=g))
        docs.append(unit(torch.randn(L, 128, generator=g)).bfloat16())
    per = max(1, n_queries // n_docs)
    queries = []
    for i in range(n_queries):
        j = min(i // per, n_docs - 1)
        L = int(torch.randint(q_len[0], q_len[1] + 1, (1,), generator=g))
        rows = torch.randint(0, docs[j].shape[0], (L,), generator=g)
        q = unit(docs[j][rows].float() + sigma_q * unit(torch.randn(L, 128, generator=g)))
        queries.append(q.bfloat16())
    return docs, queries

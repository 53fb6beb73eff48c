"""Host side of the drop-in boundary: the reference's own Python signatures, backed by libmmx.so.

Same names, argument meaning, return types and error behaviour (ordinary Python exceptions, never sys.exit) as
/root/reference/main.py: ``compute_similarity_scores`` (:440), ``compute_metrics_detailed`` (:305),
``bootstrap_confidence_interval`` (:259), ``permutation_test`` (:281), ``run_bootstrap_benchmark`` (:478) and v28's
``compute_metrics`` (run_benchmark_grand_slam_v28_publication_ready.py:203). Inputs are the caller's CPU tensors,
outputs are CPU tensors / Python floats; the work happens on the current CUDA device through libmmx.

Plus the fused fast path the harness times: ``score_and_rank`` -> device-resident rank structures without
a score matrix, ``bootstrap_from_ranks`` -> per-iteration integer counts, ``run_bootstrap_from_embeddings`` -> CSV row.

Random numbers: exactly like the reference these functions consume numpy's legacy MT19937 stream — the GLOBAL
``np.random`` state unless an explicit ``RandomState`` is passed — in the reference's order, in bulk
(``randint(0, n, (R, n))`` is the same stream as R calls of ``choice``).
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, ops
from .failure import (aggregate_failure_analysis, aggregate_failure_from_hits, analyze_query_complexity,
                      caption_feature_table,
                      get_category_from_caption)

KS = (1, 5, 10)
METRIC_KEYS = [f"{p}_R@{k}" for p in ("T2I", "I2T", "I2T_Sym") for k in KS]     # main.py:358,368,378
CSV_COLUMNS = (["Model"] + [f"{k}_{s}" for k in METRIC_KEYS for s in ("mean", "lower", "upper", "std")]
               + ["Time", "QPS", "Encoding_Time", "Img_per_sec", "_failure_analysis"])  # main.py:643-665
N_BOOTSTRAP = 10000       # main.py:265
CHUNK = 10                # main.py:453
DEVICE_SHUFFLE_MAX_N = 65536  # longest array whose shuffles are replayed on the device (K6s: table scheme up to 16 384, one-CTA bracket scheme beyond)
DEVICE_RNG = True         # draw the reference's numpy MT19937 stream on the GPU (K6); False = numpy draws on the host


# ------------------------------------------------------------------------------------------------ plumbing
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.MmxError("no CUDA device: the scoring path has no CPU fallback")
    _lib.lib()  # fail loudly if libmmx.so was not built
    return torch.device("cuda", torch.cuda.current_device())


_AUX_STREAMS: Dict[Tuple[int, str], "torch.cuda.Stream"] = {}


def _aux_stream(dev: torch.device, kind: str) -> "torch.cuda.Stream":
    """One long-lived side stream per (device, purpose). A fresh torch.cuda.Stream per call walks through torch's
    pool of 32 streams, and the caching allocator keeps a separate block pool per stream, so every call paid
    cudaMalloc for its buffers again (measured: 70-440 ms of host stall per call on 1.4 GB)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), kind)
    st = _AUX_STREAMS.get(key)
    if st is None:
        st = _AUX_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


def _h2d(t: torch.Tensor, dev: torch.device, dtype=None) -> torch.Tensor:
    """Host tensor -> device through pinned memory (async on the current stream)."""
    if t.is_cuda:
        return t if dtype is None or t.dtype == dtype else t.to(dtype)
    t = t.contiguous()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_pinned():
        p = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        p.copy_(t)
        t = p
    return t.to(dev, non_blocking=True)


def _host_read(t: torch.Tensor) -> torch.Tensor:
    """Device->host copy that first waits on the STREAM. A copy into pageable memory that has to wait for its stream
    waits inside the driver and holds up the launches of every other host thread (ops.ShufflePrefetch's worker, which
    replays the permutation tests' shuffles next to a sharded scoring pass, lost half its time to this); a stream wait
    does not."""
    torch.cuda.current_stream().synchronize()
    return t.cpu()


def _np_h2d(a: np.ndarray, dev: torch.device, dtype: torch.dtype) -> torch.Tensor:
    return _h2d(torch.from_numpy(np.ascontiguousarray(a)), dev, dtype)


def _as_bf16(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{what}: expected torch.bfloat16 embeddings (the reference's DTYPE on a bf16 GPU, "
                        f"main.py:93), got {t.dtype}")
    return t


def _as_list(x: Union[torch.Tensor, Sequence[torch.Tensor]]) -> List[torch.Tensor]:
    """ColPali side as a list of [L_i, 128] tensors (v28 semantics; a padded [n, L, 128] tensor is split per item)."""
    if isinstance(x, torch.Tensor):
        if x.dim() != 3:
            raise ValueError("ColPali embeddings must be a list of [L_i, 128] tensors or one padded [n, L, 128] tensor")
        return [x[i] for i in range(x.size(0))]
    out = []
    for t in x:
        out.append(t[0] if t.dim() == 3 and t.size(0) == 1 else t)
    return out


def _flatten(items: List[torch.Tensor], dev: torch.device) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    lens = np.array([int(t.shape[0]) for t in items], dtype=np.int64)
    off = np.zeros(len(items) + 1, dtype=np.int32)
    off[1:] = np.cumsum(lens)
    if not items:
        flat = torch.empty(0, 128, dtype=torch.bfloat16)
    elif items[0].is_cuda:
        flat = torch.cat([_as_bf16(t, "ColPali embeddings") for t in items], dim=0)
    else:
        # the reference hands over 30 000 small CPU tensors: gather them straight into ONE pinned buffer (a single
        # pass over the 1.4 GB instead of cat + pin), from where the upload is asynchronous
        for t in items:
            _as_bf16(t, "ColPali embeddings")
        flat = torch.empty((int(off[-1]), int(items[0].shape[1])), dtype=torch.bfloat16, pin_memory=True)
        torch.cat(list(items), dim=0, out=flat)
    return _h2d(flat, dev), _np_h2d(off, dev, torch.int32), lens


def _flatten_any(x, dev: torch.device) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    """Lists / padded tensors (the reference's representations) or a FlatMultiVector (no 30 000-tensor torch.cat)."""
    if isinstance(x, FlatMultiVector):
        flat, off = x.to_device(dev)
        return flat, off, x.lens
    return _flatten(_as_list(x), dev)


def _pad_flags(lens: np.ndarray, chunk: int = CHUNK) -> np.ndarray:
    """1 where pad_sequence adds a zero row to the document inside its chunk of 10 (main.py:465-467)."""
    lens = np.asarray(lens, dtype=np.int64)
    n = len(lens)
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    full = -(-n // chunk) * chunk
    padded = np.full(full, -1, dtype=np.int64)          # the last chunk may be short: fillers never win the max
    padded[:n] = lens
    cmax = padded.reshape(-1, chunk).max(axis=1)
    return (lens < np.repeat(cmax, chunk)[:n]).astype(np.uint8)


def _gt_dev(gt_sets: Sequence[Sequence[int]], dev: torch.device):
    off, idx = ops.csr_from_sets(gt_sets)
    return _np_h2d(off, dev, torch.int32), _np_h2d(idx, dev, torch.int32)


def _identity_gt(n: int, dev: torch.device):
    return (torch.arange(n + 1, dtype=torch.int32, device=dev), torch.arange(n, dtype=torch.int32, device=dev))


def _block_gt(n: int, per: int, dev: torch.device):
    return (torch.arange(0, per * n + 1, per, dtype=torch.int32, device=dev),
            torch.arange(per * n, dtype=torch.int32, device=dev))


# ------------------------------------------------------------------------------------------------ scoring
def maxsim_scores_device(q_items: List[torch.Tensor], d_items: List[torch.Tensor], dev: torch.device) -> torch.Tensor:
    q_flat, q_off, _ = _flatten(q_items, dev)
    d_flat, d_off, d_lens = _flatten(d_items, dev)
    flags = _pad_flags(d_lens)
    d_zero = _np_h2d(flags, dev, torch.uint8) if flags.any() else None
    return ops.maxsim_scores(q_flat, q_off, d_flat, d_off, d_zero)


def compute_similarity_scores(img_embeds: Union[torch.Tensor, List[torch.Tensor]],
                              txt_embeds: Union[torch.Tensor, List[torch.Tensor]],
                              m_info: Dict, processor: Any = None) -> torch.Tensor:
    """main.py:440-475. Returns fp32 scores [n_txt, n_img] on the CPU.

    ColPali: txt_embeds are the MaxSim queries, img_embeds the documents (main.py:462-467); `processor` is unused —
    its `.score` arithmetic (colpali_engine score_multi_vector, fp32 accumulate) runs in the K2 kernel.
    Dense: txt.float() @ img.float().t() on tcgen05 (bf16 products are exact in fp32)."""
    dev = _device()
    if m_info["type"] == "colpali":
        scores = maxsim_scores_device(_as_list(txt_embeds), _as_list(img_embeds), dev)
    else:
        Q = _h2d(_as_bf16(txt_embeds, "txt_embeds"), dev)
        G = _h2d(_as_bf16(img_embeds, "img_embeds"), dev)
        scores = ops.dense_scores(Q, G)
    return scores.cpu()


# ------------------------------------------------------------------------------------------------ metrics
def _recall_block(scores: torch.Tensor, gt_sets: Sequence[Sequence[int]], dev: torch.device) -> ops.ProtocolRanks:
    s = _h2d(scores if scores.dtype == torch.float32 else scores.float(), dev)
    off, idx = _gt_dev(gt_sets, dev)
    return ops.rank_from_scores(s, off, idx, want_bits=False)


def compute_metrics_detailed(scores_t2i: torch.Tensor, scores_i2t: torch.Tensor, scores_i2t_symmetric: torch.Tensor,
                             query_to_img: List[int], img_to_caps: Dict[int, List[int]],
                             img_to_single_cap: Dict[int, int], captions: List[str]) -> Dict[str, Any]:
    """main.py:305-383. Same nine ``{T2I,I2T,I2T_Sym}_R@{1,5,10}`` floats (= 100.0*correct/n) and the same
    ``_query_results`` records; a query is correct at k iff fewer than k columns score strictly above its best
    ground-truth column (identical to torch.topk membership outside exact fp32 ties)."""
    dev = _device()
    n_text = scores_t2i.size(0)
    n_img = scores_i2t.size(0)
    r_t2i = _recall_block(scores_t2i, [[query_to_img[i]] for i in range(n_text)], dev)
    r_i2t = _recall_block(scores_i2t, [img_to_caps[i] for i in range(n_img)], dev)
    r_sym = _recall_block(scores_i2t_symmetric, [[img_to_single_cap[i]] for i in range(n_img)], dev)
    metrics: Dict[str, Any] = {}
    for name, r in (("T2I", r_t2i), ("I2T", r_i2t), ("I2T_Sym", r_sym)):
        for k, v in r.recall().items():
            metrics[f"{name}_R@{k}"] = v
    correct_r1 = (r_t2i.cnt_above < 1).cpu().tolist()
    query_results = []
    for i in range(n_text):                                       # main.py:342-356
        caption = captions[i] if i < len(captions) else ""
        cx = analyze_query_complexity(caption)
        query_results.append({
            "query_idx": i, "caption": caption, "correct_r1": bool(correct_r1[i]), "has_spatial": cx["has_spatial"],
            "has_color": cx["has_color"], "has_counting": cx["has_counting"], "caption_length": cx["length"],
            "category": get_category_from_caption(caption),
        })
    metrics["_query_results"] = query_results
    return metrics


def compute_metrics(scores_t2i: torch.Tensor, scores_i2t: torch.Tensor, query_to_img: List[int],
                    img_to_caps: Dict[int, List[int]]) -> Dict[str, float]:
    """run_benchmark_grand_slam_v28_publication_ready.py:203-229 (T2I + I2T, six keys)."""
    dev = _device()
    n_text, n_img = scores_t2i.size(0), scores_i2t.size(0)
    r_t2i = _recall_block(scores_t2i, [[query_to_img[i]] for i in range(n_text)], dev)
    r_i2t = _recall_block(scores_i2t, [img_to_caps[i] for i in range(n_img)], dev)
    metrics: Dict[str, float] = {}
    for name, r in (("T2I", r_t2i), ("I2T", r_i2t)):
        for k, v in r.recall().items():
            metrics[f"{name}_R@{k}"] = v
    return metrics


# ------------------------------------------------------------------------------------------------ statistics
def _rng(rng):
    return np.random if rng is None else rng


def bootstrap_confidence_interval(data: np.ndarray, confidence: float = 0.95, rng=None) -> Tuple[float, float, float]:
    """main.py:259-278. Consumes the same numpy stream as the reference's 10 000 ``np.random.choice(data, n)`` calls;
    the 10 000 resample means are computed on the device in numpy's fp64 pairwise order, the two percentiles by
    ``np.percentile`` itself on those identical values. Returns (mean, lower, upper)."""
    dev = _device()
    data = np.asarray(data)
    n = len(data)
    d = _np_h2d(data.astype(np.float64, copy=False), dev, torch.float64)
    if isinstance(rng, ops.DeviceRng):
        idx = rng.randint(n, N_BOOTSTRAP * n).view(N_BOOTSTRAP, n)           # K6: the same stream, drawn on the GPU
    elif DEVICE_RNG:
        drng = ops.DeviceRng(_rng(rng), dev)
        idx = drng.randint(n, N_BOOTSTRAP * n).view(N_BOOTSTRAP, n)
        drng.commit()                                                        # numpy's state advances as in the reference
    else:
        idx = _np_h2d(_rng(rng).randint(0, n, size=(N_BOOTSTRAP, n)).astype(np.int32), dev, torch.int32)
    means = ops.ci_means(d, idx).cpu().numpy()                               # == N_BOOTSTRAP x choice + mean
    mean = np.mean(data)
    alpha = (1 - confidence) / 2
    lower = np.percentile(means, alpha * 100)
    upper = np.percentile(means, (1 - alpha) * 100)
    return mean, lower, upper


def skip_bootstrap_draws(drng: "ops.DeviceRng", n_units: int, n_iterations: int, n_ci_sets: int = len(KS)) -> None:
    """Advance a device generator past the numpy draws of ONE run_retrieval_bootstrap call (n_iterations x choice(n, n), then
    n_ci_sets x 10 000 x choice(values, n_iterations)) without keeping them — for ops.ShufflePrefetch, which replays the
    permutation tests that FOLLOW the models' bootstraps in the stream while the models are still being scored."""
    drng.randint(n_units, n_iterations * n_units, window=(0, 0))
    drng.randint(n_iterations, n_ci_sets * N_BOOTSTRAP * n_iterations, window=(0, 0))


def permutation_test(data1: np.ndarray, data2: np.ndarray, n_permutations: int = 10000, rng=None,
                     index_rows: Optional[torch.Tensor] = None) -> float:
    """main.py:281-302. The cumulative in-place shuffles of ``combined`` are replayed on an index array with the
    same numpy stream; the 2 x n_permutations means and |diff| run on the device in numpy's summation order.
    ``index_rows`` (int32 CUDA [n_permutations, n], ops.ShufflePrefetch) supplies shuffles that were replayed ahead of
    time from this very stream position; the generator is then left alone (ShufflePrefetch.commit moves it)."""
    dev = _device()
    data1, data2 = np.asarray(data1, dtype=np.float64), np.asarray(data2, dtype=np.float64)
    observed_diff = np.abs(np.mean(data1) - np.mean(data2))
    combined = np.concatenate([data1, data2])
    n1, n = len(data1), len(data1) + len(data2)
    r = _rng(rng)
    if index_rows is not None:
        if tuple(index_rows.shape) != (n_permutations, n):
            raise ValueError("index_rows must be [n_permutations, len(data1) + len(data2)]")
        perm_d = index_rows
    elif DEVICE_RNG and 2 <= n <= DEVICE_SHUFFLE_MAX_N and n_permutations >= 1:
        # K6s: the shuffles themselves are replayed on the device from numpy's stream; numpy's generator is moved
        # past exactly the words they consumed
        drng = r if isinstance(r, ops.DeviceRng) else ops.DeviceRng(r, dev)
        perm_d = drng.shuffle_indices(n, n_permutations)
        if drng is not r:
            drng.commit()
    else:
        if isinstance(r, ops.DeviceRng):
            raise ValueError("arrays this long are shuffled on the host: pass the numpy generator, not a DeviceRng")
        idx = np.arange(n, dtype=np.int64)
        perm = np.empty((n_permutations, n), dtype=np.int32)
        for p in range(n_permutations):
            r.shuffle(idx)                                                   # main.py:296
            perm[p] = idx
        perm_d = _np_h2d(perm, dev, torch.int32)
    diffs = ops.perm_stats(_np_h2d(combined, dev, torch.float64), n1, perm_d).cpu().numpy()
    return np.mean(diffs >= observed_diff)


# ------------------------------------------------------------------------------------------------ fused path
def score_and_rank(img_embeds, t2i_txt_embeds, all_txt_embeds, m_info: Dict, caps_per_img: int = 5,
                   want_bits: bool = True, want_top10: bool = False, want_ties: bool = False) -> Dict[str, ops.ProtocolRanks]:
    """One scoring pass over the BASE embeddings for the three protocols of run_bootstrap_benchmark
    (main.py:599-624 GT structure: query i <-> image i, image i <-> captions [5i, 5i+5)).

    Dense models never materialise a score matrix (K1). ColPali writes the final [n_q, n_doc] fp32 scores (one float
    per 5.3 MFLOP pair) and ranks them with KR; I2T uses images as MaxSim queries (v28 :375-386) and I2T_Sym the
    transposed T2I matrix exactly as main.py:624. ``want_ties`` fills ``ties`` with the tie report of SURVEY 8a (T2):
    per query, the non-ground-truth columns that score exactly the threshold (``ops.tie_summary`` counts the queries
    whose Recall@k depends on torch.topk's unspecified order among them). ``want_top10`` also fills ``top_s`` / ``top_i`` of every protocol with
    the retrieved items (`torch.topk(row, 10)` of main.py:333,362,372) — still without a dense score matrix."""
    dev = _device()
    out: Dict[str, ops.ProtocolRanks] = {}
    if m_info["type"] == "colpali":
        n = len(img_embeds)
        a_flat, a_off, a_lens = _flatten_any(all_txt_embeds, dev)
        if isinstance(img_embeds, FlatMultiVector) and isinstance(t2i_txt_embeds, FlatMultiVector):
            # host-resident images: their upload overlaps the T2I scoring, the device copy is reused for I2T
            kept: dict = {}
            s_t2i = _maxsim_streamed(t2i_txt_embeds, img_embeds, dev, keep=kept)
            (i_flat, i_off), i_lens = kept["d"], img_embeds.lens
            t_flat, t_off = kept["q"]
        else:
            i_flat, i_off, i_lens = _flatten_any(img_embeds, dev)
            t_flat, t_off, t_lens = _flatten_any(t2i_txt_embeds, dev)
            fl_i = _pad_flags(i_lens)
            s_t2i = ops.maxsim_scores(t_flat, t_off, i_flat, i_off, _np_h2d(fl_i, dev, torch.uint8) if fl_i.any() else None)
        fl_a = _pad_flags(a_lens)
        s_i2t = ops.maxsim_scores(i_flat, i_off, a_flat, a_off, _np_h2d(fl_a, dev, torch.uint8) if fl_a.any() else None)
        id_off, id_idx = _identity_gt(n, dev)
        b_off, b_idx = _block_gt(n, caps_per_img, dev)
        out["T2I"] = ops.rank_from_scores(s_t2i, id_off, id_idx, want_bits)
        out["I2T"] = ops.rank_from_scores(s_i2t, b_off, b_idx, want_bits)
        s_sym = s_t2i.t().contiguous()
        out["I2T_Sym"] = ops.rank_from_scores(s_sym, id_off, id_idx, want_bits)
        if want_top10:
            for name, sc in (("T2I", s_t2i), ("I2T", s_i2t), ("I2T_Sym", s_sym)):
                out[name].top_s, out[name].top_i = ops.topk_from_scores(sc, 10)
        if want_ties:
            for name, sc, (go, gi) in (("T2I", s_t2i, (id_off, id_idx)), ("I2T", s_i2t, (b_off, b_idx)),
                                       ("I2T_Sym", s_sym, (id_off, id_idx))):
                out[name].ties = ops.exact_tie_counts(out[name], go, gi, scores=sc)
            # tie class T6: where the zero row of pad_sequence (main.py:465) can win a maximum. The reference
            # re-chunks the RESAMPLED lists in every bootstrap iteration, so a document's padding depends on the
            # neighbours that were drawn; the bootstrap here keeps the base chunking's flags
            out["T2I"].t6 = _t6_report(t_flat, t_off, i_flat, i_off, _pad_flags(i_lens), s_t2i, dev)
            out["I2T"].t6 = _t6_report(i_flat, i_off, a_flat, a_off, fl_a, s_i2t, dev)
            out["I2T_Sym"].t6 = dict(out["T2I"].t6, note="transposed T2I matrix (main.py:624): the T2I pairs, seen per image")
    else:
        img = _h2d(_as_bf16(img_embeds, "img_embeds"), dev)
        t2i = _h2d(_as_bf16(t2i_txt_embeds, "t2i_txt_embeds"), dev)
        allc = _h2d(_as_bf16(all_txt_embeds, "all_txt_embeds"), dev)
        n = img.shape[0]
        id_off, id_idx = _identity_gt(n, dev)
        b_off, b_idx = _block_gt(n, caps_per_img, dev)
        tk = 10 if want_top10 else None
        out["T2I"] = ops.dense_rank(t2i, img, id_off, id_idx, want_bits, top_k=tk)       # [n_txt, n_img]   main.py:620
        out["I2T"] = ops.dense_rank(img, allc, b_off, b_idx, want_bits, top_k=tk)        # [n_img, 5 n_img] main.py:622 (D1 fixed)
        out["I2T_Sym"] = ops.dense_rank(img, t2i, id_off, id_idx, want_bits, top_k=tk)   # = scores_t2i.t() main.py:624
        if want_ties:
            for name, (qq, gg), (go, gi) in (("T2I", (t2i, img), (id_off, id_idx)), ("I2T", (img, allc), (b_off, b_idx)),
                                             ("I2T_Sym", (img, t2i), (id_off, id_idx))):
                out[name].ties = ops.exact_tie_counts(out[name], go, gi, Q=qq, G=gg)
    out["I2T"].col2unit = torch.arange(n * caps_per_img, dtype=torch.int32, device=dev) // caps_per_img
    return out


def _t6_report(q_flat, q_off, d_flat, d_off, flags: np.ndarray, scores: torch.Tensor, dev: torch.device) -> Dict[str, int]:
    """Tie class T6 as numbers. A (query, document) score depends on the document's zero-padding flag iff some query row
    has no positive similarity with any row of the document (the padded zero then wins that row's maximum).
    Two more K2 passes — every document padded / no document padded — find those pairs:
      pairs_sensitive_to_padding    pairs whose score differs between the two (a re-chunked bootstrap iteration of the
                                    reference, main.py:602-624, can move exactly these);
      queries_sensitive_to_padding  queries (rows) with at least one such pair;
      queries_with_clamped_maximum  rows that hold a clamped maximum against a document the BASE chunking pads — the
                                    scores this library's bootstrap uses for every iteration;
      documents_padded              documents the base chunking pads."""
    nd = int(d_off.numel()) - 1
    ones = torch.ones(nd, dtype=torch.uint8, device=dev)
    s_pad = ops.maxsim_scores(q_flat, q_off, d_flat, d_off, ones)
    s_raw = ops.maxsim_scores(q_flat, q_off, d_flat, d_off, None)
    sens = s_pad != s_raw
    base = (scores != s_raw) & sens                  # only padded documents can differ from the unpadded pass
    return {"pairs_sensitive_to_padding": int(sens.sum().item()),
            "queries_sensitive_to_padding": int(sens.any(dim=1).sum().item()),
            "queries_with_clamped_maximum": int(base.any(dim=1).sum().item()),
            "documents_padded": int(np.asarray(flags).sum())}


def base_metrics(ranks: Dict[str, ops.ProtocolRanks]) -> Dict[str, float]:
    """The nine Recall@K values of the un-resampled pass, keyed like compute_metrics_detailed."""
    m: Dict[str, float] = {}
    for name in ("T2I", "I2T", "I2T_Sym"):
        for k, v in ranks[name].recall().items():
            m[f"{name}_R@{k}"] = v
    return m


def bootstrap_from_ranks(ranks: Dict[str, ops.ProtocolRanks], sample_indices: np.ndarray, n_images: int,
                         want_hits: bool = False):
    """K3: int64 [B, 9] correct counts (METRIC_KEYS order) for the reference's resample draws
    ``sample_indices`` [B, N] (main.py:588), without rescoring. With want_hits also returns, per base T2I query,
    (R@1 hits summed over iterations, times drawn) — the inputs of aggregate_failure_analysis."""
    correct, hits, drawn = _bootstrap_counts_device(ranks, sample_indices, n_images, want_hits)
    correct = correct.cpu().numpy().astype(np.int64)
    if want_hits:
        return correct, hits.cpu().numpy().astype(np.int64), drawn.cpu().numpy()
    return correct


def _bootstrap_counts_device(ranks: Dict[str, ops.ProtocolRanks], sample_indices, n_images: int, want_hits: bool):
    """bootstrap_from_ranks without the device->host copies: (int32 CUDA [B, 9], hits int32 [nq] | None,
    drawn int64 [nq] | None)."""
    dev = _device()
    if isinstance(sample_indices, torch.Tensor) and sample_indices.is_cuda:
        si = sample_indices.to(torch.int32)
    else:
        si = _np_h2d(np.asarray(sample_indices).astype(np.int32), dev, torch.int32)
    mult = ops.bincount(si, n_images)
    cols = []
    hits = torch.zeros(ranks["T2I"].nq, dtype=torch.int32, device=dev) if want_hits else None
    for name in ("T2I", "I2T", "I2T_Sym"):
        cols.append(ops.boot_counts(ranks[name], mult, q_hits1=hits if name == "T2I" else None))
    drawn = None
    if want_hits:
        drawn = mult.sum(dim=0, dtype=torch.int64)
        qu = ranks["T2I"].q_unit
        drawn = drawn if qu is None else drawn[qu.long()]
    return torch.cat(cols, dim=1), hits, drawn


def run_bootstrap_from_embeddings(img_embeds, t2i_txt_embeds, all_txt_embeds, m_info: Dict, n_iterations: int,
                                  rng=None, t2i_captions: Optional[List[str]] = None, encoding_time: float = 0.0,
                                  t_start: Optional[float] = None, tie_report: bool = False) -> Dict[str, Any]:
    """Everything run_bootstrap_benchmark does after encoding (main.py:563-667) -> the 42-column row dict.

    numpy stream consumption matches the reference: n_iterations x choice(N, N), then for each of the nine metric
    keys in insertion order 10 000 x choice(values, n_iterations). ``tie_report`` adds ``_tie_report``: per protocol, how
    many base queries have non-ground-truth columns tied bit for bit with their threshold (tie class T2)."""
    t0 = time.time() if t_start is None else t_start
    n_images = len(img_embeds)
    aggregated: Dict[str, Any] = {"Model": m_info["name"]}
    drng = None
    if DEVICE_RNG:
        # The draws do not depend on the scores: the whole numpy stream of this row — n_iterations x choice(N, N), then
        # nine times 10 000 x choice(values, n_iterations), consecutive in the stream — is regenerated on a side stream
        # (K6p) while the embeddings cross PCIe and K1 runs; the nine CI kernels then run back to back on device-resident
        # values and ONE device->host copy brings everything back.
        dev = _device()
        main = torch.cuda.current_stream()
        side = _aux_stream(dev, "rng")
        side.wait_stream(main)
        with torch.cuda.stream(side):
            drng = ops.DeviceRng(_rng(rng), dev)
            sample_indices = drng.randint(n_images, n_iterations * n_images).view(n_iterations, n_images)  # main.py:588
            ci_idx = drng.randint(n_iterations, len(METRIC_KEYS) * N_BOOTSTRAP * n_iterations).view(
                len(METRIC_KEYS), N_BOOTSTRAP, n_iterations)                                            # main.py:269
            rng_done = torch.cuda.Event()
            rng_done.record(side)
        ranks = score_and_rank(img_embeds, t2i_txt_embeds, all_txt_embeds, m_info, want_ties=tie_report)
        main.wait_event(rng_done)
        for t in (sample_indices, ci_idx):
            t.record_stream(main)
        correct_d, hits_d, drawn_d = _bootstrap_counts_device(ranks, sample_indices, n_images, want_hits=True)
        n_t = torch.full((), float(n_images), dtype=torch.float64, device=dev)   # a fill kernel, not a blocking host copy
        values_d = 100.0 * correct_d.to(torch.float64) / n_t          # tensor / tensor: IEEE division, as numpy's
        done = torch.cuda.Event()
        done.record(main)
        means_d = ops.ci_means_batched(values_d.t().contiguous(), ci_idx)          # nine CIs, one launch
        # host work that needs no result from the device, done while the device is still busy
        caps = t2i_captions if t2i_captions is not None else [""] * n_images
        caption_table = caption_feature_table(caps)
        done.synchronize()
        total_time = time.time() - t0                                                      # main.py:636
        values, means = values_d.cpu().numpy(), means_d.cpu().numpy()
        hits, drawn = hits_d.cpu().numpy().astype(np.int64), drawn_d.cpu().numpy()
        alpha = (1 - 0.95) / 2
        # one vectorised call = the same partition + linear interpolation per row as the reference's two scalar
        # np.percentile calls per metric (main.py:273-274), with the same (inexact) q values
        bounds = np.percentile(means, [alpha * 100, (1 - alpha) * 100], axis=1)
        for j, key in enumerate(METRIC_KEYS):                                              # main.py:644-651, :259-278
            v = np.ascontiguousarray(values[:, j])
            aggregated[f"{key}_mean"] = np.mean(v)
            aggregated[f"{key}_lower"] = bounds[0, j]
            aggregated[f"{key}_upper"] = bounds[1, j]
            aggregated[f"{key}_std"] = np.std(v)
    else:
        ranks = score_and_rank(img_embeds, t2i_txt_embeds, all_txt_embeds, m_info, want_ties=tie_report)
        sample_indices = _rng(rng).randint(0, n_images, size=(n_iterations, n_images))
        correct, hits, drawn = bootstrap_from_ranks(ranks, sample_indices, n_images, want_hits=True)
        values = 100.0 * correct / n_images                                                # main.py:358,368,378
        torch.cuda.synchronize()
        total_time = time.time() - t0                                                      # main.py:636
        for j, key in enumerate(METRIC_KEYS):                                              # main.py:644-651
            v = np.ascontiguousarray(values[:, j])
            mean, lower, upper = bootstrap_confidence_interval(v, rng=rng)
            aggregated[f"{key}_mean"] = mean
            aggregated[f"{key}_lower"] = lower
            aggregated[f"{key}_upper"] = upper
            aggregated[f"{key}_std"] = np.std(v)
    enc = encoding_time if encoding_time > 0 else float("nan")
    aggregated["Time"] = total_time
    aggregated["QPS"] = n_images / enc
    aggregated["Encoding_Time"] = encoding_time
    aggregated["Img_per_sec"] = n_images / enc
    if drng is not None:
        drng.commit()
    if drng is None:
        caps = t2i_captions if t2i_captions is not None else [""] * n_images
        caption_table = None
    aggregated["_failure_analysis"] = aggregate_failure_from_hits(                                      # main.py:664
        caps, hits, drawn, table=caption_table,
        first_seen_order=_first_seen_order(sample_indices, n_images, (caption_table or caption_feature_table(caps))[1]))
    aggregated["_bootstrap_values"] = {k: values[:, j].copy() for j, k in enumerate(METRIC_KEYS)}
    aggregated["_base_metrics"] = base_metrics(ranks)
    if tie_report:
        aggregated["_tie_report"] = {name: ops.tie_summary(ranks[name]) for name in ("T2I", "I2T", "I2T_Sym")}
        if m_info["type"] == "colpali":
            for name in ("T2I", "I2T", "I2T_Sym"):
                aggregated["_tie_report"][name]["T6"] = ranks[name].t6
    return aggregated


def _first_seen_order(sample_indices, n_images: int, cats) -> List[int]:
    """Queries that introduce a category, in the reference's record order (failure.category_first_seen). Only the leading
    iterations are read back from the device (all of them for small problems)."""
    from .failure import category_first_seen
    B = int(sample_indices.shape[0])
    rows = B if B * n_images <= (1 << 20) else min(B, 8)
    head = sample_indices[:rows]
    head = head.cpu().numpy() if isinstance(head, torch.Tensor) else np.asarray(head)
    return category_first_seen(head, cats)


class FlatMultiVector:
    """Embedding hand-off format for multi-vector (ColPali) items: one flat bf16 [sum(L_i), 128] tensor
    (CPU, pinned CPU or CUDA) plus the per-item lengths — what K2 consumes, without 30 000 small tensors."""

    def __init__(self, flat: torch.Tensor, lens):
        self.flat = flat
        self.lens = np.asarray(lens, dtype=np.int64)
        if int(self.lens.sum()) != flat.shape[0]:
            raise ValueError("lens do not add up to the number of rows")

    def __len__(self) -> int:
        return len(self.lens)

    @staticmethod
    def from_list(items: Sequence[torch.Tensor]) -> "FlatMultiVector":
        items = _as_list(items)
        return FlatMultiVector(torch.cat(list(items), dim=0), [int(t.shape[0]) for t in items])

    def slice(self, a: int, b: int) -> "FlatMultiVector":
        off = np.concatenate([[0], np.cumsum(self.lens)])
        return FlatMultiVector(self.flat[int(off[a]):int(off[b])], self.lens[a:b])

    def to_device(self, dev: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
        off = np.zeros(len(self.lens) + 1, dtype=np.int32)
        off[1:] = np.cumsum(self.lens)
        return _h2d(_as_bf16(self.flat, "multi-vector embeddings"), dev), _np_h2d(off, dev, torch.int32)

    def to_list(self) -> List[torch.Tensor]:
        """The reference's representation: a python list of [L_i, 128] tensors (views, no copy)."""
        return list(self.flat.split([int(l) for l in self.lens]))

    def save(self, path: str) -> None:
        """On-disk cache: the reference persists no embeddings, so every rerun of the statistics pays
        for the encoders again. One .npz: the bf16 rows as their uint16 bit patterns plus the lengths."""
        flat = _as_bf16(self.flat, "multi-vector embeddings").detach().cpu().contiguous()
        np.savez(path, bits=flat.view(torch.int16).numpy().view(np.uint16), lens=self.lens,
                 dim=np.int64(flat.shape[1] if flat.dim() == 2 else 0))

    @staticmethod
    def load(path: str, pin: bool = True) -> "FlatMultiVector":
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        flat = torch.from_numpy(z["bits"].view(np.int16).copy()).view(torch.bfloat16)
        if pin and torch.cuda.is_available():
            flat = flat.pin_memory()           # so the upload overlaps the scoring (_maxsim_streamed)
        return FlatMultiVector(flat, z["lens"])


H2D_DOC_CHUNKS = 8   # host-resident documents are uploaded in this many pieces, each scored while the next one copies


def _maxsim_streamed(qv: FlatMultiVector, dv: FlatMultiVector, dev: torch.device,
                     flags: Optional[np.ndarray] = None, keep: Optional[dict] = None, group=None) -> torch.Tensor:
    """K2 over documents that still live in host memory: the upload of document chunk c+1 (copy stream, pinned
    source) overlaps the scoring of chunk c, so PCIe time hides behind the tensor cores. Returns fp32 [nq, nd].
    `flags`: the documents' zero-padding flags when they are not those of this list taken alone (a shard).
    `keep`: filled with the device copies ({"q": (flat, off), "d": (flat, off)}) for callers that score again."""
    nq, nd = len(qv), len(dv)
    if flags is None:
        flags = _pad_flags(dv.lens)                   # chunk-of-10 padding flags of the WHOLE list (main.py:465)
    if group is not None and not qv.flat.is_cuda:
        # replicated queries: every rank uploads 1/N of them, one all-gather over NVLink spreads the slices
        from . import sharded
        src_q = _as_bf16(qv.flat, "multi-vector embeddings")
        q_flat = sharded.gather_rows(src_q if src_q.is_pinned() else src_q.pin_memory(), dev, group)
        off = np.zeros(len(qv.lens) + 1, dtype=np.int32)
        off[1:] = np.cumsum(qv.lens)
        q_off = _np_h2d(off, dev, torch.int32)
    else:
        q_flat, q_off = qv.to_device(dev)
    n_chunks = 1 if dv.flat.is_cuda else max(1, min(H2D_DOC_CHUNKS, nd // 64))
    if n_chunks == 1:
        d_flat, d_off = dv.to_device(dev)
        if keep is not None:
            keep["q"], keep["d"] = (q_flat, q_off), (d_flat, d_off)
        return ops.maxsim_scores(q_flat, q_off, d_flat, d_off, _np_h2d(flags, dev, torch.uint8) if flags.any() else None)
    S = torch.empty((nq, nd), dtype=torch.float32, device=dev)
    main, copy = torch.cuda.current_stream(), _aux_stream(dev, "copy")
    step = -(-nd // n_chunks)
    # one device buffer for all documents (allocated on the main stream, filled chunk by chunk on the copy stream):
    # every chunk is scored from its slice, and the whole matrix is there afterwards for the role-swapped pass
    src = _as_bf16(dv.flat, "multi-vector embeddings")
    if not src.is_pinned():
        pinned = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        pinned.copy_(src)
        src = pinned
    d_all = torch.empty(src.shape, dtype=torch.bfloat16, device=dev)
    off_h = np.zeros(nd + 1, dtype=np.int64)
    off_h[1:] = np.cumsum(dv.lens)
    d_off_all = _np_h2d(off_h.astype(np.int32), dev, torch.int32)
    d_zero_all = _np_h2d(flags, dev, torch.uint8) if flags.any() else None
    staged = []
    copy.wait_stream(main)
    with torch.cuda.stream(copy):
        for a in range(0, nd, step):
            b = min(nd, a + step)
            r0, r1 = int(off_h[a]), int(off_h[b])
            d_all[r0:r1].copy_(src[r0:r1], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
            staged.append((a, b, r0, r1, ready))
    for a, b, r0, r1, ready in staged:
        main.wait_event(ready)
        sub_off = (d_off_all[a:b + 1] - r0).contiguous()
        ops.maxsim_scores(q_flat, q_off, d_all[r0:r1], sub_off, None if d_zero_all is None else d_zero_all[a:b], out=S[:, a:b])
    if keep is not None:
        keep["q"], keep["d"] = (q_flat, q_off), (d_all, d_off_all)
    return S


BITMASK_LIMIT_BYTES = 16 << 30     # above-GT bitmasks larger than this switch the dense path to LIST mode
MAXSIM_SCORES_LIMIT = 2 << 30      # single-GPU MaxSim score matrices larger than this are never materialised (ops.maxsim_rank)


def run_retrieval_bootstrap(queries, docs, query_to_doc: Sequence[int], m_info: Dict, n_iterations: int, rng=None,
                            group=None, timings: Optional[Dict[str, float]] = None, list_mode: Optional[bool] = None,
                            list_cap: Optional[int] = None, shard_sizes: Optional[Sequence[int]] = None,
                            keep: Optional[dict] = None) -> Optional[Dict[str, Any]]:
    """One retrieval protocol (T2I-style: every query has one ground-truth document) scored, ranked, bootstrapped and
    aggregated like main.py:586-651 — for the configurations whose query set is not "caption 0 of every image"
    (BASELINE.json configs 4-5: 25 000 caption queries against 5 000 / 100 000 documents).

    queries / docs: bf16 tensors [n, d] (dense) or FlatMultiVector / lists of [L_i, 128] (ColPali). The bootstrap unit
    is the document: an iteration draws ``choice(n_docs, n_docs)`` documents, each draw brings along its queries.
    With a torch.distributed ``group`` (world > 1) every rank passes ITS contiguous shard of the documents
    (``sharded.shard_bounds``) and all queries; rank 0 returns the result, the others None. ``shard_sizes`` (documents
    per rank, known to callers that made the shards) saves the collective that otherwise asks for them.
    numpy stream: n_iterations x choice(n_docs, n_docs), then 3 x 10 000 x choice(values, n_iterations) — rank 0's
    generator; its state is broadcast so that every rank can take its share of the statistics.
    Collectives per call (world > 1): [sizes], [ColPali: document lengths], generator state (2.5 KB), queries
    (host-resident: 1/N uploaded per rank, all-gathered over NVLink), thresholds all-reduce(MAX), ONE all-gather of the
    bit columns, and two small all-gathers that reassemble the bootstrap counts / CI means computed in blocks.
    ``keep`` (dict) receives the merged rank structures (``ranks``) for callers that go on (per-query vectors).

    Dense galleries whose [n_queries, n_docs] above-GT bitmask would exceed BITMASK_LIMIT_BYTES (a 10^6 x 10^6 problem
    needs 125 GB of bits) run in LIST mode (``list_mode`` forces it either way): K1 keeps at most
    ``list_cap`` above-GT columns per query and shard next to the exact count, K3 walks those lists, and the pairs
    a truncated list cannot settle are recounted exactly from the scores — same integers as the bitmask path."""
    import torch.distributed as dist
    from . import sharded
    dev = _device()
    marks = [("enter", time.perf_counter())]
    dev_marks: List[Tuple[str, Any]] = []                # (phase, CUDA event) on the main stream, only when timings are asked for

    def mark(name: str, stream=None) -> None:
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream if stream is not None else torch.cuda.current_stream())
            dev_marks.append((name, e))
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if world > 1 and group is None:
        group = dist.group.WORLD
    q2d = np.asarray(query_to_doc, dtype=np.int32)
    nq = len(q2d)
    colpali = m_info["type"] == "colpali"
    if colpali:
        qv = queries if isinstance(queries, FlatMultiVector) else FlatMultiVector.from_list(queries)
        dv = docs if isinstance(docs, FlatMultiVector) else FlatMultiVector.from_list(docs)
        n_local = len(dv)
    else:
        n_local = docs.shape[0]
    if world > 1 and shard_sizes is None:
        sizes = torch.zeros((world, 1), dtype=torch.int64, device=dev)
        dist.all_gather(list(sizes.unbind(0)), torch.tensor([n_local], dtype=torch.int64, device=dev), group=group)
        torch.cuda.current_stream().synchronize()                        # (see _host_read)
        shard_sizes = sizes.view(-1).tolist()                            # one device->host read
    if world > 1:
        ends = np.cumsum(np.asarray(shard_sizes, dtype=np.int64))
        bounds = [(int(e - sz), int(e)) for e, sz in zip(ends, shard_sizes)]
        if bounds[rank][1] - bounds[rank][0] != n_local:
            raise ValueError("shard_sizes does not match this rank's documents")
    else:
        bounds = [(0, n_local)]
    g_base, n_docs = bounds[rank][0], bounds[-1][1]
    shard_flags = None
    if colpali and world > 1:
        # the reference pads inside chunks of 10 consecutive documents of the WHOLE list (main.py:465): a shard that
        # does not start on a multiple of 10 needs its neighbours' lengths to know which of its documents get a zero
        # row. One fixed-size all-gather of the (padded) length vectors, one device->host read.
        n_max = max(e - b_ for b_, e in bounds)
        mine = torch.zeros(n_max, dtype=torch.int32)
        mine[:n_local] = torch.from_numpy(np.asarray(dv.lens, dtype=np.int32))
        all_l = torch.empty((world, n_max), dtype=torch.int32, device=dev)
        dist.all_gather(list(all_l.unbind(0)), mine.to(dev), group=group)
        all_h = _host_read(all_l).numpy()
        lens_all = np.concatenate([all_h[r, : e - b_] for r, (b_, e) in enumerate(bounds)])
        shard_flags = _pad_flags(lens_all)[g_base:g_base + n_local]
    # ---- the reference's numpy stream (the draws do not depend on the scores). K6p on a side stream regenerates it
    # while the embeddings cross PCIe and the scoring kernel runs — on EVERY rank (the same 2.5 KB state, broadcast
    # from rank 0), so that each can bootstrap its own block of iterations afterwards.
    r = _rng(rng)
    if list_mode is None:
        list_mode = (not colpali) and nq * n_docs // 8 > BITMASK_LIMIT_BYTES
    if list_mode and colpali:
        raise ValueError("LIST mode applies to the dense path (the MaxSim path materialises its score matrix)")
    split_stats = world > 1 and DEVICE_RNG
    if rank == 0 or split_stats:
        if DEVICE_RNG:
            drng = ops.DeviceRng(r, dev)
            if split_stats:
                dist.broadcast(drng.state, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            side = _aux_stream(dev, "rng")
            side.wait_stream(torch.cuda.current_stream())
            rng_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timings is not None else None
            with torch.cuda.stream(side):
                if rng_ev:
                    rng_ev[0].record(side)
                # main.py:588. With the statistics split over the ranks, each keeps the rows of ITS block of iterations
                # (the stream is replayed in full everywhere — its positions depend on every rejection before them —
                # but segments outside the block skip their write pass)
                b0, b1 = sharded.stats_block(n_iterations, group) if split_stats else (0, n_iterations)
                sample_indices = drng.randint(n_docs, n_iterations * n_docs, window=(b0 * n_docs, (b1 - b0) * n_docs))
                sample_indices = (sample_indices.view(b1 - b0, n_docs) if sample_indices is not None
                                  else torch.empty((0, n_docs), dtype=torch.int32, device=dev))
                ci_idx = drng.randint(n_iterations, len(KS) * N_BOOTSTRAP * n_iterations).view(
                    len(KS), N_BOOTSTRAP, n_iterations)                   # main.py:269, the three CIs' draws in stream order
                rng_done = torch.cuda.Event()
                rng_done.record(side)
                if rng_ev:
                    rng_ev[1].record(side)
    marks.append(("rng_queued", time.perf_counter()))
    mark("start")
    gt_off = torch.arange(nq + 1, dtype=torch.int32, device=dev)
    gt_idx = _np_h2d(q2d, dev, torch.int32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    # ---- score + threshold pass on this rank's shard (async)
    fused_rank = colpali and world == 1 and 4 * nq * n_local > MAXSIM_SCORES_LIMIT
    if fused_rank:
        # one GPU and a score matrix that would not be small: score + rank block by block, no [nq, n_docs] matrix
        q_flat, _ = qv.to_device(dev)
        d_flat, d_off = dv.to_device(dev)
        fl = _pad_flags(dv.lens)
        ev[0].record()
        local = ops.maxsim_rank(q_flat, qv.lens, d_flat, d_off, gt_off, gt_idx,
                                _np_h2d(fl, dev, torch.uint8) if fl.any() else None, g_base=0)
        ev[1].record()
    elif colpali:
        ev[0].record()
        S = _maxsim_streamed(qv, dv, dev, shard_flags, group=group if world > 1 else None)
        ev[1].record()
        local = ops.rank_from_scores(S, gt_off, gt_idx, g_base=g_base, threshold_only=True)
    else:
        Q = queries if queries.is_cuda else sharded.gather_rows(_as_bf16(queries, "queries"), dev, group if world > 1 else None)
        Q, G = _as_bf16(Q, "queries"), _h2d(_as_bf16(docs, "docs"), dev)
        ev[0].record()
        local = ops.dense_rank(Q, G, gt_off, gt_idx, g_base=g_base, threshold_only=True)
        ev[1].record()
    if rank == 0 and not DEVICE_RNG:
        # host mode: numpy draws while the GPU is busy with the kernels queued above
        sample_indices = _np_h2d(r.randint(0, n_docs, size=(n_iterations, n_docs)).astype(np.int32), dev,
                                 torch.int32)                                                   # main.py:588
        ci_idx = _np_h2d(r.randint(0, n_iterations, size=(len(KS), N_BOOTSTRAP, n_iterations)).astype(np.int32), dev,
                         torch.int32)
    marks.append(("score_queued", time.perf_counter()))
    mark("first_pass")
    thr = sharded.merge_thresholds(local.thr, group)
    mark("thresholds_all_reduce")
    if fused_rank:
        pass                                             # thresholds, counts and bit rows are already there
    elif colpali:
        local = ops.rank_from_scores(S, gt_off, gt_idx, g_base=g_base, thr=thr)
    elif list_mode:
        local = ops.dense_rank_lists(Q, G, gt_off, gt_idx, g_base=g_base, thr=thr, cap=list_cap or ops.LIST_CAP)
        recount = ops.dense_unlisted_resolver(Q, G, thr, None, g_base)        # this shard's share of an exact recount
    else:
        local = ops.dense_rank(Q, G, gt_off, gt_idx, g_base=g_base, thr=thr)
    mark("second_pass")
    ranks = sharded.merge_ranks(local, bounds, group)
    mark("merge_all_gather")
    ranks.q_unit = gt_idx
    if keep is not None:
        keep["ranks"] = ranks
    if not split_stats:
        # one GPU and the host-RNG mode keep the statistics on rank 0
        if rank != 0:
            if list_mode:
                sharded.unlisted_round(recount, None, None, group)           # serve rank 0's recount, if there is one
            return None
        if DEVICE_RNG:
            torch.cuda.current_stream().wait_event(rng_done)
        mult = ops.bincount(sample_indices, n_docs)
        if list_mode:
            rounds, k3_stats = [], {}

            def resolve(pairs, mult_rows):
                rounds.append(1)
                return sharded.unlisted_round(recount, pairs, mult_rows, group)
            correct = ops.boot_counts_lists(ranks, mult, resolve, stats=k3_stats)                    # [B, 3] int32
            if not rounds:
                sharded.unlisted_round(recount, None, None, group)           # nothing unsettled: release the other ranks
        else:
            correct = ops.boot_counts(ranks, mult)                                                   # [B, 3] int32
        # query copies drawn in each iteration (integer sum: the same doubles as summing float64 copies)
        denom = ops.query_copies_drawn(mult, gt_idx).to(torch.float64)
        values = 100.0 * correct.to(torch.float64) / denom[:, None]                                  # main.py:358
        means = ops.ci_means_batched(values.t().contiguous(), ci_idx)                                # three CIs, one launch
    else:
        torch.cuda.current_stream().wait_event(rng_done)
        mark("wait_for_numpy_stream")
        for t in (sample_indices, ci_idx):
            t.record_stream(torch.cuda.current_stream())
        k3_stats = {}
        boot_fn = sharded.list_boot_fn(recount, group, k3_stats) if list_mode else None
        values, means = sharded.bootstrap_stats(ranks, sample_indices, ci_idx, gt_idx, n_docs, group, boot_fn=boot_fn,
                                                mark=mark if timings is not None else None, n_iterations=n_iterations)
        if rank != 0:
            return None
    marks.append(("stats_queued", time.perf_counter()))
    torch.cuda.current_stream().synchronize()                            # (see _host_read)
    if DEVICE_RNG:
        drng.commit()                                                    # numpy's generator ends where the reference's would
    values_h = values.cpu().numpy()
    means_h = [m.cpu().numpy() for m in means]
    marks.append(("results_on_host", time.perf_counter()))
    out: Dict[str, Any] = {"Model": m_info.get("name", ""), "n_queries": nq, "n_docs": n_docs}
    alpha = (1 - 0.95) / 2
    for j, k in enumerate(KS):
        v = np.ascontiguousarray(values_h[:, j])
        out[f"R@{k}_mean"] = np.mean(v)
        out[f"R@{k}_lower"] = np.percentile(means_h[j], alpha * 100)
        out[f"R@{k}_upper"] = np.percentile(means_h[j], (1 - alpha) * 100)
        out[f"R@{k}_std"] = np.std(v)
    out["base"] = ranks.recall()
    out["_bootstrap_values"] = values_h
    if list_mode:
        out["_list_mode"] = {"cap": int(ranks.list_i.shape[2]), "shards": int(ranks.list_i.shape[0]),
                             "unresolved_pairs": k3_stats.get("unresolved_pairs", 0)}
    if timings is not None:
        timings["score_ms"] = ev[0].elapsed_time(ev[1])
        # device time between consecutive phase marks on the main stream (this rank), and the numpy-stream replay on its own
        timings["device_ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(dev_marks[:-1], dev_marks[1:])}
        if DEVICE_RNG and (rank == 0 or split_stats) and rng_ev:
            timings["device_ms"]["numpy_stream_replay (side stream)"] = rng_ev[0].elapsed_time(rng_ev[1])
        marks.append(("return", time.perf_counter()))
        timings["wall_ms"] = {k: 1e3 * (t - marks[0][1]) for k, t in marks[1:]}     # host checkpoints since entry
    return out


def run_bootstrap_benchmark(model, processor, m_info: Dict, dataset, n_iterations: int, *, encode_data=None,
                            load_cached_image=None, get_all_captions=None) -> Dict[str, Any]:
    """main.py:478-667. The dataset walk and the encoders stay upstream (out of scope): pass the reference's own
    ``encode_data`` / ``load_cached_image`` / ``get_all_captions`` callables. Everything after the encoders runs on
    the B200 path. Returns the row dict of main.py:643-665 ({} when no valid image, as the reference)."""
    if encode_data is None or load_cached_image is None or get_all_captions is None:
        raise TypeError("run_bootstrap_benchmark needs the upstream encode_data / load_cached_image / "
                        "get_all_captions callables (the encoders are not part of this library)")
    images, t2i_captions, all_captions = [], [], []
    for item in dataset:                                                                   # main.py:508-525
        img = load_cached_image(item)
        if img is None:
            continue
        caps_raw = get_all_captions(item, "sentences")
        if len(caps_raw) < 5:
            continue
        caps = caps_raw[:5]
        images.append(img)
        t2i_captions.append(caps[0])
        all_captions.extend(caps)
    if not images:
        return {}
    t_start = time.time()
    img_embeds, _ = encode_data(model, processor, m_info, images, [])                      # main.py:551-557
    _, t2i_txt_embeds = encode_data(model, processor, m_info, [], t2i_captions)
    _, all_txt_embeds = encode_data(model, processor, m_info, [], all_captions)
    encoding_time = time.time() - t_start
    row = run_bootstrap_from_embeddings(img_embeds, t2i_txt_embeds, all_txt_embeds, m_info, n_iterations,
                                        t2i_captions=t2i_captions, encoding_time=encoding_time, t_start=t_start)
    return {k: v for k, v in row.items() if k in CSV_COLUMNS}


def write_results_csv(rows: List[Dict[str, Any]], path: str) -> None:
    """main.py:834-837: pd.DataFrame(rows).to_csv(path, index=False) with the reference's 42 columns."""
    import pandas as pd
    pd.DataFrame([{k: r[k] for k in CSV_COLUMNS if k in r} for r in rows]).to_csv(path, index=False)

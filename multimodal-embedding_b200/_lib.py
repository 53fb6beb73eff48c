"""ctypes binding of libmmx.so (the C ABI declared in include/mmx.h).

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``) and loaded with
plain ctypes: no torch types cross the boundary, only device pointers, sizes and a CUDA stream handle. There is
no fallback — if the library is missing, or the device is not a B200, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from pathlib import Path
from typing import Optional

_PKG = Path(__file__).resolve().parent
CSRC = _PKG / "csrc"
LIB_PATH = CSRC / "libmmx.so"
SOURCES = ["mmx_host.cu", "maxsim.cu", "dense_rank.cu", "rank_scores.cu", "topk.cu", "stats.cu", "rng.cu", "shuffle.cu",
           "bits.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


class MmxError(RuntimeError):
    """Raised when a libmmx entry point returns non-zero (message from mmx_last_error)."""


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise MmxError("nvcc not found; libmmx.so cannot be built")
    return exe


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into csrc/libmmx.so for sm_100a (cross-compiles without a GPU) and make sure the MT19937
    jump-ahead table of K6p (csrc/mt19937_jump_table.npy, ~15 s of host arithmetic once) sits next to it."""
    from . import mtjump
    mtjump.load_table(build_if_missing=True)
    mtjump.load_table(build_if_missing=True, long=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    deps = srcs + list(CSRC.glob("*.cuh")) + [_PKG.parent / "include" / "mmx.h"]
    if not force and LIB_PATH.exists():
        newest = max(p.stat().st_mtime for p in deps)
        if LIB_PATH.stat().st_mtime >= newest:
            return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, srcs)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
    if res.returncode != 0:
        raise MmxError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


_c = ctypes
_i32p = _c.c_void_p
_SIGNATURES = {
    "mmx_last_error": (_c.c_char_p, []),
    "mmx_abi_version": (_c.c_int, []),
    "mmx_device_check": (_c.c_int, []),
    "mmx_maxsim_workspace_bytes": (_c.c_int, [_c.c_int64, _c.c_int64, _c.c_int64, _c.POINTER(_c.c_size_t)]),
    "mmx_maxsim_scores": (_c.c_int, [
        _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64,      # q_tok, q_off, nq, n_qtok
        _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int64,      # d_tok, d_off, nd, n_dtok
        _c.c_int, _c.c_void_p,                                 # dim, d_zero
        _c.c_void_p, _c.c_int64,                               # scores, ld_scores
        _c.c_void_p, _c.c_size_t, _c.c_void_p,                 # workspace, bytes, stream
    ]),
    "mmx_dense_workspace_bytes": (_c.c_int, [_c.c_int64, _c.c_int64, _c.POINTER(_c.c_size_t)]),
    "mmx_dense_rank": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int,   # Q, nq, G, ng, d
        _c.c_void_p, _c.c_void_p, _c.c_int64,                         # gt_off, gt_idx, g_base
        _c.c_int, _c.c_void_p,                                        # mode, thr
        _c.c_void_p, _c.c_void_p, _c.c_int64,                         # cnt_above, above_bits, words_per_row
        _c.c_void_p, _c.c_int64,                                      # scores, ld_scores
        _c.c_void_p, _c.c_size_t, _c.c_void_p,                        # workspace, bytes, stream
    ]),
    "mmx_rank_from_scores": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_int64,              # scores, nq, ng, ld
        _c.c_void_p, _c.c_void_p, _c.c_int64,                         # gt_off, gt_idx, g_base
        _c.c_int, _c.c_void_p,                                        # mode, thr
        _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_void_p,            # cnt_above, above_bits, words_per_row, stream
    ]),
    "mmx_topk_from_scores": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int64, _c.c_int,    # scores, nq, ng, ld, g_base, k
        _c.c_void_p, _c.c_void_p, _c.c_void_p,                                    # top_s, top_i, stream
    ]),
    "mmx_dense_topk_workspace_bytes": (_c.c_int, [_c.c_int64, _c.c_int64, _c.POINTER(_c.c_size_t)]),
    "mmx_dense_topk": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int64, _c.c_int,   # Q, nq, G, ng, d, g_base, k
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_void_p,                    # top_s, top_i, ws, bytes, stream
    ]),
    "mmx_dense_rank_topk": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int,   # Q, nq, G, ng, d
        _c.c_void_p, _c.c_void_p, _c.c_int64,                         # gt_off, gt_idx, g_base
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64,            # thr, cnt_above, above_bits, words_per_row
        _c.c_int, _c.c_void_p, _c.c_void_p,                           # k, top_s, top_i
        _c.c_void_p, _c.c_size_t, _c.c_void_p,                        # workspace, bytes, stream
    ]),
    "mmx_topk_select": (_c.c_int, [
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int,    # cand_s, cand_i, cand_n, nq, cap, k
        _c.c_void_p, _c.c_void_p, _c.c_void_p,                                    # top_s, top_i, stream
    ]),
    "mmx_merge_topk": (_c.c_int, [
        _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_int,                 # gathered_s, gathered_i, n_shards, nq, k
        _c.c_void_p, _c.c_void_p, _c.c_void_p,                                    # top_s, top_i, stream
    ]),
    "mmx_bincount": (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,  # idx,B,n_draws,n_units,scratch,mult,stream
    ]),
    "mmx_above_lists": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int64, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p,  # bits,nq,wpr,col2unit,L,list_t,stream
    ]),
    "mmx_boot_counts": (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_int64,  # list_t, L, cnt_above, bits, wpr
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64,  # col2unit, q_unit, mult, B, n_units
        _c.c_int64, _c.c_void_p, _c.c_void_p, _c.c_void_p,            # nq, correct, q_hits1, stream
    ]),
    "mmx_copies_drawn": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p]),
    "mmx_dense_rank_lists": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int64,   # Q, nq, G, ng, d, g_base
        _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p,  # thr, cnt_above, list_i, list_n, cap, stream
    ]),
    "mmx_boot_counts_lists_workspace_bytes": (_c.c_int, [_c.c_int, _c.c_int64, _c.POINTER(_c.c_size_t)]),
    "mmx_boot_counts_lists": (_c.c_int, [
        _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int64, _c.c_int,     # list_i, list_n, n_shards, nq, cap
        _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p,              # col2unit, L, cnt_above, q_unit
        _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p, _c.c_void_p,  # mult, B, n_units, correct, q_hits1
        _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p,  # unres, unres_n, unres_cap, workspace, bytes, stream
    ]),
    "mmx_ci_means": (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p,   # data, n, idx, R, means, stream
    ]),
    "mmx_ci_means_batched": (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,   # data, n, idx, R, batch, means, stream
    ]),
    "mmx_mt19937_randint": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_int64, _c.c_void_p, _c.c_void_p]),
    "mmx_mt19937_par_scratch_bytes": (_c.c_int, [_c.c_int, _c.POINTER(_c.c_size_t)]),
    "mmx_mt19937_par_set_bit_positions": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int]),
    "mmx_mt19937_par_set_slack": (_c.c_int, [_c.c_double, _c.c_double]),
    "mmx_mt19937_randint_par": (_c.c_int, [
        _c.c_void_p, _c.c_uint32, _c.c_int64, _c.c_void_p,            # state, high, count, out
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p,   # polys, n_poly, seg_blocks, scratch, bytes, stream
    ]),
    "mmx_mt19937_randint_par_window": (_c.c_int, [
        _c.c_void_p, _c.c_uint32, _c.c_int64, _c.c_void_p, _c.c_int64, _c.c_int64,   # state, high, count, out, win_begin, win_count
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p,   # polys, n_poly, seg_blocks, scratch, bytes, stream
    ]),
    "mmx_mt19937_raw_par": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_void_p,                                     # state, count, out
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_size_t, _c.c_void_p,   # polys, n_poly, seg_blocks, scratch, bytes, stream
    ]),
    "mmx_perm_shuffle_workspace_bytes": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int64, _c.POINTER(_c.c_size_t)]),
    "mmx_perm_shuffle": (_c.c_int, [
        _c.c_void_p, _c.c_int64, _c.c_int, _c.c_int, _c.c_void_p, _c.c_void_p,    # words, n_words, n, P, perm, info
        _c.c_void_p, _c.c_size_t, _c.c_void_p,                                    # workspace, bytes, stream
    ]),
    "mmx_place_bit_columns": (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int64, _c.c_int64, _c.c_void_p,               # gathered, n_shards, nq, wmax, shard_begin_h
        _c.c_void_p, _c.c_int64, _c.c_void_p, _c.c_void_p,                        # above_bits, words_per_row, cnt_above, stream
    ]),
    "mmx_perm_stats": (_c.c_int, [
        _c.c_void_p, _c.c_int, _c.c_int, _c.c_void_p, _c.c_int, _c.c_void_p, _c.c_void_p,  # combined,n1,n2,perm,P,diffs,stream
    ]),
}

_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Load libmmx.so (once). Raises MmxError when it has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise MmxError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libmmx has no CPU or PyTorch fallback)")
        h = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            if hasattr(h, name):
                fn = getattr(h, name)
                fn.restype = res
                fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MmxError(lib().mmx_last_error().decode("utf-8", "replace") or f"libmmx error {rc}")


def declared_symbols() -> list[str]:
    """Names declared in include/mmx.h (used by the CPU test that checks the library exports all of them)."""
    import re
    text = (_PKG.parent / "include" / "mmx.h").read_text()
    return sorted(set(re.findall(r"\b(mmx_[a-z0-9_]+)\s*\(", text)))

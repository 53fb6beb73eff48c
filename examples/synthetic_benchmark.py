#!/usr/bin/env python
"""The reference's whole post-encoder flow on synthetic embeddings, start to finish on the B200 path:

    per model:  run_bootstrap_benchmark (main.py:478-667)  ->  row of the 42-column CSV (main.py:834-837)
    afterwards: pairwise permutation tests on the per-iteration arrays + the analysis report
                (analyze_statistical_results.py, with real p-values instead of the CI-overlap approximation)

    python examples/synthetic_benchmark.py --images 5000 --iterations 1000 --out /tmp/mmx_demo

The "encoders" are synthetic: unit-norm image embeddings, captions = noisy copies; ColPali-style models are
lists of [L_i, 128] multi-vectors. Everything after them is the library: no score matrix for the dense models, device
RNG replaying numpy's stream, bit-identical statistics."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mmx  # noqa: E402


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def dense_model(n, dim, sigma, seed):
    g = torch.Generator().manual_seed(seed)
    img = unit(torch.randn(n, dim, generator=g))
    all_txt = unit(img.repeat_interleave(5, 0) + sigma * unit(torch.randn(5 * n, dim, generator=g)))
    return img.bfloat16(), all_txt[0::5].contiguous().bfloat16(), all_txt.bfloat16()


def colpali_model(n, patches, sigma, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(12, 30, (5 * n,), generator=g)
    P = unit(torch.randn(n * patches, 128, generator=g))
    owner = torch.repeat_interleave(torch.arange(5 * n), lens) // 5
    rows = torch.randint(0, patches, (int(lens.sum()),), generator=g)
    T = unit(P[owner * patches + rows] + sigma * unit(torch.randn(int(lens.sum()), 128, generator=g))).bfloat16()
    off = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)])
    t2i_rows = torch.cat([torch.arange(int(off[5 * i]), int(off[5 * i + 1])) for i in range(n)])
    return (mmx.FlatMultiVector(P.bfloat16(), np.full(n, patches)), mmx.FlatMultiVector(T[t2i_rows], lens.numpy()[0::5]),
            mmx.FlatMultiVector(T, lens.numpy()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=5000)
    ap.add_argument("--iterations", type=int, default=1000)
    ap.add_argument("--patches", type=int, default=1030)
    ap.add_argument("--permutations", type=int, default=10000)
    ap.add_argument("--out", default="mmx_demo")
    ap.add_argument("--seed", type=int, default=42)
    a = ap.parse_args()
    out = Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    n = a.images
    captions = [("a red bus next to two people %d" if i % 3 == 0 else "a cat on a couch %d") % i for i in range(n)]
    models = [("Synth-CLIP-L", "dense", lambda: dense_model(n, 768, 8.0, 1)),
              ("Synth-SigLIP-400M", "siglip", lambda: dense_model(n, 1152, 10.0, 2)),
              ("Synth-CLIP-H", "dense", lambda: dense_model(n, 1024, 9.0, 3)),
              ("Synth-ColPali", "colpali", lambda: colpali_model(n, a.patches, 3.6, 4))]
    np.random.seed(a.seed)                                    # the reference seeds the global stream (main.py:99)
    rows = []
    for name, kind, make in models:
        img, t2i, all_txt = make()
        t0 = time.time()
        row = mmx.run_bootstrap_from_embeddings(img, t2i, all_txt, {"type": kind, "name": name}, a.iterations,
                                                t2i_captions=captions, encoding_time=1.0, tie_report=True)
        torch.cuda.synchronize()
        print(f"{name:20s} {time.time() - t0:7.3f} s   T2I R@1 {row['T2I_R@1_mean']:.2f} "
              f"[{row['T2I_R@1_lower']:.2f}, {row['T2I_R@1_upper']:.2f}]   I2T R@1 {row['I2T_R@1_mean']:.2f}   "
              f"ties {row['_tie_report']['T2I']['queries_with_exact_ties']}", flush=True)
        rows.append(row)
    csv = out / "benchmark_v29_statistical_results.csv"
    mmx.write_results_csv(rows, str(csv))
    import pandas as pd
    data = mmx.collect_bootstrap_data(rows)
    t0 = time.time()
    df = mmx.add_permutation_columns(pd.read_csv(csv), data, n_permutations=a.permutations)
    df.to_csv(out / "benchmark_v29_with_permutation_tests.csv", index=False)
    mmx.generate_report(df, str(out / "analysis_report.md"), bootstrap_data=data, n_permutations=a.permutations)
    print(f"permutation tests + report: {time.time() - t0:.3f} s -> {out}")


if __name__ == "__main__":
    main()

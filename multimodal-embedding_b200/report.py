"""Report generator with REAL pairwise permutation tests.

The reference's analysis script approximates significance from confidence-interval overlap because it only sees the
CSV (analyze_statistical_results.py:47-102; its ``np.erf`` and ``json.loads(str(dict))`` calls fail at run time,
SURVEY 5.2 D5). The hot path keeps every model's per-iteration metric arrays (``row["_bootstrap_values"]`` — what
main.py:790 ``all_model_bootstrap_data`` was meant to hold), so the matrix here comes from
``permutation_test`` (main.py:281-302, K5 on the device) on those arrays. Outputs: the p-value matrix, additive
``perm_p_<metric>`` CSV columns (each model against the best model of that metric) and a markdown report with the
sections of the reference's report.
"""
from __future__ import annotations

import ast
import math
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

from .api import permutation_test

BootstrapData = Dict[str, Dict[str, np.ndarray]]       # model -> metric key -> per-iteration values


def collect_bootstrap_data(rows: Iterable[Dict[str, Any]]) -> BootstrapData:
    """model name -> ``_bootstrap_values`` of the rows returned by run_bootstrap_from_embeddings."""
    return {r["Model"]: r["_bootstrap_values"] for r in rows if "_bootstrap_values" in r}


def pairwise_permutation_tests(bootstrap_data: BootstrapData, metric: str = "T2I_R@1", n_permutations: int = 10000,
                               rng=None, models: Optional[Sequence[str]] = None):
    """Symmetric p-value matrix (pandas DataFrame, diagonal 1.0) of ``permutation_test`` on the per-iteration arrays
    of `metric`. numpy stream: pairs (i, j), i < j, in `models` order, n_permutations shuffles each."""
    import pandas as pd
    models = list(bootstrap_data) if models is None else list(models)
    m = len(models)
    p = np.ones((m, m), dtype=np.float64)
    for i in range(m):
        for j in range(i + 1, m):
            p[i, j] = p[j, i] = permutation_test(bootstrap_data[models[i]][metric], bootstrap_data[models[j]][metric],
                                                 n_permutations, rng=rng)
    return pd.DataFrame(p, index=models, columns=models)


def ci_overlap_significance(df, metric: str, alpha: float = 0.05):
    """The reference's CSV-only approximation (analyze_statistical_results.py:47-102) with its erf call repaired:
    disjoint CIs -> 0.01, otherwise 2 * (1 - Phi(|mean_i - mean_j| / pooled_std))."""
    import pandas as pd
    models = df["Model"].tolist()
    col = {s: df.set_index("Model")[f"{metric}_{s}"] for s in ("mean", "lower", "upper", "std")}
    out = np.ones((len(models), len(models)))
    for i, a in enumerate(models):
        for j, b in enumerate(models):
            if i == j:
                continue
            if col["upper"][a] < col["lower"][b] or col["upper"][b] < col["lower"][a]:
                out[i, j] = 0.01
                continue
            pooled = math.sqrt((col["std"][a] ** 2 + col["std"][b] ** 2) / 2)
            z = abs(col["mean"][a] - col["mean"][b]) / pooled if pooled > 0 else 0.0
            out[i, j] = min(2 * (1 - 0.5 * (1 + math.erf(z / math.sqrt(2)))), 1.0)
    return pd.DataFrame(out, index=models, columns=models)


def compute_pairwise_significance(df, metric: str, alpha: float = 0.05, bootstrap_data: Optional[BootstrapData] = None,
                                  n_permutations: int = 10000, rng=None):
    """Same name and leading arguments as analyze_statistical_results.py:47. With ``bootstrap_data`` the matrix holds
    permutation-test p-values; without it, the reference's CI-overlap approximation."""
    if bootstrap_data is None:
        return ci_overlap_significance(df, metric, alpha)
    return pairwise_permutation_tests(bootstrap_data, metric, n_permutations, rng, models=df["Model"].tolist())


def add_permutation_columns(df, bootstrap_data: BootstrapData, metrics: Sequence[str] = ("T2I_R@1", "I2T_R@1", "I2T_Sym_R@1"),
                            n_permutations: int = 10000, rng=None):
    """Additive CSV columns ``perm_p_<metric>``: p-value of each model against the best model (highest mean) of that
    metric, 1.0 for the best model itself. The reference's 42 columns are left untouched."""
    df = df.copy()
    for metric in metrics:
        best = df.loc[df[f"{metric}_mean"].idxmax(), "Model"]
        vals = []
        for model in df["Model"]:
            vals.append(1.0 if model == best else float(permutation_test(
                bootstrap_data[model][metric], bootstrap_data[best][metric], n_permutations, rng=rng)))
        df[f"perm_p_{metric}"] = vals
    return df


def stars(p: float) -> str:
    return "***" if p < 0.001 else "**" if p < 0.01 else "*" if p < 0.05 else "ns"


def _table(header: Sequence[str], rows: Iterable[Sequence[str]], align: Optional[Sequence[str]] = None) -> List[str]:
    align = align or [":--"] * len(header)
    out = ["| " + " | ".join(header) + " |", "|" + "|".join(align) + "|"]
    out += ["| " + " | ".join(map(str, r)) + " |" for r in rows]
    return out + [""]


def _ci(row, key: str) -> str:
    return f"{row[key + '_mean']:.1f}% [{row[key + '_lower']:.1f}, {row[key + '_upper']:.1f}]"


def _failure(row) -> Dict[str, Any]:
    fa = row.get("_failure_analysis", {})
    if isinstance(fa, str):                       # the CSV holds str(dict) (main.py:664), which json cannot parse
        try:
            fa = ast.literal_eval(fa)
        except (ValueError, SyntaxError):
            fa = {}
    return fa if isinstance(fa, dict) else {}


def generate_report(df, output_path: str, alpha: float = 0.05, bootstrap_data: Optional[BootstrapData] = None,
                    n_permutations: int = 10000, rng=None) -> str:
    """Markdown report with the reference's sections (ranked CI tables for the three protocols, T2I R@1 significance
    matrix, failure analysis, per-category accuracy, speed). Returns the text and writes it to `output_path`."""
    by_t2i = df.sort_values("T2I_R@1_mean", ascending=False)
    models = by_t2i["Model"].tolist()
    lines = ["# Statistical Analysis Report", "",
             "Bootstrap confidence intervals, pairwise significance, failure analysis and per-category accuracy.", ""]
    lines += ["## Text-to-Image Retrieval (95% CI)", ""]
    lines += _table(["Rank", "Model", "R@1", "R@5", "R@10"],
                    [(i, f"**{r['Model']}**", _ci(r, "T2I_R@1"), _ci(r, "T2I_R@5"), _ci(r, "T2I_R@10"))
                     for i, (_, r) in enumerate(by_t2i.iterrows(), 1)])
    sig = compute_pairwise_significance(df, "T2I_R@1", alpha, bootstrap_data, n_permutations, rng)
    how = "permutation test on the per-iteration bootstrap values" if bootstrap_data is not None else \
        "CI-overlap approximation (no bootstrap samples supplied)"
    lines += [f"### Pairwise significance, T2I R@1 ({how})", ""]
    lines += _table(["Model"] + models,
                    [[a] + ["-" if a == b else stars(sig.loc[a, b]) for b in models] for a in models],
                    [":--"] + ["--:"] * len(models))
    lines += ["`***` p < 0.001, `**` p < 0.01, `*` p < 0.05, `ns` not significant", ""]
    for title, key in (("Image-to-Text, standard protocol (5 captions per image)", "I2T"),
                       ("Image-to-Text, symmetric protocol (1 caption per image)", "I2T_Sym")):
        ranked = df.sort_values(f"{key}_R@1_mean", ascending=False)
        lines += [f"## {title}", ""]
        lines += _table(["Rank", "Model", "R@1", "R@5", "R@10", "Gap vs T2I"],
                        [(i, r["Model"], _ci(r, f"{key}_R@1"), _ci(r, f"{key}_R@5"), _ci(r, f"{key}_R@10"),
                          f"{r[key + '_R@1_mean'] - r['T2I_R@1_mean']:+.1f}pp")
                         for i, (_, r) in enumerate(ranked.iterrows(), 1)])
    if "_failure_analysis" in df.columns:
        fas = {r["Model"]: _failure(r) for _, r in by_t2i.iterrows()}
        keys = ["overall_accuracy", "accuracy_has_spatial", "accuracy_not_has_spatial", "accuracy_has_color",
                "accuracy_not_has_color", "accuracy_has_counting", "accuracy_not_has_counting"]
        lines += ["## Failure analysis (T2I R@1 by query feature)", ""]
        lines += _table(["Model", "Overall", "Spatial", "No spatial", "Color", "No color", "Counting", "No counting"],
                        [[m] + [f"{fas[m].get(k, 0):.1f}%" for k in keys] for m in models])
        cats = sorted({c for fa in fas.values() for c in fa.get("accuracy_by_category", {})})
        if cats:
            lines += ["## Per-category accuracy (COCO supercategories)", ""]
            lines += _table(["Model"] + [c.title() for c in cats],
                            [[m] + [f"{fas[m].get('accuracy_by_category', {}).get(c, 0):.1f}%" for c in cats]
                             for m in models])
    lines += ["## Speed", ""]
    lines += _table(["Model", "QPS", "Encoding time", "T2I R@1", "Efficiency (R@1/100 x QPS)"],
                    [(r["Model"], f"{r.get('QPS', 0):.1f}", f"{r.get('Encoding_Time', 0):.1f}s",
                      f"{r['T2I_R@1_mean']:.1f}%", f"{r['T2I_R@1_mean'] / 100 * r.get('QPS', 0):.1f}")
                     for _, r in by_t2i.iterrows()])
    if len(models) > 1:
        p = float(sig.loc[models[0], models[1]])
        verdict = "statistically significant" if p < alpha else "not statistically significant"
        lines += ["## Conclusion", "", f"**{models[0]}** has the highest T2I R@1 ({_ci(by_t2i.iloc[0], 'T2I_R@1')}); "
                  f"the difference to {models[1]} is {verdict} (p = {p:.4f}, alpha = {alpha}).", ""]
    text = "\n".join(lines)
    with open(output_path, "w") as f:
        f.write(text)
    return text

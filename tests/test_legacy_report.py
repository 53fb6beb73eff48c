"""legacy metric variants (v19 MRR, Winoground) and the report generator.

CPU part: the oracle against golden strings produced by the reference's own v19 function, and the report's host logic.
GPU part (-m gpu): the CUDA-backed functions against the same golden strings and the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from helpers import synth_dense, unit


@pytest.fixture(scope="module")
def legacy_golden(golden_dir):
    z = np.load(golden_dir / "dense.npz")
    g = json.loads((golden_dir / "legacy_v19.json").read_text())
    return torch.from_numpy(z["scores_t2i"]), torch.from_numpy(z["scores_i2t"]), g


def _cases(s_t2i, s_i2t, g):
    n = s_t2i.shape[0]
    yield "blocks5", (s_t2i, s_i2t, list(range(n)), {i: list(range(5 * i, 5 * i + 5)) for i in range(n)})
    yield "defaults", (s_t2i, s_t2i.t().contiguous(), None, None)
    yield "ragged", (s_t2i, s_i2t, g["ragged_q2i"], {int(k): v for k, v in g["ragged_i2c"].items()})
    yield "narrow", (s_t2i[:, :7].contiguous(), s_i2t[:, :4].contiguous(), [i % 7 for i in range(n)],
                     {i: [i % 4] for i in range(n)})


def test_oracle_v19_metrics_match_reference_strings(legacy_golden):
    s_t2i, s_i2t, g = legacy_golden
    for name, (a, b, q2i, i2c) in _cases(s_t2i, s_i2t, g):
        assert oracle.compute_bidirectional_metrics(a, b, q2i, i2c) == g[name], name


def test_winoground_host_counters_match_oracle():
    import mmx
    rs = np.random.RandomState(0)
    blocks = torch.from_numpy(rs.normal(size=(400, 2, 2)).astype(np.float32))
    blocks[3] = 0.5                                  # exact ties count as failures (strict >)
    assert mmx.winoground_scores(blocks) == oracle.winoground_counts(blocks)
    with pytest.raises(ValueError):
        mmx.winoground_scores(torch.zeros(4, 2, 3))


def _fake_rows(n_models=4, seed=0):
    import mmx
    rs = np.random.RandomState(seed)
    rows = []
    for m in range(n_models):
        row = {"Model": f"model-{m}"}
        vals = {}
        for k in mmx.METRIC_KEYS:
            v = np.round(rs.normal(40 + 3 * m, 0.7, size=200) * 50) / 50
            vals[k] = v
            row[f"{k}_mean"], row[f"{k}_std"] = float(np.mean(v)), float(np.std(v))
            row[f"{k}_lower"], row[f"{k}_upper"] = row[f"{k}_mean"] - 0.1, row[f"{k}_mean"] + 0.1
        row.update({"Time": 1.0, "QPS": 100.0 + m, "Encoding_Time": 50.0, "Img_per_sec": 100.0,
                    "_failure_analysis": str({"overall_accuracy": 40.0 + m, "accuracy_has_color": 30.0,
                                              "accuracy_by_category": {"animal": 50.0, "food": 45.5}})})
        row["_bootstrap_values"] = vals
        rows.append(row)
    return rows


def test_report_from_csv_only(tmp_path):
    """Without bootstrap samples the report falls back to the reference's CI-overlap approximation (with its erf call
    repaired) and parses the str(dict) failure column the reference's json.loads cannot read."""
    import pandas as pd
    import mmx
    rows = _fake_rows()
    csv = tmp_path / "res.csv"
    mmx.write_results_csv(rows, str(csv))
    df = pd.read_csv(csv)
    assert list(df.columns) == mmx.CSV_COLUMNS
    sig = mmx.compute_pairwise_significance(df, "T2I_R@1")
    assert sig.shape == (4, 4) and np.allclose(np.diag(sig.values), 1.0) and np.allclose(sig.values, sig.values.T)
    assert sig.loc["model-0", "model-3"] == 0.01                      # disjoint intervals
    text = mmx.generate_report(df, str(tmp_path / "report.md"))
    assert "| 1 | **model-3** |" in text and "CI-overlap approximation" in text
    assert "| model-3 | 43.0% |" in text and "Animal" in text and "50.0%" in text
    assert (tmp_path / "report.md").read_text() == text


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_v19_metrics_on_device_match_reference_strings(legacy_golden):
    import mmx
    s_t2i, s_i2t, g = legacy_golden
    for name, (a, b, q2i, i2c) in _cases(s_t2i, s_i2t, g):
        assert mmx.compute_bidirectional_metrics(None, None, a, b, q2i, None, i2c) == g[name], name


@pytest.mark.gpu
def test_winoground_from_embeddings_matches_per_example_reference_loop():
    import mmx
    g = torch.Generator().manual_seed(4)
    n, d = 400, 256
    base = unit(torch.randn(n, d, generator=g)).repeat_interleave(2, 0)      # the two images of an example are close
    img = unit(base + 0.2 * unit(torch.randn(2 * n, d, generator=g)))
    txt = unit(img + 3.0 * unit(torch.randn(2 * n, d, generator=g)))
    img, txt = img.bfloat16(), txt.bfloat16()
    blocks = torch.stack([oracle.dense_scores(img[2 * e:2 * e + 2], txt[2 * e:2 * e + 2]) for e in range(n)])
    want = oracle.winoground_counts(blocks)                                         # v28 :440-452 per example
    got = mmx.winoground_from_embeddings(img, txt, {"type": "dense", "name": "w"})
    assert got == want and 0 < want["Wino Group"] < 100
    # ColPali flavour: ragged captions, equal-length images inside an example
    from helpers import synth_colpali
    docs, queries = synth_colpali(40, 40, doc_len=90, q_len=(5, 14), sigma_q=2.0, seed=9)
    bl = torch.stack([oracle.maxsim_scores(queries[2 * e:2 * e + 2], docs[2 * e:2 * e + 2]) for e in range(20)])
    got = mmx.winoground_from_embeddings(docs, queries, {"type": "colpali", "name": "w"})
    assert got == oracle.winoground_counts(bl)


@pytest.mark.gpu
def test_pairwise_permutation_matrix_matches_reference_stream(tmp_path):
    """The significance matrix consumes numpy's stream pair by pair exactly like calling the reference's
    permutation_test in the same order; the perm_p_* columns are additive."""
    import pandas as pd
    import mmx
    rows = _fake_rows(3, seed=2)
    data = mmx.collect_bootstrap_data(rows)
    got = mmx.pairwise_permutation_tests(data, "T2I_R@1", 500, rng=np.random.RandomState(3))
    rs = np.random.RandomState(3)
    names = list(data)
    for i in range(3):
        for j in range(i + 1, 3):
            p = oracle.permutation_test(data[names[i]]["T2I_R@1"], data[names[j]]["T2I_R@1"], 500, rng=rs)
            assert got.loc[names[i], names[j]] == p == got.loc[names[j], names[i]]
    df = pd.DataFrame([{k: r[k] for k in mmx.CSV_COLUMNS} for r in rows])
    df2 = mmx.add_permutation_columns(df, data, ("T2I_R@1",), 300, rng=np.random.RandomState(5))
    assert list(df2.columns)[:42] == mmx.CSV_COLUMNS and df2["perm_p_T2I_R@1"].iloc[2] == 1.0
    text = mmx.generate_report(df, str(tmp_path / "r.md"), bootstrap_data=data, n_permutations=300,
                               rng=np.random.RandomState(6))
    assert "permutation test on the per-iteration bootstrap values" in text


@pytest.mark.gpu
def test_example_flow_end_to_end(tmp_path):
    """examples/synthetic_benchmark.py at a small size: four models -> CSV with the reference's 42 columns -> additive
    perm_p_* columns -> report with a permutation-test significance matrix."""
    import subprocess, sys
    import pandas as pd
    import mmx
    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, str(root / "examples" / "synthetic_benchmark.py"), "--images", "300", "--iterations",
                          "60", "--patches", "64", "--permutations", "400", "--out", str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    df = pd.read_csv(tmp_path / "benchmark_v29_statistical_results.csv")
    assert list(df.columns) == mmx.CSV_COLUMNS and len(df) == 4
    df2 = pd.read_csv(tmp_path / "benchmark_v29_with_permutation_tests.csv")
    assert list(df2.columns)[:42] == mmx.CSV_COLUMNS and "perm_p_T2I_R@1" in df2.columns
    assert "permutation test on the per-iteration bootstrap values" in (tmp_path / "analysis_report.md").read_text()

"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol (no compute calls),
the failure-analysis mirror matches the reference's golden output, schema constants, padding flags, loud failure
without a GPU."""
import json

import numpy as np
import pytest
import torch

import mmx
import oracle
from helpers import from_bits


def test_library_builds_and_exports_every_declared_symbol():
    path = mmx.build()
    assert path.exists()
    h = mmx._lib.lib()
    declared = mmx._lib.declared_symbols()
    assert len(declared) >= 13
    missing = [s for s in declared if not hasattr(h, s)]
    assert missing == []
    assert h.mmx_abi_version() == 1
    assert set(declared) <= set(mmx._lib._SIGNATURES) | {"mmx_last_error"}


def test_csv_schema_matches_reference_order():
    assert mmx.CSV_COLUMNS == oracle.CSV_COLUMNS
    assert len(mmx.CSV_COLUMNS) == 42
    assert mmx.METRIC_KEYS == oracle.METRIC_KEYS


def test_failure_analysis_matches_reference_golden(golden_dir):
    z = np.load(golden_dir / "dense.npz")
    caps = [str(c) for c in z["captions"]]
    hits = z["correct_r1"].astype(np.int64)
    got = mmx.failure.aggregate_failure_from_hits(caps, hits, np.ones(len(caps), dtype=np.int64))
    want = json.loads((golden_dir / "failure_analysis.json").read_text())
    assert json.loads(json.dumps(got, sort_keys=True)) == want
    recs = [{"caption": c, "correct_r1": bool(h)} for c, h in zip(caps, hits)]
    got2 = mmx.aggregate_failure_analysis([recs[:20], recs[20:]])
    assert json.loads(json.dumps(got2, sort_keys=True)) == want
    for c in caps[:5]:
        assert mmx.analyze_query_complexity(c) == oracle.metrics.analyze_query_complexity(c)
        assert mmx.get_category_from_caption(c) == oracle.metrics.get_category_from_caption(c)
    # the caption-only half precomputed (what the row builder does while the GPU is busy) and its small cache
    table = mmx.failure.caption_feature_table(caps)
    assert mmx.failure.caption_feature_table(list(caps)) is table                       # same captions: cached
    got3 = mmx.failure.aggregate_failure_from_hits(caps, hits, np.ones(len(caps), dtype=np.int64), table=table)
    assert json.loads(json.dumps(got3, sort_keys=True)) == want
    for k in range(12):                                                                 # the cache keeps the last few
        mmx.failure.caption_feature_table([f"a dog {k}"])
    assert len(mmx.failure._TABLE_CACHE) <= 8
    assert mmx.failure.caption_feature_table([]) [2].shape == (0,)


def test_pad_flags_and_flatten():
    lens = np.array([3, 3, 5, 1, 1, 1, 1, 1, 1, 1, 7, 7])
    assert list(mmx.api._pad_flags(lens)) == list(oracle.maxsim_pad_flags(lens))
    assert list(mmx.api._pad_flags(lens)) == [1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 0]


def test_csr_from_sets():
    off, idx = mmx.ops.csr_from_sets([[3], [], [1, 2]])
    assert off.tolist() == [0, 1, 1, 3] and idx.tolist() == [3, 1, 2]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the loud failure on a GPU-less host")
def test_no_cpu_fallback():
    img = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(mmx.MmxError):
        mmx.compute_similarity_scores(img, img, {"type": "dense"}, None)
    with pytest.raises(mmx.MmxError):
        mmx.bootstrap_confidence_interval(np.arange(10.0))
    with pytest.raises(mmx.MmxError):
        mmx.ops.dense_scores(img, img)


def test_flat_multivector_disk_cache_roundtrip(tmp_path):
    g = torch.Generator().manual_seed(0)
    items = [torch.randn(int(l), 128, generator=g).bfloat16() for l in (3, 1, 17, 1030, 0, 5)]
    fv = mmx.FlatMultiVector.from_list(items)
    fv.save(str(tmp_path / "emb"))
    back = mmx.FlatMultiVector.load(str(tmp_path / "emb"), pin=False)
    assert list(back.lens) == [3, 1, 17, 1030, 0, 5] and back.flat.dtype == torch.bfloat16
    assert torch.equal(back.flat.view(torch.int16), fv.flat.view(torch.int16))
    assert all(torch.equal(a.view(torch.int16), b.view(torch.int16)) for a, b in zip(back.to_list(), items))
    part = back.slice(2, 4)
    assert len(part) == 2 and part.flat.shape[0] == 17 + 1030


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (or the examples) may import or execute it."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    files = list((root / "multimodal-embedding_b200").glob("*.py")) + list((root / "examples").glob("*.py")) + [root / "mmx.py"]
    bad = [f.name for f in files if re.search(r"^\s*(import|from)\s+oracle\b", f.read_text(), flags=re.M)]
    assert bad == []
    csrc = (root / "multimodal-embedding_b200" / "csrc")
    assert not any("oracle" in p.read_text() for p in list(csrc.glob("*.cu")) + list(csrc.glob("*.cuh")))

"""mmx — B200-native retrieval scoring for the multimodal-embedding benchmark.

Drop-in for everything sariekr/multimodal-embedding's ``main.py`` does after the encoders return embeddings:
the reference's own function names and signatures (see :mod:`api`), executed by hand-written sm_100a CUDA in
``csrc/libmmx.so`` behind the C ABI of ``include/mmx.h``. Import as ``import mmx`` (root shim) or
``importlib.import_module("multimodal-embedding_b200")``. No CPU fallback: without the built library and a B200 the
compute functions raise.
"""
from . import _lib, ops, sharded, failure, legacy, report                   # noqa: F401
from ._lib import MmxError, build                                           # noqa: F401
from .api import (                                                          # noqa: F401
    CSV_COLUMNS, METRIC_KEYS, FlatMultiVector, base_metrics, run_retrieval_bootstrap, bootstrap_confidence_interval, bootstrap_from_ranks, compute_metrics,
    compute_metrics_detailed, compute_similarity_scores, permutation_test, run_bootstrap_benchmark,
    run_bootstrap_from_embeddings, score_and_rank, write_results_csv,
)
from .legacy import compute_bidirectional_metrics, winoground_from_embeddings, winoground_scores  # noqa: F401
from .report import (add_permutation_columns, collect_bootstrap_data, compute_pairwise_significance,  # noqa: F401
                     generate_report, pairwise_permutation_tests)
from .failure import aggregate_failure_analysis, analyze_query_complexity, get_category_from_caption  # noqa: F401

__version__ = "0.1.0"

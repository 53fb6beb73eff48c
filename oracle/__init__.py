"""CPU oracle for the retrieval-scoring path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or the CPU arm.
The product path (``multimodal-embedding_b200``) never imports this package and has no CPU fallback.

Every function restates the algorithm of ``/root/reference`` (sariekr/multimodal-embedding) and cites the
file:line it follows. Pinning: ``tests/golden/make_golden.py`` imports the reference's own ``main.py`` in the build
container, runs its functions on seeded synthetic inputs and commits inputs + outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this oracle against those vectors. The MaxSim arithmetic lives in the
third-party ``colpali_engine`` (requirements.txt:9, unpinned; paper.md:146 names 0.3.1) which is not vendored; its
expression is anchored on the identical in-container HF port
``transformers.models.colpali.processing_colpali.ColPaliProcessor.score_retrieval`` driven through the
reference's own chunk loop (main.py:451-470).
"""
from .scoring import (  # noqa: F401
    dense_scores, maxsim_scores, maxsim_pad_flags, flatten_multivector,
)
from .metrics import (  # noqa: F401
    compute_metrics_detailed, compute_metrics, above_gt_counts, recall_from_counts, topk_lists, METRIC_KEYS,
)
from .stats import (  # noqa: F401
    pairwise_sum, np_mean, np_std, percentile_linear, bootstrap_confidence_interval, permutation_test,
    ci_resample_indices, permutation_indices,
)
from .bootstrap import (  # noqa: F401
    bootstrap_rescore_reference, bootstrap_identity, aggregate_row, CSV_COLUMNS,
)
from .legacy import compute_bidirectional_metrics, winoground_counts  # noqa: F401

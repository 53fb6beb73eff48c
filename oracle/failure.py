"""Oracle: failure analysis aggregation (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows /root/reference/main.py:669-722 (``aggregate_failure_analysis``), literal."""
from collections import defaultdict
from typing import Any, Dict, List


def aggregate_failure_analysis(all_query_results: List[List[Dict]]) -> Dict[str, Any]:
    all_results = []
    for it in all_query_results:
        all_results.extend(it)
    if not all_results:
        return {}
    analysis: Dict[str, Any] = {}
    total = len(all_results)
    correct = sum(r["correct_r1"] for r in all_results)
    analysis["overall_accuracy"] = 100.0 * correct / total if total > 0 else 0.0
    for feature in ["has_spatial", "has_color", "has_counting"]:
        w = [r for r in all_results if r[feature]]
        wo = [r for r in all_results if not r[feature]]
        if w:
            analysis[f"accuracy_{feature}"] = 100.0 * sum(r["correct_r1"] for r in w) / len(w)
        if wo:
            analysis[f"accuracy_not_{feature}"] = 100.0 * sum(r["correct_r1"] for r in wo) / len(wo)
    cat = defaultdict(list)
    for r in all_results:
        if r["category"]:
            cat[r["category"]].append(r["correct_r1"])
    analysis["accuracy_by_category"] = {c: 100.0 * sum(v) / len(v) for c, v in cat.items() if v}
    for low, high in [(0, 5), (6, 10), (11, 15), (16, 100)]:
        b = [r for r in all_results if low <= r["caption_length"] <= high]
        if b:
            analysis[f"accuracy_length_{low}_{high}"] = 100.0 * sum(r["correct_r1"] for r in b) / len(b)
    return analysis

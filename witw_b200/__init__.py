"""witw_b200 -- B200 (sm_100a) kernels for the cross-view retrieval hot path of IQTLabs/WITW.

Polar transform, orientation-searched correlation / crop / distance, rank and top-k
evaluation, behind the reference's own function names (see ops.py), implemented as
hand-written CUDA in libwitw_b200.so (C ABI: include/witw_b200.h).  No CPU fallback.
"""
from . import _lib, ops
from ._lib import WitwError
from .ops import (
    GalleryBuilder,
    GalleryIndex,
    PolarTransform,
    QueryBatch,
    baseline_ranks,
    bilinear_interpolate,
    correlation,
    correlation_scores,
    crop_overhead,
    evaluate_ranks,
    evaluate_ranks_prepared,
    heatmap_scores,
    l2_distance,
    match,
    normalized_polar,
    polar_grid,
    polar_transform,
    rank_from_distances,
    recall_from_ranks,
    sweep_tc,
    tc_supported,
    topk_from_distances,
    true_match_distances,
)
from .install import install, uninstall
from .sharded import evaluate_ranks_sharded, shard_bounds

__all__ = [
    "GalleryBuilder", "GalleryIndex", "PolarTransform", "QueryBatch", "WitwError", "baseline_ranks", "bilinear_interpolate", "correlation",
    "correlation_scores", "crop_overhead", "evaluate_ranks", "evaluate_ranks_prepared", "evaluate_ranks_sharded",
    "heatmap_scores", "install", "l2_distance", "match", "normalized_polar", "polar_grid", "polar_transform", "rank_from_distances",
    "recall_from_ranks", "shard_bounds", "sweep_tc", "tc_supported", "topk_from_distances", "true_match_distances",
    "uninstall",
]

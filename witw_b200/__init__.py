"""witw_b200 -- B200 (sm_100a) kernels for the cross-view retrieval hot path of IQTLabs/WITW.

Polar transform, orientation-searched correlation / crop / distance, rank and top-k
evaluation, behind the reference's own function names (see ops.py), implemented as
hand-written CUDA in libwitw_b200.so (C ABI: include/witw_b200.h).  No CPU fallback.
"""
from . import _lib, ops
from ._lib import WitwError
from .ops import (
    Deferral,
    RankEvaluation,
    GalleryBuilder,
    GalleryIndex,
    ImageNormalization,
    PolarTransform,
    QueryBatch,
    Resize,
    baseline_ranks,
    bilinear_interpolate,
    correlation,
    correlation_scores,
    crop_overhead,
    evaluate_ranks,
    evaluate_ranks_prepared,
    exact_columns,
    finish_tc,
    heatmap_scores,
    l2_distance,
    match,
    match_distance,
    normalized_polar,
    polar_grid,
    polar_transform,
    prepare_pair,
    rank_from_distances,
    recall_from_ranks,
    resize_normalize,
    sweep_tc,
    tc_supported,
    topk_from_distances,
    triplet_loss,
    true_match_distances,
)
from .heatmap import heatmap_sweep, prefetch_to_device, prepare_tiles, streamed_polar
from .install import install, uninstall
from . import peer, sharded
from .sharded import ShardedEvaluation, evaluate_ranks_sharded, shard_bounds

__all__ = [
    "ShardedEvaluation", "heatmap_sweep", "prefetch_to_device", "prepare_tiles", "streamed_polar", "RankEvaluation", "Deferral", "exact_columns", "finish_tc", "GalleryBuilder", "GalleryIndex", "ImageNormalization", "PolarTransform", "QueryBatch", "Resize", "WitwError", "baseline_ranks", "bilinear_interpolate", "correlation",
    "correlation_scores", "crop_overhead", "evaluate_ranks", "evaluate_ranks_prepared", "evaluate_ranks_sharded",
    "heatmap_scores", "install", "l2_distance", "match", "match_distance", "normalized_polar", "polar_grid", "polar_transform", "prepare_pair", "rank_from_distances",
    "recall_from_ranks", "resize_normalize", "shard_bounds", "sweep_tc", "tc_supported", "topk_from_distances", "triplet_loss", "true_match_distances",
    "uninstall",
]

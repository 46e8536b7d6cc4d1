"""ctypes binding of libwitw_b200.so (C ABI declared in include/witw_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is
raised.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C witw_b200/csrc``.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwitw_b200.so")

_lib = None

# name -> (restype, argtypes); must list every symbol include/witw_b200.h declares
SIGNATURES = {
    "witw_last_error": (c_char_p, []),
    "witw_version": (c_int, []),
    "witw_device_check": (c_int, []),
    "witw_stream_l2_window": (c_int, [c_void_p, c_size_t, c_void_p]),
    "witw_polar_grid": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p]),
    "witw_bilinear_lut": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "witw_bilinear_gather_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p]),
    "witw_polar_plan_bytes": (c_size_t, [c_int, c_int, c_int]),
    "witw_polar_plan_build": (c_int, [c_int, c_int, c_int, c_void_p]),
    "witw_polar_resample_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "witw_norm_lut": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "witw_bilinear_gather_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_int, c_void_p]),
    "witw_polar_plan_bytes_u8": (c_size_t, [c_int, c_int, c_int]),
    "witw_polar_plan_build_u8": (c_int, [c_int, c_int, c_int, c_void_p]),
    "witw_polar_resample_u8": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "witw_resize_plan_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "witw_resize_plan_build": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "witw_resize_norm": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "witw_match_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "witw_match_pairs_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "witw_crop_gather_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "witw_l2_distance_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "witw_crop_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "witw_l2_distance_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "witw_match_backward_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                        c_int, c_void_p]),
    "witw_triplet_loss_f32": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "witw_gallery_operand_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "witw_query_operand_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "witw_gallery_prep": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "witw_query_prep": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "witw_match_tc_topk_slots": (c_int, [c_int64, c_int64]),
    "witw_match_tc": (c_int, [c_void_p, c_void_p]),
    "witw_spectral_rows_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "witw_match_pairs_spec_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                          c_void_p]),
    "witw_match_columns_spec_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                                            c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "witw_finish_spec_f32": (c_int, [c_void_p, c_void_p]),
    "witw_finish_scratch_bytes": (c_size_t, [c_int64, c_int]),
    "witw_sizeof_sweep_args": (c_size_t, []),
    "witw_sizeof_finish_args": (c_size_t, []),
    "witw_spec_supported": (c_int, [c_int, c_int, c_int]),
    "witw_spec_gallery_operand_bytes": (c_size_t, [c_int64, c_int]),
    "witw_spec_query_operand_bytes": (c_size_t, [c_int64, c_int]),
    "witw_spec_gallery_prep": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p]),
    "witw_spec_query_prep": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "witw_match_spec_topk_slots": (c_int, [c_int64, c_int64]),
    "witw_match_spec": (c_int, [c_void_p, c_void_p]),
    "witw_match_spec_variant": (c_int, [c_int]),
    "witw_peer_exchange_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "witw_peer_alloc": (c_int, [c_size_t, c_void_p, c_void_p]),
    "witw_peer_open": (c_int, [c_void_p, c_void_p]),
    "witw_peer_close": (c_int, [c_void_p]),
    "witw_peer_free": (c_int, [c_void_p]),
    "witw_peer_thresholds": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_int, c_int, c_uint32, c_void_p, c_void_p]),
    "witw_peer_results": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_uint32,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "witw_rank_from_dist_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "witw_l2_rank_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "witw_topk_slices": (c_int, [c_int64, c_int64]),
    "witw_topk_from_dist_f32": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_int32, c_void_p]),
    "witw_topk_select_scratch_bytes": (c_size_t, [c_int64, c_int]),
    "witw_topk_select_f32": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "witw_topk_merge": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
}


class SweepArgs(ctypes.Structure):
    """witw_sweep_args of include/witw_b200.h (field for field)."""
    _fields_ = [
        ("gal_op", c_void_p), ("gal_scale", c_void_p), ("gal_aux", c_void_p), ("qry_op", c_void_p), ("qry_aux", c_void_p),
        ("G", c_int64), ("Q", c_int64), ("CH", c_int32), ("sw", c_int32), ("g_index_offset", c_int32), ("topk", c_int32),
        ("dist", c_void_p), ("ori", c_void_p), ("d_true", c_void_p), ("true_idx", c_void_p), ("rank_count", c_void_p),
        ("topk_key", c_void_p), ("topk_idx", c_void_p), ("list_g", c_void_p), ("list_n", c_void_p),
        ("list_cap", c_int32), ("err_sigmas", c_float), ("fix_rel", c_float),
    ]


class FinishArgs(ctypes.Structure):
    """witw_finish_args of include/witw_b200.h (field for field)."""
    _fields_ = [
        ("gal_spec", c_void_p), ("crop_inv_norm", c_void_p), ("qry_spec", c_void_p), ("q_inv_norm", c_void_p),
        ("G", c_int64), ("Q", c_int64), ("CH", c_int32), ("g_index_offset", c_int32),
        ("list_g", c_void_p), ("list_n", c_void_p), ("list_cap", c_int32), ("kc", c_int32),
        ("d_true", c_void_p), ("rank_count", c_void_p), ("dist", c_void_p), ("ori", c_void_p),
        ("cand_key", c_void_p), ("cand_idx", c_void_p), ("out_dist", c_void_p), ("out_idx", c_void_p),
        ("k_out", c_int32), ("reserved", c_int32), ("qflag", c_void_p), ("n_flagged", c_void_p), ("scratch", c_void_p),
    ]


class WitwError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "witw_b200: %s is missing -- build it with `make -C witw_b200/csrc` "
            "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.witw_sizeof_sweep_args() != ctypes.sizeof(SweepArgs) or lib.witw_sizeof_finish_args() != ctypes.sizeof(FinishArgs):
        raise ImportError("witw_b200: %s was built from another include/witw_b200.h (argument structures differ); rebuild it" % LIB_PATH)
    _lib = lib
    return lib


def last_error():
    return load().witw_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise WitwError("%s failed (%d): %s" % (what or "libwitw_b200 call", rc, last_error()))


def call(name, *args):
    """Call a C-ABI function that returns a status code and raise on failure."""
    check(getattr(load(), name)(*args), name)

"""ctypes binding of libfeta_b200.so (the C ABI declared in include/feta_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C feta_tmlr_b200/csrc``.
There is NO CPU fallback: if the shared object is missing the import of any op raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfeta_b200.so")

_P = c_void_p  # every device pointer crosses as an opaque address

# name -> (restype, argtypes); mirrors include/feta_b200.h one to one
SIGNATURES = {
    "feta_version": (c_int, []),
    "feta_last_error_string": (c_char_p, []),
    "feta_launch_count": (c_int64, []),
    "feta_cheb_plan_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "feta_cheb_plan_build": (c_int, [_P, c_int64, _P, c_int, c_int64, c_int64, c_float,
                                     _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "feta_graph_plan_build": (c_int, [_P, c_int64, _P, c_int, c_int64, c_int64, c_int, c_float,
                                      _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "feta_cheb_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "feta_cheb_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P,
                              c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "feta_cheb_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64,
                              _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int,
                              _P, c_size_t, _P]),
    "feta_arma_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                              c_int64, c_int64, c_int, c_int, c_int, _P]),
    "feta_arma_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                              c_int64, c_int64, c_int, c_int, c_int, _P]),
    "feta_attn_fwd": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, c_int64, c_int64, _P,
                              c_int, c_int, c_int, c_int, c_float, c_int, _P]),
    "feta_attn_bwd": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P, _P,
                              c_int64, c_int64, c_int, c_int, c_int, c_int, c_float, _P]),
    "feta_attn_fwd_dropout": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P,
                                      c_int, c_int, c_int, c_int, c_float, _P]),
    "feta_attn_bwd_dropout": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P,
                                      _P, c_int64, c_int64, c_int, c_int, c_int, c_int, c_float, _P]),
    "feta_static_context": (c_int, [_P, _P, c_int, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "feta_colsum_partial_floats": (c_int64, [c_int]),
    "feta_colsum": (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    "feta_attn_rows_supported": (c_int, [c_int, c_int]),
    "feta_attn_rows_coeff": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float,
                                     c_int64, _P]),
    "feta_attn_rows_fwd": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, c_int64, c_int64, _P,
                                   c_int, c_int, c_int, c_int, c_float, _P]),
    "feta_attn_rows_bwd": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P,
                                   c_int64, c_int64, c_int, c_int, c_int, c_int, c_float, _P]),
    "feta_linear_wgrad_slices": (c_int, [c_int64]),
    "feta_linear_wgrad": (c_int, [_P, _P, _P, _P, _P, c_size_t, _P, c_int64, c_int, c_int, _P]),
    "feta_linear_tc_supported": (c_int, [c_int, c_int]),
    "feta_linear_tc5_supported": (c_int, [c_int, c_int]),
    "feta_layer_tail_supported": (c_int, [c_int, c_int]),
    "feta_layer_tail_fwd": (c_int, [_P] * 22 + [c_int64, c_int, c_int, c_float, c_float, _P]),
    "feta_linear_layernorm_supported": (c_int, [c_int, c_int]),
    "feta_linear_layernorm_simt_supported": (c_int, [c_int, c_int]),
    "feta_lnbwd_linear_dx_blocks": (c_int, [c_int64]),
    "feta_lnbwd_linear_dx": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "feta_ln_fold": (c_int, [_P, c_int, c_int, _P, _P, _P]),
    "feta_linear_layernorm_fwd_ex": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float,
                                             c_int, _P]),
    "feta_linear_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float, _P]),
    "feta_linear_simt_supported": (c_int, [c_int, c_int]),
    "feta_linear_fwd_ex": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "feta_linear_dx_ex": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "feta_linear_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "feta_linear_dx": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "feta_add_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, _P]),
    "feta_add_layernorm_bwd_blocks": (c_int, [c_int64]),
    "feta_add_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "feta_add_layernorm_bwd_fold": (c_int, [_P, c_int64, c_int, _P, _P, _P]),
    "feta_add_batchnorm_blocks": (c_int, [c_int64]),
    "feta_add_batchnorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_float, c_float,
                                       c_int64, c_int, _P]),
    "feta_add_batchnorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "feta_adam_step": (c_int, [_P, _P, _P, _P, c_int64, _P, c_float, c_float, c_float, c_float, c_float, _P, _P]),
    "feta_coeff_scalar": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int64, _P]),
    "feta_coeff_pool_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "feta_coeff_pool_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, _P]),
    "feta_pack_heads": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "feta_pack_heads_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "feta_unpack_heads": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "feta_unpack_heads_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, c_int, _P]),
    "feta_segment_mean_fwd": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "feta_segment_mean_bwd": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "feta_masked_mean_fwd": (c_int, [_P, c_int64, c_int64, _P, _P, c_int, c_int, c_int, _P]),
    "feta_masked_mean_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "feta_gather_rows": (c_int, [_P, c_int64, c_int64, _P, _P, c_int64, c_int, _P]),
    "feta_scatter_rows": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int, _P]),
    "feta_collate_indices": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P,
                                     c_int, c_int, c_int64, c_int64, _P]),
    "feta_collate_edges_static": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, c_int, c_int64, c_int64, _P]),
    "feta_collate_pad_rows": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "feta_collate_pad_pe": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, _P]),
}


class FetaError(RuntimeError):
    pass


_lib = None


def load():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "feta_tmlr_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C feta_tmlr_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().feta_last_error_string().decode("utf-8", "replace")
        raise FetaError("%s failed (code %d): %s" % (what, rc, msg))


def launch_count():
    return int(load().feta_launch_count())

"""ctypes binding of the C ABI in include/abcsmc_b200.h (libabcsmc_b200.so, hand-written sm_100a kernels).

There is no CPU fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libabcsmc_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "abcsmc_b200.h")

_dp = C.POINTER(C.c_double)
_i64 = C.c_int64
_vp = C.c_void_p

# name -> (restype, argtypes); pointers are passed as void* (host or device addresses)
_SIGS = {
    "abcb200_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "abcb200_destroy": (C.c_int, [_vp]),
    "abcb200_set_stream": (C.c_int, [_vp, _vp]),
    "abcb200_synchronize": (C.c_int, [_vp]),
    "abcb200_last_error": (C.c_char_p, [_vp]),
    "abcb200_launch_count": (C.c_uint64, [_vp]),
    "abcb200_exact_test_count": (C.c_uint64, [_vp]),
    "abcb200_stat": (C.c_uint64, [_vp, C.c_int]),
    "abcb200_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "abcb200_host_free": (C.c_int, [_vp]),
    "abcb200_set_timers": (C.c_int, [_vp, C.c_int, C.c_uint32]),
    "abcb200_stage_ms": (C.c_double, [_vp, C.c_int]),
    "abcb200_kernel_ms": (C.c_double, [_vp, C.c_int]),
    "abcb200_rank_pls": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, C.c_int, _vp, C.c_double, C.c_int, _i64, _vp, _vp, _vp, _vp]),
    "abcb200_rank_pls_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, C.c_int, _vp, C.c_double, C.c_int, _i64, _vp, _vp, _vp, _vp]),
    "abcb200_rank_simple": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp, _i64, _vp, _vp]),
    "abcb200_rank_simple_dev": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp, _i64, _vp, _vp]),
    "abcb200_doubled_variance": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_doubled_variance_dev": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_doubled_variance_gather_dev": (C.c_int, [_vp, _vp, _i64, _vp, _i64, C.c_int, _vp, _vp]),
    "abcb200_sample_predictive_priors": (C.c_int, [_vp, C.c_uint64, _i64, _vp, _vp, _i64, _i64, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp]),
    "abcb200_sample_predictive_priors_dev": (C.c_int, [_vp, C.c_uint64, _i64, _vp, _vp, _i64, _i64, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp]),
    "abcb200_setup_mvn_sampler": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_sample_mvn_predictive_priors": (C.c_int, [_vp, C.c_uint64, _i64, _vp, _vp, _i64, _i64, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _i64, _vp, _vp]),
    "abcb200_weights_set0": (C.c_int, [_vp, _i64, _vp]),
    "abcb200_weights": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, C.c_int, C.c_int, _vp]),
    "abcb200_weights_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, C.c_int, C.c_int, _vp]),
    "abcb200_weights_unnorm_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "abcb200_scale_weights_dev": (C.c_int, [_vp, _vp, _i64, _vp]),
    "abcb200_group_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "abcb200_group_unique_id": (C.c_int, [_vp, C.c_size_t]),
    "abcb200_group_create_rank": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "abcb200_group_destroy": (C.c_int, [_vp]),
    "abcb200_group_size": (C.c_int, [_vp]),
    "abcb200_group_local_size": (C.c_int, [_vp]),
    "abcb200_group_ctx": (_vp, [_vp, C.c_int]),
    "abcb200_group_last_error": (C.c_char_p, [_vp]),
    "abcb200_weights_sharded": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, C.c_int, C.c_int, _vp]),
    "abcb200_weights_sharded_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "abcb200_chain_create": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "abcb200_chain_destroy": (C.c_int, [_vp]),
    "abcb200_chain_sets": (C.c_int, [_vp]),
    "abcb200_chain_nparams": (C.c_int, [_vp]),
    "abcb200_set_tie_order": (C.c_int, [_vp, C.c_int]),
    "abcb200_tie_order_stdsort": (C.c_int, [_vp, _i64, _i64, _vp]),
    "abcb200_chain_process_set": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, _vp, C.c_int, C.c_double, C.c_int, _i64, _vp, _vp, _vp, _vp,
                                            _vp, _vp, _vp, _vp, _vp]),
    "abcb200_chain_state": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "abcb200_chain_restore": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, C.c_int]),
    "abcb200_db_last_error": (C.c_char_p, []),
    "abcb200_db_set_shape": (C.c_int, [C.c_char_p, C.c_int, _vp, _vp, _vp]),
    "abcb200_db_load_set": (C.c_int, [C.c_char_p, C.c_int, _i64, C.c_int, C.c_int, _vp, _i64, _vp, _i64, _vp, _vp]),
    "abcb200_db_write_ranks": (C.c_int, [C.c_char_p, _vp, _i64]),
    "abcb200_chain_process_db_set": (C.c_int, [_vp, C.c_char_p, C.c_int, _vp, C.c_int, C.c_double, C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "abcb200_colwise_moments": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp, _vp]),
    "abcb200_colwise_z_scores": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp, _vp, _vp, _i64]),
    "abcb200_gram": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, C.c_int, _vp, _vp]),
    "abcb200_euclidean": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp, _vp]),
    "abcb200_ordered": (C.c_int, [_vp, _vp, _i64, _vp]),
    "abcb200_ordered_top": (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "abcb200_wilcoxon": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "abcb200_pls_fit": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "abcb200_pls_free": (C.c_int, [_vp]),
    "abcb200_pls_get": (C.c_int, [_vp, C.c_char, _vp]),
    "abcb200_pls_scores": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_pls_coefficients": (C.c_int, [_vp, C.c_int, _vp]),
    "abcb200_pls_fitted_values": (C.c_int, [_vp, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_pls_residuals": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_pls_sse": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_pls_explained_variance": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, _vp]),
    "abcb200_residual_select": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp]),
    "abcb200_pls_cv_loo": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp, _vp]),
    "abcb200_pls_cv_lso": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _i64, _i64, C.c_int, C.c_double,
                                     _vp, _vp, _vp]),
    "abcb200_pls_cv_new_data": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, C.c_int, C.c_double, _vp, _vp]),
}

_lib = None


class Abcb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"abcb200 error {code}: {msg}")
        self.code = code


def declared_symbols():
    """Entry points declared in include/abcsmc_b200.h."""
    with open(HEADER_PATH) as f:
        txt = f.read()
    return sorted(set(re.findall(r"\b(abcb200_[a-z0-9_]+)\s*\(", txt)))


def build(force=False):
    """Compile libabcsmc_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib

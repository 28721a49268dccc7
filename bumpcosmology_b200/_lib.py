"""ctypes binding of the C ABI declared in include/bump.h (libbump_b200.so, built in-tree).

There is no CPU fallback: if the shared library is missing this module raises, and if no sm_100 GPU is
visible `bump_ctx_create` fails with BUMP_E_NOGPU.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BUMP_LIB_PATH", os.path.join(_PKG, "libbump_b200.so"))   # override: tuning builds only

NTHETA = 14
NTHETA_MAX = 15
OUT_LOGLIKE, OUT_LOG_MU_SEL, OUT_LOG_MU2, OUT_NEFF_SEL = 0, 1, 2, 3
OUT_DLOGLIKE, OUT_DLOG_MU = 4, 19
OUT_NVALID_EVT, OUT_NVALID_SEL, OUT_NOBS, OUT_NSEL, OUT_STATUS = 34, 35, 36, 37, 38
E_INVALID, E_CUDA, E_NOGPU, E_NCCL, E_EXCHANGE = 1, 2, 3, 4, 5
OUT_HEADER = 40
PARTIAL_LEN = 128
FLAG_WA, FLAG_NO_GRAPH, FLAG_NO_SORT, FLAG_FIXED_COSMO = 1, 2, 4, 8

_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes); exactly the entry points of include/bump.h
SIGNATURES = {
    "bump_version": (C.c_char_p, []),
    "bump_last_error": (C.c_char_p, []),
    "bump_device_count": (C.c_int, []),
    "bump_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint32]),
    "bump_ctx_destroy": (None, [C.c_void_p]),
    "bump_ctx_clone": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "bump_upload_events": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, _dp, _dp, _dp, _dp]),
    "bump_upload_injections": (C.c_int, [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, C.c_double]),
    "bump_set_fixed_dvdzdt": (C.c_int, [C.c_void_p, _dp, C.c_int64]),
    "bump_out_len": (C.c_int64, [C.c_void_p]),
    "bump_eval": (C.c_int, [C.c_void_p, _dp, _dp]),
    "bump_eval_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bump_eval_partial": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "bump_merge_partials": (C.c_int, [_dp, C.c_int, _dp]),
    "bump_eval_partial_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bump_finalize_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bump_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "bump_comm_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "bump_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bump_p2p_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "bump_p2p_detach": (C.c_int, [C.c_void_p]),
    "bump_p2p_set_timeout": (C.c_int, [C.c_void_p, C.c_double]),
    "bump_debug_warp_times": (C.c_int, [C.c_void_p, _dp, C.c_int64]),
    "bump_debug_timeline": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int64]),
    "bump_debug_tables": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int64]),
    "bump_debug_math": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int64, _dp]),
    "bump_time_evals": (C.c_int, [C.c_void_p, _dp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "bump_launches_per_eval": (C.c_int, [C.c_void_p]),
    "bump_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "bump_ctx_flags": (C.c_int, [C.c_void_p]),
    "bump_nuts_chain": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_double, C.c_int, _dp, _dp,
                                  _dp, _dp, _dp, _dp, _dp]),
    "bump_nuts_potential": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "bump_nuts_prior_terms": (C.c_int, [_dp, _dp, _dp, _dp]),
    "bump_nuts_chain_cb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                     C.c_double, C.c_int, _dp, _dp, _dp, _dp, _dp]),
}
NUTS_NSTAT, NUTS_NDET = 5, 8
POTENTIAL_CB = C.CFUNCTYPE(C.c_double, C.c_void_p, _dp, _dp)   # double (*)(void* user, const double* u, double* grad)

_lib = None


class BumpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bump_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load libbump_b200.so (idempotent).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m bumpcosmology_b200._build` "
                          "(needs nvcc; sm_100a). bump_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise BumpError(code, load().bump_last_error().decode())


def as_dp(a):
    return a.ctypes.data_as(_dp)

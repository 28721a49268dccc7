"""ctypes wrapper of oracle/bump_cpu.cpp (fused C++/OpenMP CPU port) — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__ and bench.py's CPU legs may import this.  `make -C oracle` builds the library."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libbump_cpu.so")
_dp = C.POINTER(C.c_double)
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.bcpu_create.restype = C.c_void_p
        lib.bcpu_create.argtypes = [C.c_int64, C.c_int64, _dp, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp, C.c_double]
        lib.bcpu_destroy.argtypes = [C.c_void_p]
        lib.bcpu_eval.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        lib.bcpu_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def _c(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class CpuPort:
    def __init__(self, m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw):
        self.lib = load()
        ev = [_c(x) for x in (m1s_det, qs, dls, pdraw)]
        sel = [_c(x).ravel() for x in (m1s_det_sel, qs_sel, dls_sel, pdraw_sel)]
        self.nobs, self.nsamp = ev[0].shape
        self.nsel = sel[0].shape[0]
        self._h = C.c_void_p(self.lib.bcpu_create(self.nobs, self.nsamp, *[x.ctypes.data_as(_dp) for x in ev],
                                                  self.nsel, *[x.ctypes.data_as(_dp) for x in sel], float(Ndraw)))
        self._out = np.empty(40 + self.nobs)
        self.threads = int(self.lib.bcpu_max_threads())

    def evaluate(self, theta, nthreads=0):
        th = _c(theta)[:14].copy()
        self.lib.bcpu_eval(self._h, th.ctypes.data_as(_dp), self._out.ctypes.data_as(_dp), int(nthreads))
        o = self._out
        nobs = self.nobs
        return {"loglike": float(o[0]), "log_mu_sel": float(o[1]), "log_mu2": float(o[2]), "neff_sel": float(o[3]),
                "dloglike": o[4:18].copy(), "dlog_mu_sel": o[19:33].copy(), "neff": o[40:].copy(), "nobs": nobs,
                "selfactor": -nobs * float(o[1]), "logl": float(o[0]) - nobs * float(o[1])}

    def close(self):
        if self._h:
            self.lib.bcpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""ctypes wrapper of oracle/bump_cpu.cpp (fused C++/OpenMP CPU port) — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__ and bench.py's CPU legs may import this.  `make -C oracle` builds the library."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libbump_cpu.so")                 # portable build (x86-64-v3)
NATIVE_LIB_PATH = os.path.join(_HERE, "_build", "libbump_cpu_native.so")   # -march=native, built on the run host
_NATIVE_STAMP = os.path.join(_HERE, "_build", "native.stamp")
_dp = C.POINTER(C.c_double)
_lib = None
loaded_path = None


def host_threads():
    """Threads this process may use: its affinity mask, not OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def _host_cpu_id():
    """Model name + ISA flags of this host: a -march=native build is only valid where it was made."""
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
    except OSError:
        return "unknown"
    keep = [ln.split(":", 1)[1].strip() for ln in txt.splitlines() if ln.startswith(("model name", "flags"))][:2]
    return " | ".join(keep)


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return LIB_PATH


def build_native():
    """-march=native build for THIS host (rebuilt when the host's CPU changed since the last build).  Returns the
    path, or None if it cannot be built here (no compiler)."""
    cpu = _host_cpu_id()
    try:
        if os.path.exists(NATIVE_LIB_PATH) and os.path.exists(_NATIVE_STAMP):
            with open(_NATIVE_STAMP) as f:
                if f.read() == cpu:
                    return NATIVE_LIB_PATH
        subprocess.run(["make", "-s", "-C", _HERE, "native"], check=True, capture_output=True, timeout=300)
        with open(_NATIVE_STAMP, "w") as f:
            f.write(cpu)
        return NATIVE_LIB_PATH
    except Exception:  # noqa: BLE001
        return None


def load(native=False):
    """native=True (bench.py's CPU arms): prefer the -march=native build made on this host; tests use the portable one."""
    global _lib, loaded_path
    if _lib is None:
        path = (build_native() if native else None) or LIB_PATH
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        loaded_path = path
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
    def __init__(self, m1s_det, qs, dls, pdraw, m1s_det_sel, qs_sel, dls_sel, pdraw_sel, Ndraw, native=False):
        self.lib = load(native)
        ev = [_c(x) for x in (m1s_det, qs, dls, pdraw)]
        sel = [_c(x).ravel() for x in (m1s_det_sel, qs_sel, dls_sel, pdraw_sel)]
        self.nobs, self.nsamp = ev[0].shape
        self.nsel = sel[0].shape[0]
        self._h = C.c_void_p(self.lib.bcpu_create(self.nobs, self.nsamp, *[x.ctypes.data_as(_dp) for x in ev],
                                                  self.nsel, *[x.ctypes.data_as(_dp) for x in sel], float(Ndraw)))
        self._out = np.empty(40 + self.nobs)
        self.threads = host_threads()

    def evaluate(self, theta, nthreads=None):
        """nthreads: OpenMP team size, passed explicitly on every call (default: every core this process may run on,
        whatever OMP_NUM_THREADS says)."""
        th = _c(theta)[:14].copy()
        n = self.threads if not nthreads else int(nthreads)
        self.lib.bcpu_eval(self._h, th.ctypes.data_as(_dp), self._out.ctypes.data_as(_dp), n)
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

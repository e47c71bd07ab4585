"""ctypes binding of libemrihost.so (csrc/emrihost.c): native host-side producers -- the trajectory ODE and the
Schwarzschild fundamental frequencies.  Optional: when the library has not been built the pure-Python/SciPy twins
in trajectory/inspiral.py and utils/utility.py are used (these are input producers, not the accelerated path)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libemrihost.so")
_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            return None
        lib = C.CDLL(LIB_PATH)
        lib.emrihost_schwarzschild_frequencies.argtypes = [_dp, _dp, C.c_int64, _dp, _dp]
        lib.emrihost_trajectory.argtypes = [C.c_double] * 9 + [C.c_int] + [_dp] * 7
        lib.emrihost_trajectory.restype = C.c_int
        lib.emrihost_trajectory_batch.argtypes = [C.c_int64] + [_dp] * 6 + [C.c_double] * 3 + [C.c_int] + [_dp] * 7 + [_ip]
        lib.emrihost_trajectory_batch_mt.argtypes = lib.emrihost_trajectory_batch.argtypes + [C.c_int]
        lib.emrihost_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def frequencies(p, e):
    lib = load()
    p = np.ascontiguousarray(p, dtype=np.float64)
    e = np.ascontiguousarray(e, dtype=np.float64)
    a, b = np.empty_like(p), np.empty_like(p)
    lib.emrihost_schwarzschild_frequencies(p.ravel(), e.ravel(), p.size, a.ravel(), b.ravel())
    return a, b


def trajectory_batch(M, mu, p0, e0, Phi_phi0, Phi_r0, T, rtol, atol, max_len, nthreads=1):
    """Integrate nb walkers (nthreads > 1: a per-call pthread team).  Returns (arrays [nb, max_len] x 7: t, p, e, Phi_phi,
    Phi_r, f_phi [Hz], f_r [Hz]; lens [nb], negative = error code)."""
    lib = load()
    f = lambda x: np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    M, mu, p0, e0, Phi_phi0, Phi_r0 = map(f, (M, mu, p0, e0, Phi_phi0, Phi_r0))
    nb = len(M)
    out = [np.empty((nb, max_len)) for _ in range(7)]
    lens = np.zeros(nb, dtype=np.int32)
    if nthreads > 1 and nb > 1:
        lib.emrihost_trajectory_batch_mt(nb, M, mu, p0, e0, Phi_phi0, Phi_r0, float(T), float(rtol), float(atol), int(max_len),
                                         *out, lens, int(nthreads))
    else:
        lib.emrihost_trajectory_batch(nb, M, mu, p0, e0, Phi_phi0, Phi_r0, float(T), float(rtol), float(atol), int(max_len), *out, lens)
    return out, lens

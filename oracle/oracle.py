"""ctypes front-end of the CPU oracle (oracle/emrifd_oracle.c).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  "parity unpinned" against
FastEMRIWaveforms itself (not installable here); pinned against SciPy/mpmath/lisatools-derived goldens.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
MAXBR = 4

BRANCH_DTYPE = np.dtype([
    ("mode", np.int32), ("dir", np.int32), ("ja", np.int32), ("jb", np.int32),
    ("closed_end", np.int32), ("pad", np.int32), ("start", np.int64), ("end", np.int64),
    ("xa", np.float64), ("xb", np.float64), ("Fa", np.float64), ("Fb", np.float64)])


def build(force=False):
    """Compile the oracle with its Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "emrifd_oracle.c")
    libs = [os.path.join(_BUILD, f) for f in ("liboracle_f64.so", "liboracle_quad.so")]
    need = force or not all(os.path.exists(f) and os.path.getmtime(f) >= os.path.getmtime(src) for f in libs)
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class Oracle:
    """One precision flavour of the oracle: 'quad' (binary128 evaluation, the truth) or 'f64'."""

    def __init__(self, flavour="quad"):
        build()
        self.flavour = flavour
        self.lib = C.CDLL(os.path.join(_BUILD, f"liboracle_{flavour}.so"))
        L = self.lib
        assert L.orc_sizeof_branch() == BRANCH_DTYPE.itemsize
        L.orc_spline_build.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp]
        L.orc_spline_eval.argtypes = [_dp, _dp, C.c_int, C.c_int, _dp, C.c_int64, _dp]
        L.orc_segment_build.argtypes = [_dp, _dp, C.c_int, C.c_int, _ip, _ip, C.c_int64, C.c_double,
                                        C.c_void_p, C.c_void_p, _ip]
        L.orc_mode_sum.argtypes = [_dp, _dp, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int64, C.c_double,
                                   C.c_void_p, C.c_void_p, _ip, C.c_int, C.c_double, C.c_double,
                                   C.c_double, C.c_int64, C.c_int64, _dp, _dp, C.c_void_p]
        L.orc_inner_product.argtypes = [_dp, _dp, C.c_int, C.c_int64, _dp, C.c_void_p]
        L.orc_inner_product.restype = C.c_double
        L.orc_loglike.argtypes = [_dp, _dp, _dp, C.c_int, C.c_int64, _dp]
        L.orc_spa_R.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_spa_S.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_set_k13_mode.argtypes = [C.c_int]

    def set_k13_mode(self, mode):
        """"exact" (default) or "few": FastEMRIWaveforms' SPAFunc truncations (14-term ascending series for |X| <= 7, 9-term
        asymptotic above; SURVEY.md A.2 [UPSTREAM-MEMORY])."""
        self.lib.orc_set_k13_mode({"exact": 0, "few": 1}[mode])

    # -- A3 ---------------------------------------------------------------------------------
    def spline_build(self, t, y):
        t = np.ascontiguousarray(t, dtype=np.float64)
        y = np.ascontiguousarray(np.atleast_2d(y), dtype=np.float64)
        R, L = y.shape
        coeff = np.zeros((L, R, 4))
        rc = self.lib.orc_spline_build(t, y, L, R, coeff)
        if rc:
            raise ValueError(f"oracle spline_build failed rc={rc}")
        return coeff

    def spline_eval(self, t, coeff, tnew):
        t = np.ascontiguousarray(t, dtype=np.float64)
        tnew = np.ascontiguousarray(tnew, dtype=np.float64)
        L, R, _ = coeff.shape
        out = np.zeros((R, len(tnew)))
        self.lib.orc_spline_eval(t, np.ascontiguousarray(coeff), L, R, tnew, len(tnew), out)
        return out

    # -- A4 ---------------------------------------------------------------------------------
    def segment_build(self, t, coeff, m_arr, n_arr, N, val=0.0, fpos=None):
        L, R, _ = coeff.shape
        K = (R - 4) // 2
        br = np.zeros((K, MAXBR), dtype=BRANCH_DTYPE)
        nbr = np.zeros(K, dtype=np.int32)
        fp = None if fpos is None else np.ascontiguousarray(fpos, dtype=np.float64)
        rc = self.lib.orc_segment_build(
            np.ascontiguousarray(t, dtype=np.float64), np.ascontiguousarray(coeff), L, K,
            np.ascontiguousarray(m_arr, dtype=np.int32), np.ascontiguousarray(n_arr, dtype=np.int32),
            int(N), float(val), None if fp is None else fp.ctypes.data, br.ctypes.data, nbr)
        if rc:
            raise ValueError(f"oracle segment_build failed rc={rc}")
        return br, nbr

    # -- A5-A7 ------------------------------------------------------------------------------
    def mode_sum(self, t, coeff, m_arr, n_arr, ylms, N, branches, nbr, val=0.0, fpos=None,
                 include_minus_m=True, scale=1.0, cos2psi=1.0, sin2psi=0.0, out_lo=0, out_n=None):
        L, R, _ = coeff.shape
        K = (R - 4) // 2
        out_n = int(N - out_lo if out_n is None else out_n)
        hp = np.zeros(2 * out_n)
        hc = np.zeros(2 * out_n)
        fp = None if fpos is None else np.ascontiguousarray(fpos, dtype=np.float64)
        nev = C.c_int64(0)
        yl = np.ascontiguousarray(ylms, dtype=np.complex128).view(np.float64)
        rc = self.lib.orc_mode_sum(
            np.ascontiguousarray(t, dtype=np.float64), np.ascontiguousarray(coeff), L, K,
            np.ascontiguousarray(m_arr, dtype=np.int32), np.ascontiguousarray(n_arr, dtype=np.int32),
            yl, int(N), float(val), None if fp is None else fp.ctypes.data,
            branches.ctypes.data, np.ascontiguousarray(nbr, dtype=np.int32), int(include_minus_m),
            float(scale), float(cos2psi), float(sin2psi), int(out_lo), out_n, hp, hc, C.addressof(nev))
        if rc:
            raise ValueError(f"oracle mode_sum failed rc={rc}")
        self.last_n_eval = nev.value
        return hp.view(np.complex128), hc.view(np.complex128)

    # -- A10/A11 ----------------------------------------------------------------------------
    def inner_product(self, a, b, freqs, psd=None):
        a = np.ascontiguousarray(np.atleast_2d(a), dtype=np.complex128)
        b = np.ascontiguousarray(np.atleast_2d(b), dtype=np.complex128)
        nch, n = a.shape
        ps = None if psd is None else np.ascontiguousarray(psd, dtype=np.float64)
        return self.lib.orc_inner_product(a.view(np.float64), b.view(np.float64), nch, n,
                                          np.ascontiguousarray(freqs, dtype=np.float64),
                                          None if ps is None else ps.ctypes.data)

    def loglike(self, d_whitened, h, noise_factor):
        d = np.ascontiguousarray(np.atleast_2d(d_whitened), dtype=np.complex128)
        hh = np.ascontiguousarray(np.atleast_2d(h), dtype=np.complex128)
        w = np.ascontiguousarray(np.atleast_2d(noise_factor), dtype=np.float64)
        nch, n = d.shape
        out = np.zeros(3)
        self.lib.orc_loglike(d.view(np.float64), hh.view(np.float64), w, nch, n, out)
        return out  # ll, 4*sum Re(d* h w), 4*sum |h w|^2

    def spa_R(self, X):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_spa_R(float(X), C.byref(a), C.byref(b))
        return complex(a.value, b.value)

    def spa_S(self, X):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_spa_S(float(X), C.byref(a), C.byref(b))
        return complex(a.value, b.value)

    # -- whole FDInterpolatedModeSum.sum restated -------------------------------------------
    def fd_sum(self, t, teuk_modes, ylms, Phi_phi, Phi_r, m_arr, n_arr, f_phi, f_r, N, val=0.0,
               fpos=None, **kw):
        """y_all rows = [Re A | Im A | f_phi, f_r, Phi_phi, Phi_r]; returns (hp, hc, coeff, branches, nbr)."""
        y = np.concatenate([teuk_modes.T.real, teuk_modes.T.imag,
                            np.stack([f_phi, f_r, Phi_phi, Phi_r])])
        coeff = self.spline_build(t, y)
        br, nbr = self.segment_build(t, coeff, m_arr, n_arr, N, val, fpos)
        hp, hc = self.mode_sum(t, coeff, m_arr, n_arr, ylms, N, br, nbr, val, fpos, **kw)
        return hp, hc, coeff, br, nbr


# ---- mode selection by power (numpy restatement; checker only) ----------------------------
def mode_select_ref(teuk_modes, ylms, m0mask, eps):
    """few.utils.modeselector.ModeSelector.__call__ semantics (call site
    Tutorial_FrequencyDomain_Waveforms.ipynb:122-131 through ``eps=``), with the summation orders frozen so that a
    device implementation can be compared index-for-index:
    power = |[A, conj(A[:, m>0])] * ylms|^2 (each operation rounded); per time sample sort descending (ties: ascending
    index), sequential cumsum, keep entry s iff s == 0 or cumsum[s-1] < total * (1 - eps); union over samples; -m picks
    fold onto their +m partner.  ``total`` is summed as 256 strided sequential partials, a 32-lane shuffle-down tree
    per group of 32 partials, then sequentially over the 8 groups.  Returns the sorted kept +m indices."""
    teuk_modes = np.asarray(teuk_modes, dtype=np.complex128)
    m0mask = np.asarray(m0mask, dtype=bool)
    M = teuk_modes.shape[1]
    full = np.concatenate([teuk_modes, np.conj(teuk_modes[:, m0mask])], axis=1)
    ar, ai, yr, yi = full.real, full.imag, ylms.real[None, :], ylms.imag[None, :]
    re, im = ar * yr - ai * yi, ar * yi + ai * yr
    power = re * re + im * im
    ntot = power.shape[1]
    src = np.concatenate([np.arange(M), np.where(m0mask)[0]])
    keep = np.zeros(M, dtype=bool)
    for row in power:
        pad = np.zeros(((ntot + 255) // 256) * 256)
        pad[:ntot] = row
        part = np.zeros(256)
        for chunk in pad.reshape(-1, 256):        # sequential strided partials (adding the zero padding is exact)
            part = part + chunk
        v = part.reshape(8, 32).copy()
        for o in (16, 8, 4, 2, 1):                # shuffle-down tree: lane l += lane l+o (lanes >= 32-o add junk, unused)
            v[:, :o] = v[:, :o] + v[:, o:2 * o]
        total = 0.0
        for q in range(8):
            total = total + v[q, 0]
        thresh = total * (1.0 - eps)
        order = np.lexsort((np.arange(ntot), -row))
        cs = 0.0
        for s, i in enumerate(order):
            if s > 0 and not (cs < thresh):
                break
            keep[src[i]] = True
            cs = cs + row[i]
    return np.where(keep)[0]


def ylm_ref(l, m, theta, phi):
    """-2Y_lm(theta, phi) by the explicit Wigner-d sum with exact integer factorials (few.utils.ylm.GetYlms values;
    Tutorial_FD_construction_single_mode.ipynb:87).  Scalar, checker only."""
    from math import factorial, sqrt, pi, cos, sin
    mp, mm = m, 2
    cb, sb = cos(theta / 2.0), sin(theta / 2.0)
    tot = 0.0
    for k in range(max(0, mm - mp), min(l + mm, l - mp) + 1):
        den = factorial(l + mm - k) * factorial(k) * factorial(l - k - mp) * factorial(k - mm + mp)
        tot += (-1.0) ** (k - mm + mp) / den * cb ** (2 * l - 2 * k + mm - mp) * sb ** (2 * k - mm + mp)
    d = sqrt(factorial(l + mp) * factorial(l - mp) * factorial(l + mm) * factorial(l - mm)) * tot
    return sqrt((2 * l + 1) / (4.0 * pi)) * d * complex(cos(m * phi), sin(m * phi))


# ---- optimised CPU baseline (timed by bench.py only; validated against the oracle in tests/test_oracle_cpu.py) -------------
def _cpu_tag():
    """The fast baseline is compiled with -march=native, so its file name carries the host CPU model: a library built on
    another machine (the build container vs the GPU box) is never loaded."""
    import hashlib
    try:
        model = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except OSError:
        model = []
    return hashlib.sha1("".join(model).encode()).hexdigest()[:10]


class FastCPU:
    """ctypes front-end of oracle/emrifd_cpu_fast.c: one walker's f >= 0 waveform and/or likelihood sums on the implicit grid,
    from the oracle's (bit-exact) spline coefficients and work-list.  ``Oracle('f64')`` supplies those."""

    def __init__(self, oracle=None):
        self.orc = oracle or Oracle("f64")
        path = os.path.join(_BUILD, f"libcpufast_{_cpu_tag()}.so")
        src = os.path.join(_HERE, "emrifd_cpu_fast.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "cpufast", f"CPUFAST={path}"])
        self.lib = C.CDLL(path)
        self.lib.cpuf_sum.argtypes = [_dp, _dp, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int64, C.c_double, C.c_void_p, _ip, C.c_int,
                                      C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _dp,
                                      C.c_void_p]
        self.lib.cpuf_set_num_threads.argtypes = [C.c_int]

    def num_threads(self):
        return int(self.lib.cpuf_num_threads())

    def set_num_threads(self, n):
        self.lib.cpuf_set_num_threads(int(n))
        self.orc.lib.orc_set_num_threads(int(n))

    def prepare(self, it, N, val):
        """Spline coefficients + work-list of one walker (the oracle's exact double code)."""
        y = np.concatenate([it["teuk_modes"].T.real, it["teuk_modes"].T.imag, np.stack([it["f_phi"], it["f_r"], it["Phi_phi"], it["Phi_r"]])])
        coeff = self.orc.spline_build(it["t"], y)
        br, nbr = self.orc.segment_build(it["t"], coeff, it["m_arr"], it["n_arr"], N, val)
        return coeff, br, nbr

    def sum(self, it, N, val, data_w=None, wfac=None, want_h=True, include_minus_m=True, prepared=None):
        """Returns (hp, hc [npos] or None, like3, n_eval)."""
        coeff, br, nbr = prepared or self.prepare(it, N, val)
        L, R, _ = coeff.shape
        K = (R - 4) // 2
        npos = (N + 1) // 2
        hp = np.zeros(2 * npos) if want_h else None
        hc = np.zeros(2 * npos) if want_h else None
        like3 = np.zeros(3)
        nev = C.c_int64(0)
        dw = None if data_w is None else np.ascontiguousarray(data_w, dtype=np.complex128).view(np.float64)
        wf = None if wfac is None else np.ascontiguousarray(wfac, dtype=np.float64)
        ylm = np.ascontiguousarray(it["ylms"], dtype=np.complex128).view(np.float64)
        rc = self.lib.cpuf_sum(np.ascontiguousarray(it["t"], dtype=np.float64), np.ascontiguousarray(coeff), L, K,
                               np.ascontiguousarray(it["m_arr"], dtype=np.int32), np.ascontiguousarray(it["n_arr"], dtype=np.int32), ylm,
                               int(N), float(val), br.ctypes.data, np.ascontiguousarray(nbr, dtype=np.int32), int(include_minus_m),
                               float(it.get("scale", 1.0)), float(it.get("cos2psi", 1.0)), float(it.get("sin2psi", 0.0)),
                               None if dw is None else dw.ctypes.data, None if wf is None else wf.ctypes.data,
                               None if hp is None else hp.ctypes.data, None if hc is None else hc.ctypes.data, like3, C.addressof(nev))
        if rc:
            raise ValueError(f"cpuf_sum failed rc={rc}")
        return (hp.view(np.complex128) if want_h else None, hc.view(np.complex128) if want_h else None, like3, nev.value)

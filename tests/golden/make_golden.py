"""Generate the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference, scipy, mpmath):
    python tests/golden/make_golden.py

What pins what:
  lisatools_golden.npz  outputs of the REAL reference functions imported from /root/reference:
                        lisatools.diagnostic.inner_product (diagnostic.py:14-186),
                        lisatools.sampling.likelihood.Likelihood.inject_signal/__call__
                        (likelihood.py:80-334) with eryn's TransformContainer (transform.py:181-226),
                        on seeded random inputs, PSD = SciPy CubicSpline over LISA_Alloc_Sh.txt
                        exactly as FDutils.py:4-5 builds it (FDutils itself needs matplotlib).
  spline_golden.npz     SciPy CubicSpline (not-a-knot) coefficients + PSD spline values incl. f = 0.
  k13_golden.npz        K_{1/3}(-iX) e^{-iX} from mpmath (50 digits) and scipy.special.kv
                        (the notebook's evaluation, Tutorial_FD_construction_single_mode.ipynb:600).
  waveform_*.npz        sparse inputs + binary128-oracle h+, hx, work-list for small cases
                        (regression pins; the oracle itself is pinned by the files above and by the
                        FD-vs-FFT(TD) first-principles test).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def lisatools_td_golden():
    """inner_product(..., dt=) branch of the reference (diagnostic.py:49-67): rfft of the time series, DC dropped, zero padding
    of the shorter signal."""
    import warnings
    sys.path.insert(0, os.path.join(REF, "LISAanalysistools"))
    from lisatools.diagnostic import inner_product, snr
    from scipy.interpolate import CubicSpline
    S = np.genfromtxt(os.path.join(REF, "LISA_Alloc_Sh.txt"))
    Sh_X = CubicSpline(S[:, 0], S[:, 1])
    rng = np.random.default_rng(77)
    n, dt = 4096, 10.0
    t = np.arange(n) * dt
    x = [np.sin(2 * np.pi * 2e-3 * t + 0.3 * k) * 1e-20 + 1e-21 * rng.normal(size=n) for k in range(2)]
    y = [x[k] + 2e-21 * rng.normal(size=n) for k in range(2)]
    y_short = [c[:4000] for c in y]
    psd = Sh_X(np.fft.rfftfreq(n, dt)[1:])
    out = dict(x=np.asarray(x), y=np.asarray(y), dt=dt, psd=psd)
    out["ip_xy"] = inner_product(x, y, dt=dt, PSD=psd)
    out["ip_xy_norm"] = inner_product(x, y, dt=dt, PSD=psd, normalize=True)
    out["snr_x"] = snr(x, dt=dt, PSD=psd)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["ip_xy_short"] = inner_product(x, y_short, dt=dt, PSD=psd)
    np.savez(os.path.join(HERE, "lisatools_td_golden.npz"), **out)
    print("lisatools_td_golden.npz", out["ip_xy"], out["ip_xy_norm"], out["snr_x"], out["ip_xy_short"])


def lisatools_golden():
    sys.path.insert(0, os.path.join(REF, "LISAanalysistools"))
    sys.path.insert(0, os.path.join(REF, "Eryn"))
    from lisatools.diagnostic import inner_product, snr
    from lisatools.sampling.likelihood import Likelihood
    from eryn.utils import TransformContainer
    from scipy.interpolate import CubicSpline
    S = np.genfromtxt(os.path.join(REF, "LISA_Alloc_Sh.txt"))
    Sh_X = CubicSpline(S[:, 0], S[:, 1])           # FDutils.py:4-5
    get_sensitivity = lambda f: Sh_X(f)            # FDutils.py:21-33
    rng = np.random.default_rng(2601996)
    n = 2001
    f = np.linspace(0.0, 0.02, n)                  # includes f = 0 (extrapolated PSD, emri_pe.py:283)
    psd = get_sensitivity(f)
    env = np.exp(-((f - 0.004) / 0.002) ** 2) * 1e-18
    a = [env * (rng.normal(size=n) + 1j * rng.normal(size=n)) for _ in range(2)]
    b = [a[i] + 0.3 * env * (rng.normal(size=n) + 1j * rng.normal(size=n)) for i in range(2)]
    out = dict(f=f, psd=psd, a=np.asarray(a), b=np.asarray(b))
    out["ip_ab"] = inner_product(a, b, f_arr=f, PSD=psd)
    out["ip_ab_norm"] = inner_product(a, b, f_arr=f, PSD=psd, normalize=True)
    out["ip_ab_complex"] = inner_product(a, b, f_arr=f, PSD=psd, complex=True)
    out["ip_ab_sig1"] = inner_product(a, b, f_arr=f, PSD=psd, normalize="sig1")
    out["ip_a0b0"] = inner_product(a[0], b[0], f_arr=f, PSD=psd)
    out["ip_df"] = inner_product(a, b, df=f[1] - f[0], PSD=psd)
    out["snr_a"] = snr(a, f_arr=f, PSD=psd)
    # Likelihood with a table-lookup template model: params (idx, amp) -> amp * templates[idx]
    templates = np.asarray([[a[c] * (1.0 + 0.05 * k) + 0.02 * k * b[c] for c in range(2)] for k in range(5)])

    def model(idx, amp, **kw):
        return [amp * templates[int(idx)][0], amp * templates[int(idx)][1]]

    like = Likelihood(model, 2, f_arr=f, vectorized=False, transpose_params=False, subset=2)
    like.inject_signal(data_stream=[a[0], a[1]], noise_fn=[get_sensitivity, get_sensitivity], noise_kwargs=[{}, {}])
    params = np.array([[0, 1.0], [1, 1.0], [2, 0.9], [3, 1.1], [4, 1.0]])
    out["templates"] = templates
    out["like_params"] = params
    out["like_ll"] = like(params)
    out["like_noise_factor"] = np.asarray(like.noise_factor)
    out["like_injection"] = np.asarray(like.injection_channels)
    # TransformContainer as configured in emri_pe.py:161-206
    fill_dict = {"ndim_full": 14, "fill_values": np.array([0.0, 1.0, 2.45, np.pi / 3, np.pi / 3, np.pi / 3, np.pi / 3, 0.0]),
                 "fill_inds": np.array([2, 5, 6, 7, 8, 9, 10, 12])}
    tc = TransformContainer(parameter_transforms={(0, 1): lambda logM, logeta: [np.exp(logM), np.exp(logM) * np.exp(logeta)]},
                            fill_dict=fill_dict)
    p6 = np.array([[np.log(1e6), np.log(1e-5), 12.0, 0.35, 1.0, 2.0], [np.log(5e5), np.log(3e-5), 11.0, 0.2, 0.5, 0.1]])
    out["tc_in"], out["tc_out"] = p6, tc.both_transforms(p6)
    np.savez_compressed(os.path.join(HERE, "lisatools_golden.npz"), **out)
    print("lisatools golden: ip_ab", out["ip_ab"], "ll", out["like_ll"])


def spline_golden():
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(11)
    t = np.sort(rng.uniform(0, 3.0e7, 48))
    t[0] = 0.0
    y = np.vstack([np.cumsum(rng.normal(size=48)), 1e5 * np.sqrt(1 + t / 1e6), np.sin(t / 4e6), 1e-3 * (1 + t / 3e7) ** 2.5])
    cs = CubicSpline(t, y, axis=1)
    coeff = np.moveaxis(cs.c, 0, -1)[..., ::-1]          # [L-1][R][4] = (y, c1, c2, c3)
    S = np.genfromtxt(os.path.join(REF, "LISA_Alloc_Sh.txt"))
    fq = np.concatenate([[0.0, 1e-6, 5e-6], np.geomspace(1e-5, 1.0, 200), [1.2]])
    np.savez_compressed(os.path.join(HERE, "spline_golden.npz"), t=t, y=y, coeff=coeff, tq=np.linspace(-1e5, 3.05e7, 301),
                        yq=cs(np.linspace(-1e5, 3.05e7, 301)), psd_f=fq, psd_val=CubicSpline(S[:, 0], S[:, 1])(fq))


def k13_golden():
    import mpmath as mp
    from scipy.special import kv
    mp.mp.dps = 50
    X = np.concatenate([np.geomspace(1e-8, 1.0, 25), np.linspace(1.0, 40.0, 79), np.geomspace(40.0, 1e9, 30)])
    ref = np.array([complex(mp.besselk(mp.mpf(1) / 3, -1j * mp.mpf(float(x))) * mp.exp(-1j * mp.mpf(float(x)))) for x in X])
    sci = kv(1.0 / 3.0, -1j * X) * np.exp(-1j * X)
    np.savez_compressed(os.path.join(HERE, "k13_golden.npz"), X=X, Q_mpmath=ref, Q_scipy=sci)
    print("k13: scipy vs mpmath max rel", np.max(np.abs(sci - ref) / np.abs(ref)))


def waveform_golden():
    from helpers import CASES, make_item, oracle_waveform
    from oracle.oracle import Oracle
    from emri_frequencydomainwaveforms_b200.waveform import FastSchwarzschildEccentricFlux
    gen = FastSchwarzschildEccentricFlux(sum_kwargs=dict(pad_output=True, output_type="fd", odd_len=True))
    oq = Oracle("quad")
    for name in ("plunge", "ecc_many"):
        it = make_item(gen, name, dt=100.0)
        hp, hc, coeff, br, nbr = oracle_waveform(oq, it)
        nz = np.nonzero((hp != 0) | (hc != 0))[0]
        nz_all = nz
        nz = nz[:: max(1, len(nz) // 16000)]       # keep the fixture small: a strided subset of the support
        keys = ("t", "p", "e", "teuk_modes", "ylms", "Phi_phi", "Phi_r", "m_arr", "n_arr", "l_arr", "f_phi", "f_r")
        np.savez_compressed(os.path.join(HERE, f"waveform_{name}.npz"), **{k: it[k] for k in keys},
                            scale=it["scale"], M=it["M"], mu=it["mu"], T=it["T"], dt=it["dt"], N=it["N"],
                            nnz_total=len(nz_all), support_lo=nz_all.min(), support_hi=nz_all.max(), nz_index=nz, hp_nz=hp[nz], hc_nz=hc[nz], branches=br, nbr=nbr, n_eval=oq.last_n_eval)
        print(name, "N", it["N"], "L", len(it["t"]), "K", len(it["m_arr"]), "nnz", len(nz), "evals", oq.last_n_eval)


if __name__ == "__main__":
    lisatools_td_golden()
    lisatools_golden()
    spline_golden()
    k13_golden()
    waveform_golden()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))

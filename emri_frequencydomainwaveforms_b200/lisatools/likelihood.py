"""Likelihood on the FD path -- mirror of LISAanalysistools/lisatools/sampling/likelihood.py:13-334.

Two pieces:

* ``Likelihood`` keeps the reference constructor / ``inject_signal`` / ``get_ll`` / ``__call__``
  semantics (noise weighting d*sqrt(df/S), ``ll = -1/2 * 4 * sum |d~ - h~|^2``, ``subset`` chunking,
  parameter transforms, the ``template_model.get_ll`` plugin hook of likelihood.py:70-72,330-331).
  The reductions run in ``emrifd_loglike`` on the GPU.
* ``FDTemplateModel`` is the plugin: an object with ``get_ll(params, data, noise_factor, **kw)`` that
  the reference's own ``Likelihood`` (or ours) adopts verbatim with ``fill_data_noise=True``.  It
  evaluates the whole walker batch in ONE fused launch sequence (spline -> segmentation -> mode sum
  + |d~ - h~|^2), never materialising h(f) in HBM.
"""
import os

import numpy as np

from .. import _lib, engine


class FDTemplateModel:
    """Batched FD template + likelihood plugin around a ``GenerateEMRIWaveform``-shaped generator."""

    def __init__(self, waveform_generator, f_arr=None, device=None, producers="auto", chunk=64):
        self.gen = waveform_generator
        self.chunk = int(chunk)     # walkers per launch sequence; larger parameter batches are pipelined chunk by chunk
        self._side = None           # (handle, stream) of the producer pipeline
        self.base = waveform_generator.waveform_generator       # FastSchwarzschildEccentricFlux
        self._device = device
        self._data_ref = None
        self._data = None
        self.f_arr = f_arr
        self.last_h2d_bytes = 0
        if producers == "auto":
            from .. import _hostlib
            amp, ig = self.base.amplitude_generator, self.base.inspiral_generator
            producers = ("device" if hasattr(amp, "device_call") and getattr(ig, "use_native", False)
                         and _hostlib.load() is not None else "host")
        self.producers = producers          # "device": Ylm / mode selection / compaction on the GPU; "host": NumPy producers

    @property
    def handle(self):
        return _lib.get_handle(self._device)

    # ---- single template, list of [h+, hx] on f >= 0 (what Likelihood.get_ll calls per walker) ----
    def __call__(self, *params, **kwargs):
        kw = dict(kwargs)
        kw.pop("mask_positive", None)
        out = self.gen(*params, mask_positive=True, **kw)
        if not isinstance(out, (list, tuple)):
            # h+ - i hx of two COMPLEX frequency-domain channels cannot be split back into them
            raise ValueError("FDTemplateModel needs a generator built with return_list=True (frequency-domain h+, hx are complex).")
        return [out[0], out[1]]

    # ---- host producers for a batch of full parameter vectors [nb, 14] ----------------------------
    def prepare_batch(self, params, T=1.0, dt=10.0, eps=1e-5, mode_selection=None, **kwargs):
        items, ok = [], []
        for row in np.atleast_2d(params):
            M, mu, a, p0, e0, x0, dist, qS, phiS, qK, phiK, Phi_phi0, Phi_theta0, Phi_r0 = row[:14]
            theta, phi, c2, s2 = self.gen._transform(qS, phiS, qK, phiK)
            try:
                it = self.base.prepare(M, mu, p0, e0, theta, phi, dist=dist, Phi_phi0=Phi_phi0, Phi_r0=Phi_r0,
                                       T=T, dt=dt, eps=eps, mode_selection=mode_selection)
                it["cos2psi"], it["sin2psi"] = c2, s2
                items.append(it)
                ok.append(True)
            except ValueError:
                ok.append(False)      # out-of-domain draw: ll = NaN, Eryn maps it to -1e300 (red_blue.py:282-284)
        return items, np.asarray(ok)

    def set_data(self, data, noise_factor):
        """data: whitened injection channels [2][n]; noise_factor [2][n] (likelihood.py:213-220)."""
        import torch
        h = self.handle
        # The injected data live on the (process-wide, per-device) handle: another model or a Likelihood may have replaced
        # them since this model's last call.  Re-upload unless this model still owns the handle's data AND is handed the very
        # same objects (held by reference, so their ids cannot be recycled).
        same = self._data_ref is not None and self._data_ref[0] is data and self._data_ref[1] is noise_factor
        if same and getattr(h, "data_owner", None) is self:
            return
        to_np = lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
        d = np.ascontiguousarray(np.stack([to_np(c) for c in data]), dtype=np.complex128)
        w = np.ascontiguousarray(np.stack([to_np(c) for c in noise_factor]), dtype=np.float64)
        if d.shape[0] != 2 or w.shape != d.shape:
            raise ValueError("data and noise_factor must be [2, n] (plus, cross).")
        if np.isnan(w[0, 0]):
            # the reference skips the DC bin when the PSD is NaN there (start_ind = 1, likelihood.py:268): a zero weight and a
            # zero datum drop the bin from every sum just the same
            w, d = w.copy(), d.copy()
            w[:, 0] = 0.0
            d[:, 0] = 0.0
        dd = torch.from_numpy(d.view(np.float64)).to(h.torch_device)
        ww = torch.from_numpy(w).to(h.torch_device)
        h.check(h.lib.emrifd_set_data(h.h, dd.data_ptr(), ww.data_ptr(), d.shape[1]))
        self._data, self._data_ref, self.n_data = (dd, ww), (data, noise_factor), d.shape[1]
        h.data_owner = self

    def get_ll(self, params, data=None, noise_factor=None, T=1.0, dt=10.0, eps=1e-5, f_arr=None,
               include_minus_m=True, mode_selection=None, N=None, **kwargs):
        """params [nb, 14] (already transformed/filled) -> ll [nb] (numpy).  ``f_arr`` two-sided grid
        (emri_pe.py:339-349) or, if None, the implicit fftfreq grid of length ``N`` = 2*n_data-1."""
        import torch
        if not isinstance(params, np.ndarray):
            raise ValueError("params must be np.ndarray.")
        if data is not None:
            self.set_data(data, noise_factor)
        if self._data is None:
            raise ValueError("No data set: pass (data, noise_factor) or call set_data first.")
        h = self.handle
        if getattr(h, "data_owner", None) is not self:     # someone else used the handle since: point it back at this model's data
            h.check(h.lib.emrifd_set_data(h.h, self._data[0].data_ptr(), self._data[1].data_ptr(), self.n_data))
            h.data_owner = self
        f_arr = self.f_arr if f_arr is None else f_arr
        if f_arr is not None:
            f_host = f_arr.detach().cpu().numpy() if torch.is_tensor(f_arr) else np.asarray(f_arr)
            Ngrid, fpos = engine.grid_from_frequency(f_host)
            if getattr(self, "_fpos_ref", None) is not f_arr:     # cached by reference (an id() could be recycled)
                self._fpos_dev = torch.from_numpy(fpos).to(h.torch_device)
                self._fpos_ref = f_arr
            fpos_dev, val = self._fpos_dev, 0.0
        else:
            Ngrid = 2 * self.n_data - 1 if N is None else int(N)
            fpos_dev, val = None, 1.0 / (Ngrid * dt)
        if (Ngrid + 1) // 2 != self.n_data:
            raise ValueError("frequency grid and injected data have different lengths")
        producers = kwargs.pop("producers", self.producers)
        if producers == "device" and mode_selection is None:
            # Ylm, mode selection and compaction on the device; only the sparse tracks cross PCIe
            P = np.atleast_2d(params)
            from ..waveform import ssb_transform_batch
            ang = np.stack(ssb_transform_batch(P[:, 7], P[:, 8], P[:, 9], P[:, 10],
                                               detector_frame=getattr(self.gen, "frame", "detector") == "detector"), axis=1)
            if len(P) > self.chunk:
                return self._get_ll_pipelined(P, ang, h, Ngrid, val, fpos_dev, T, dt, eps, include_minus_m)
            db, ok = self.base.prepare_batch_device(P[:, 0], P[:, 1], P[:, 3], P[:, 4], ang[:, 0], ang[:, 1], dist=P[:, 6],
                                                    Phi_phi0=P[:, 11], Phi_r0=P[:, 13], T=T, dt=dt, eps=eps,
                                                    cos2psi=ang[:, 2], sin2psi=ang[:, 3], handle=h)
            ll = np.full(len(ok), np.nan)
            if db is not None:
                self.last_h2d_bytes = db.h2d_bytes
                out = engine.run_loglike(db, Ngrid, val, fpos_dev, include_minus_m=include_minus_m).cpu().numpy()
                # a walker the device refused (non-monotone knots, too many branches) comes back as NaN -- Eryn maps that to
                # -1e300 (Eryn/eryn/moves/red_blue.py:282-284); the other walkers of the batch are unaffected
                self.last_walker_status = np.zeros(len(ok), dtype=np.int32)
                self.last_walker_status[ok] = h.walker_status(db.pb.B)
                ll[ok] = out[:, 0]
                self.last_dh_hh = out[:, 1:]
            return ll
        items, ok = self.prepare_batch(params, T=T, dt=dt, eps=eps, mode_selection=mode_selection)
        ll = np.full(len(ok), np.nan)
        if items:
            pb = engine.PackedBatch(items)
            self.last_h2d_bytes = pb.h2d_bytes()
            out = engine.run_loglike_host(pb, h, Ngrid, val, fpos_dev, include_minus_m=include_minus_m)
            self.last_walker_status = np.zeros(len(ok), dtype=np.int32)
            self.last_walker_status[ok] = h.walker_status(pb.B)
            ll[ok] = out[:, 0]
            self.last_dh_hh = out[:, 1:]
        return ll


def _fdtm_pipelined(self, P, ang, h, Ngrid, val, fpos_dev, T, dt, eps, include_minus_m):
    """Large parameter batches (a whole tempered ensemble in one call): chunks of ``self.chunk`` walkers, the producers of
    chunk k+1 -- host trajectory ODE (threaded), H2D of the sparse tracks, device amplitudes / Ylm / mode selection /
    compaction, on a side stream through a handle of their own -- overlap the spline / segment / sum + likelihood of chunk k
    on the caller's stream.  Results are identical to chunk-by-chunk calls (walkers are independent)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    nb, ck = len(P), self.chunk
    NW = 2     # producer threads: two chunks are being prepared while a third is summed
    if self._side is None or self._side[0][0].device != h.device:
        self._side = []
        for _ in range(NW):
            st = torch.cuda.Stream(device=h.torch_device)
            self._side.append((_lib.Handle(h.device, st.cuda_stream), st))
    starts = list(range(0, nb, ck))
    cores = len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    nthr = max(1, min(cores // NW, 16))

    import time
    trace = [] if os.environ.get("EMRIFD_TRACE") else None

    def prep(i):
        t0 = time.perf_counter()
        h2, side = self._side[i % NW]
        sl = slice(starts[i], min(starts[i] + ck, nb))
        with torch.cuda.stream(side):
            db, ok = self.base.prepare_batch_device(P[sl, 0], P[sl, 1], P[sl, 3], P[sl, 4], ang[sl, 0], ang[sl, 1], dist=P[sl, 6],
                                                    Phi_phi0=P[sl, 11], Phi_r0=P[sl, 13], T=T, dt=dt, eps=eps,
                                                    cos2psi=ang[sl, 2], sin2psi=ang[sl, 3], handle=h2, nthreads=nthr)
            ev = torch.cuda.Event()
            ev.record(side)
        if trace is not None:
            trace.append(("prep", i, round(1e3 * (time.perf_counter() - t0), 2), getattr(db, "stage_ms", None)))
        return db, ok, ev, sl

    ll = np.full(nb, np.nan)
    self.last_walker_status = np.zeros(nb, dtype=np.int32)
    self.last_dh_hh = np.full((nb, 2), np.nan)
    self.last_h2d_bytes = 0
    main = torch.cuda.current_stream(h.torch_device)
    import sys
    # two Python threads hand the GIL back and forth a dozen times per chunk; with the default 5 ms switch interval every
    # hand-over can stall the other thread for milliseconds -- as long as a whole chunk takes on the device
    sw = sys.getswitchinterval()
    sys.setswitchinterval(2e-5)
    try:
        with ThreadPoolExecutor(max_workers=NW) as ex:
            futs = {i: ex.submit(prep, i) for i in range(min(NW, len(starts)))}
            pending = None                     # (like tensor, db, ok, slice) of the chunk whose kernels are in flight
            for i in range(len(starts) + 1):
                cur = None
                if i < len(starts):
                    db, ok, ev, sl = futs.pop(i).result()
                    if i + NW < len(starts):
                        futs[i + NW] = ex.submit(prep, i + NW)
                    if db is not None:
                        main.wait_event(ev)
                        db.handle = h          # the sum runs on the caller's handle (its stream, its injected data)
                        self.last_h2d_bytes += db.h2d_bytes
                        cur = (engine.run_loglike(db, Ngrid, val, fpos_dev, include_minus_m=include_minus_m), db, ok, sl,
                               h.walker_status_async(db.pb.B))
                if pending is not None:        # read the previous chunk back while this one runs
                    like_t, db_p, ok_p, sl_p, st_p = pending
                    out = like_t.cpu().numpy()
                    idx = np.arange(sl_p.start, sl_p.stop)[ok_p]
                    self.last_walker_status[idx] = st_p.cpu().numpy()
                    ll[idx] = out[:, 0]
                    self.last_dh_hh[idx] = out[:, 1:]
                pending = cur
    finally:
        sys.setswitchinterval(sw)
    if trace is not None:
        print("EMRIFD_TRACE", trace, flush=True)
    return ll


FDTemplateModel._get_ll_pipelined = _fdtm_pipelined


class Likelihood:
    """Mirror of lisatools.sampling.likelihood.Likelihood for frequency-domain templates."""

    def __init__(self, template_model, num_channels, dt=None, df=None, f_arr=None, parameter_transforms=None,
                 use_gpu=True, vectorized=False, separate_d_h=False, return_cupy=False, fill_data_noise=False,
                 transpose_params=False, subset=None, device=None):
        if dt is None and df is None and f_arr is None:
            raise ValueError("Must provide dt, df or f_arr.")
        if df is None and f_arr is None:
            raise ValueError("Time-domain likelihoods (dt= only) are outside the FD hot path.")
        if isinstance(template_model, list):
            raise ValueError("For single likelihood, template model cannot be a list.")
        self.template_model, self.num_channels = template_model, num_channels
        self.dt, self.df, self.f_arr = dt, df, f_arr
        self.parameter_transforms, self.subset = parameter_transforms, subset
        self.transpose_params, self.vectorized = transpose_params, vectorized
        self.fill_data_noise, self.separate_d_h = fill_data_noise, separate_d_h
        self.frequency_domain = True
        self._device = device
        if hasattr(template_model, "get_ll"):
            self._plugin_ll, self.like_here = template_model.get_ll, False
        else:
            self.fill_data_noise, self.like_here = False, True

    @property
    def handle(self):
        return _lib.get_handle(self._device)

    def _transform(self, params):
        if self.parameter_transforms is not None:
            key = list(self.parameter_transforms.keys())[0]
            params = self.parameter_transforms[key].both_transforms(params)
        return params

    def inject_signal(self, data_stream=None, params=None, waveform_kwargs={}, noise_fn=None, noise_kwargs={},
                      add_noise=False):
        import torch
        to_np = lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
        if params is not None:
            params = self._transform(params)
            inj = [to_np(c) for c in self.template_model(*params, **waveform_kwargs)]
        elif data_stream is not None:
            if not isinstance(data_stream, list):
                raise ValueError("If data_stream is provided, it must be as a list.")
            inj = [to_np(c) for c in data_stream]
        else:
            raise ValueError("Must provide data_stream or params kwargs to inject signal.")
        self.injection_length = len(inj[0])
        for c in inj:
            if len(c) != self.injection_length:
                raise ValueError("Length of all injection channels must match.")
        if len(inj) != self.num_channels:
            raise ValueError("Number of channels from template_model does not match number of channels declare by user.")
        if add_noise and self.df is not None and not getattr(self, "noise_has_been_added", False):
            raise NotImplementedError   # as the reference does (likelihood.py:183-184); with f_arr add_noise is ignored there too
        nf = noise_fn if isinstance(noise_fn, list) else [noise_fn] * self.num_channels
        if len(nf) == 1:
            nf = nf * self.num_channels
        if len(nf) != self.num_channels:
            raise ValueError("Number of noise functions does not match number of channels declared by user.")
        nk = noise_kwargs if isinstance(noise_kwargs, list) else [noise_kwargs] * self.num_channels
        if len(nk) == 1:
            nk = nk * self.num_channels
        freqs = np.arange(self.injection_length) * self.df if self.df is not None else to_np(self.f_arr)
        psd = [np.asarray(to_np(fn(freqs, **kw))) for fn, kw in zip(nf, nk)]
        diff_freqs = np.zeros_like(freqs)
        diff_freqs[1:] = np.diff(freqs)
        diff_freqs[0] = diff_freqs[1]
        self.base_injections = inj
        self.noise_factor = np.asarray([(diff_freqs / p) ** 0.5 for p in psd])
        whitened = np.asarray([c * w for c, w in zip(inj, self.noise_factor)])
        if hasattr(self, "injection_channels"):
            self.injection_channels = self.injection_channels + whitened
        else:
            self.injection_channels = whitened
        self.freqs, self.psd, self.data_length = freqs, psd, self.injection_length
        h = self.handle
        d_up, w_up = np.ascontiguousarray(self.injection_channels), np.ascontiguousarray(self.noise_factor)
        if np.isnan(w_up[0, 0]):      # PSD undefined at f = 0: the reference starts its sums at bin 1 (likelihood.py:268)
            d_up, w_up = d_up.copy(), w_up.copy()
            d_up[:, 0] = 0.0
            w_up[:, 0] = 0.0
        self._d_dev = torch.from_numpy(d_up.view(np.float64)).to(h.torch_device)
        self._w_dev = torch.from_numpy(w_up).to(h.torch_device)

    def get_ll(self, params, *args, **kwargs):
        import torch
        if not self.like_here:
            return self._plugin_ll(params, *args, **kwargs)
        if self.separate_d_h:
            raise NotImplementedError   # as the reference (likelihood.py:253-254)
        h = self.handle
        if self.vectorized:
            chans = self.template_model(*params, *args, **kwargs)
            tm = torch.stack([torch.as_tensor(c) for c in chans], dim=1) if isinstance(chans, (list, tuple)) else torch.as_tensor(chans)
        else:
            tm = torch.stack([torch.stack([torch.as_tensor(c).to(h.torch_device) for c in self.template_model(*p, *args, **kwargs)])
                              for p in params])
        tm = tm.to(device=h.torch_device, dtype=torch.complex128).contiguous()
        B = tm.shape[0]
        if tm.shape[1:] != (self.num_channels, self.data_length) or self.num_channels != 2:
            raise ValueError("templates must be [num_likes, 2, data_length]")
        out = torch.empty((B, 3), dtype=torch.float64, device=h.torch_device)
        h.check(h.lib.emrifd_set_data(h.h, self._d_dev.data_ptr(), self._w_dev.data_ptr(), self.data_length))
        h.data_owner = self
        h.check(h.lib.emrifd_loglike(h.h, tm.data_ptr(), B, out.data_ptr()))
        return np.atleast_1d(out[:, 0].cpu().numpy())

    def __call__(self, params, *args, **kwargs):
        if not isinstance(params, np.ndarray):
            raise ValueError("params must be np.ndarray.")
        params = self._transform(params)
        if self.transpose_params:
            params = params.T
            subset_axis = 1
        else:
            subset_axis = 0
        num_likes = params.shape[subset_axis]
        inds_likes = np.arange(num_likes)
        if self.subset is not None:
            if not isinstance(self.subset, int):
                raise ValueError("Subset must be int.")
            inds_subset = np.split(inds_likes, np.arange(self.subset, num_likes, self.subset))
        else:
            inds_subset = [inds_likes]
        out_ll = []
        for inds in inds_subset:
            args_in = (params[inds],) if subset_axis == 0 else (params[:, inds],)
            args_in += args
            if self.fill_data_noise:
                args_in += (self.injection_channels, self.noise_factor)
            out_ll.append(self.get_ll(*args_in, **kwargs))
        return np.concatenate(out_ll, axis=0)

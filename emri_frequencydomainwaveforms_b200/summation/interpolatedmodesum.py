"""``CubicSplineInterpolant`` -- mirror of ``few.summation.interpolatedmodesum.CubicSplineInterpolant``
as used by the reference (Tutorial_FD_construction_single_mode.ipynb:176,201,380): not-a-knot cubic
splines of ``ninterps`` rows sharing one knot vector, built and evaluated on the GPU
(``emrifd_spline_build`` / ``emrifd_spline_eval``)."""
import numpy as np

from .. import _lib


class CubicSplineInterpolant:
    """``CubicSplineInterpolant(t, y_all)``; ``spline(t_new)`` returns ``[ninterps, len(t_new)]``.

    Inputs may be numpy arrays or torch tensors; results are torch CUDA tensors (float64).
    ``interp_array`` exposes the coefficients with few's logical shape ``(4, length, ninterps)``
    = [y | c1 | c2 | c3][knot][interp] (a permuted view of the knot-major quad layout the kernels use).
    """

    def __init__(self, t, y_all, use_gpu=True, device=None, **kwargs):
        import torch
        self.handle = _lib.get_handle(device)
        dev = self.handle.torch_device
        t = torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t, dtype=torch.float64).to(dev).contiguous()
        y = torch.as_tensor(np.asarray(y_all) if not torch.is_tensor(y_all) else y_all, dtype=torch.float64).to(dev)
        if y.ndim == 1:
            y = y[None, :]
        y = y.contiguous()
        if t.ndim != 1 or y.shape[1] != t.shape[0]:
            raise ValueError("t must be 1D [length] and y_all [ninterps, length].")
        self.t, self.ninterps, self.length = t, int(y.shape[0]), int(y.shape[1])
        self.coeff = torch.empty((self.length, self.ninterps, 4), dtype=torch.float64, device=dev)
        h = self.handle
        h.check(h.lib.emrifd_spline_build(h.h, t.data_ptr(), y.data_ptr(), self.length, self.ninterps,
                                          self.length, 1, self.coeff.data_ptr()))
        h.status()

    @property
    def interp_array(self):
        return self.coeff.permute(2, 0, 1)

    def __call__(self, tnew, deriv_order=0):
        import torch
        if deriv_order != 0:
            raise ValueError("Only deriv_order=0 is available on this path.")
        dev = self.handle.torch_device
        tn = torch.as_tensor(np.asarray(tnew) if not torch.is_tensor(tnew) else tnew, dtype=torch.float64).to(dev)
        shape = tn.shape
        tn = tn.reshape(-1).contiguous()
        out = torch.empty((self.ninterps, tn.numel()), dtype=torch.float64, device=dev)
        h = self.handle
        h.check(h.lib.emrifd_spline_eval(h.h, self.t.data_ptr(), self.coeff.data_ptr(), self.length,
                                         self.ninterps, tn.data_ptr(), tn.numel(), out.data_ptr()))
        return out.reshape((self.ninterps,) + tuple(shape))

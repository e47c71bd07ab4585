"""Parameter fill + mass transform used by the reference's likelihood call.

The reference wires an eryn ``TransformContainer`` (Eryn/eryn/utils/transform.py:181-226) with
``fill_inds = [2, 5, 6, 7, 8, 9, 10, 12]`` and ``(0, 1): (logM, log eta) -> (M, M eta)``
(emri_pe.py:161-206).  Eryn's own container keeps working with our Likelihood; this helper is the
dependency-free equivalent for the sharded batch driver.
"""
import numpy as np


def fill_and_transform(params, fill_inds, fill_values, ndim_full=14):
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    out = np.zeros(params.shape[:-1] + (ndim_full,))
    test_inds = np.delete(np.arange(ndim_full), fill_inds)
    out[..., test_inds] = params
    out[..., fill_inds] = fill_values
    M = np.exp(out[..., 0])
    out[..., 1] = M * np.exp(out[..., 1])
    out[..., 0] = M
    return out


class TransformContainer:
    """Minimal stand-in with eryn's ``both_transforms`` entry point for the EMRI configuration."""

    def __init__(self, fill_dict):
        self.fill_dict = fill_dict

    def both_transforms(self, params, **kwargs):
        return fill_and_transform(params, self.fill_dict["fill_inds"], self.fill_dict["fill_values"],
                                  self.fill_dict.get("ndim_full", 14))

"""ctypes binding of libemrifd.so (include/emrifd.h).  There is NO CPU fallback: if the CUDA
library is missing or no CUDA device is present, every product entry point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EMRIFD_LIB", os.path.join(_HERE, "csrc", "libemrifd.so"))

MAX_BRANCHES = 4
INCLUDE_MINUS_M = 1
MASK_POSITIVE = 2

BRANCH_DTYPE = np.dtype([
    ("mode", np.int32), ("dir", np.int32), ("ja", np.int32), ("jb", np.int32),
    ("closed_end", np.int32), ("pad", np.int32), ("start", np.int64), ("end", np.int64),
    ("xa", np.float64), ("xb", np.float64), ("Fa", np.float64), ("Fb", np.float64)])

WALKER_DTYPE = np.dtype([
    ("L", np.int32), ("K", np.int32), ("knot_off", np.int64), ("teuk_off", np.int64),
    ("mode_off", np.int64), ("coeff_off", np.int64), ("out_off", np.int64),
    ("scale", np.float64), ("cos2psi", np.float64), ("sin2psi", np.float64)])

ERRORS = {-1: "EMRIFD_ERR_INVALID", -2: "EMRIFD_ERR_TOO_FEW_KNOTS", -3: "EMRIFD_ERR_KNOT_ORDER",
          -4: "EMRIFD_ERR_BRANCHES", -5: "EMRIFD_ERR_NOMEM", -6: "EMRIFD_ERR_CUDA",
          -7: "EMRIFD_ERR_TOO_MANY_KNOTS", -8: "EMRIFD_ERR_NO_DATA"}

# every symbol include/emrifd.h declares (tests check that the built library exports all of them)
SYMBOLS = [
    "emrifd_version", "emrifd_sizeof_branch", "emrifd_sizeof_walker", "emrifd_create", "emrifd_destroy",
    "emrifd_set_stream", "emrifd_synchronize", "emrifd_last_error", "emrifd_spline_build",
    "emrifd_spline_eval", "emrifd_batch_spline", "emrifd_batch_segment", "emrifd_batch_sum",
    "emrifd_fd_waveform_batch", "emrifd_batch_status", "emrifd_set_data", "emrifd_inner_product",
    "emrifd_loglike", "emrifd_loglike_batch_host", "emrifd_bench_fp64_fma", "emrifd_launch_count",
    "emrifd_sum_kernel_time", "emrifd_mode_select", "emrifd_ylm_batch", "emrifd_mode_compact_count",
    "emrifd_mode_compact_gather", "emrifd_tile_bins", "emrifd_batch_sum_cyclic",
    "emrifd_synth_amplitude", "emrifd_walker_status", "emrifd_walker_status_dev", "emrifd_set_k13_mode",
    "emrifd_window_taps", "emrifd_band_energy", "emrifd_band_convolve", "emrifd_sum_kernel_times", "emrifd_set_overlap"]

_lib = None


class EmrifdError(RuntimeError):
    pass


def load():
    """Load libemrifd.so; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmrifdError(
            f"{LIB_PATH} not found: build it with `python -m emri_frequencydomainwaveforms_b200.csrc.build` "
            "(or __graft_entry__.build()).  This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, dbl, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    lib.emrifd_create.argtypes = [i32, vp, C.POINTER(vp)]
    lib.emrifd_destroy.argtypes = [vp]
    lib.emrifd_set_stream.argtypes = [vp, vp]
    lib.emrifd_synchronize.argtypes = [vp]
    lib.emrifd_last_error.argtypes = [vp]
    lib.emrifd_last_error.restype = C.c_char_p
    lib.emrifd_spline_build.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp]
    lib.emrifd_spline_eval.argtypes = [vp, vp, vp, i64, i64, vp, i64, vp]
    lib.emrifd_batch_spline.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.emrifd_batch_segment.argtypes = [vp, vp, i64, vp, vp, vp, vp, i64, dbl, vp, vp, vp]
    lib.emrifd_batch_sum.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, dbl, vp, i32, i64, i64, vp, vp, vp]
    lib.emrifd_batch_sum_cyclic.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, dbl, vp, i32, i64, i64, vp, vp, vp]
    lib.emrifd_fd_waveform_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, dbl, vp, i32,
                                             vp, vp, vp, vp, vp]
    lib.emrifd_batch_status.argtypes = [vp]
    lib.emrifd_walker_status.argtypes = [vp, i64, vp]
    lib.emrifd_walker_status_dev.argtypes = [vp, i64, vp]
    lib.emrifd_set_k13_mode.argtypes = [vp, i32]
    lib.emrifd_set_overlap.argtypes = [vp, i32]
    lib.emrifd_set_data.argtypes = [vp, vp, vp, i64]
    lib.emrifd_inner_product.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp]
    lib.emrifd_loglike.argtypes = [vp, vp, i64, vp]
    lib.emrifd_loglike_batch_host.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, dbl, vp, i32, vp, vp, vp]
    lib.emrifd_mode_select.argtypes = [vp, vp, i64, i64, vp, vp, vp, i64, i64, dbl, vp]
    lib.emrifd_ylm_batch.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp, i64, i32, vp]
    lib.emrifd_synth_amplitude.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, i64, i32, i32, vp]
    lib.emrifd_mode_compact_count.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.emrifd_mode_compact_gather.argtypes = [vp, vp, i64, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.emrifd_window_taps.argtypes = [vp, vp, i64, i32, vp]
    lib.emrifd_band_energy.argtypes = [vp, vp, i64, i32, vp]
    lib.emrifd_band_convolve.argtypes = [vp, vp, i32, vp, i64, i64, i64, i64, vp]
    lib.emrifd_bench_fp64_fma.argtypes = [vp, i32, C.POINTER(dbl)]
    lib.emrifd_launch_count.argtypes = [vp]
    lib.emrifd_launch_count.restype = i64
    lib.emrifd_sum_kernel_time.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(i64)]
    lib.emrifd_sum_kernel_times.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64)]
    assert lib.emrifd_sizeof_branch() == BRANCH_DTYPE.itemsize
    assert lib.emrifd_sizeof_walker() == WALKER_DTYPE.itemsize
    _lib = lib
    return lib


class Handle:
    """One emrifd handle per (process, device, stream)."""

    def __init__(self, device=None, stream=None):
        import torch
        if not torch.cuda.is_available():
            raise EmrifdError("No CUDA device: the FD mode-sum path runs only on the GPU (no CPU fallback).")
        self.lib = load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        if stream is None:
            stream = torch.cuda.current_stream(self.torch_device).cuda_stream
        self.stream = int(stream)
        self.data_owner = None      # the object whose injected data emrifd_set_data last pointed this handle at
        hp = C.c_void_p()
        rc = self.lib.emrifd_create(self.device, C.c_void_p(self.stream), C.byref(hp))
        if rc != 0:
            raise EmrifdError(f"emrifd_create failed: {ERRORS.get(rc, rc)}")
        self.h = hp

    def check(self, rc):
        if rc == 0:
            return
        msg = self.lib.emrifd_last_error(self.h).decode()
        name = ERRORS.get(rc, str(rc))
        # argument / trajectory problems mirror FEW's ValueError behaviour (check_mode_by_mode.py:218-219,328-330)
        if rc in (-1, -2, -3, -4, -7, -8):
            raise ValueError(f"{name}: {msg}")
        raise EmrifdError(f"{name}: {msg}")

    def status(self):
        """Raise ValueError if ANY walker of the last batch failed (the single-waveform behaviour of FEW's sanity checks)."""
        self.check(self.lib.emrifd_batch_status(self.h))

    def walker_status(self, B):
        """Per-walker status codes of the last batch (0 = ok).  Failed walkers have h = 0 and ll = NaN; the others are
        unaffected -- the per-walker contract of the reference's samplers (Eryn/eryn/moves/red_blue.py:282-284)."""
        out = np.zeros(int(B), dtype=np.int32)
        self.check(self.lib.emrifd_walker_status(self.h, int(B), out.ctypes.data))
        self.lib.emrifd_batch_status(self.h)   # clear the sticky batch word: the failures have been reported per walker
        return out

    def walker_status_async(self, B):
        """The same status words as a device int32 tensor, copied on the handle's stream without a sync."""
        import torch
        out = torch.empty(int(B), dtype=torch.int32, device=self.torch_device)
        self.check(self.lib.emrifd_walker_status_dev(self.h, int(B), out.data_ptr()))
        return out

    def set_k13_mode(self, mode):
        """"exact" (default, <= 3e-15) or "few" (FastEMRIWaveforms' 14-term / 9-term SPAFunc with its seam at |X| = 7)."""
        codes = {"exact": 0, "few": 1}
        if mode not in codes:
            raise ValueError("k13_mode must be 'exact' or 'few'")
        self.check(self.lib.emrifd_set_k13_mode(self.h, codes[mode]))

    def synchronize(self):
        self.check(self.lib.emrifd_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.emrifd_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.emrifd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def get_handle(device=None):
    """Process-wide lazily created handle per device (created post-fork, cf. emri_pe.py:545)."""
    import torch
    if device is None:
        if not torch.cuda.is_available():
            raise EmrifdError("No CUDA device: the FD mode-sum path runs only on the GPU (no CPU fallback).")
        device = torch.cuda.current_device()
    key = (os.getpid(), int(device))
    if key not in _handles:
        _handles[key] = Handle(device)
    h = _handles[key]
    # follow torch's CURRENT stream (a `with torch.cuda.stream(s):` block, a non-default-stream worker): tensors are allocated
    # and uploaded on it by the Python layer, so the kernels must be ordered on it too
    cur = int(torch.cuda.current_stream(h.torch_device).cuda_stream)
    if cur != h.stream:
        h.check(h.lib.emrifd_set_stream(h.h, C.c_void_p(cur)))
        h.stream = cur
    return h


def ptr(x):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()

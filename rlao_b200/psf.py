"""Host wrapper of aoenv_psf_peak: peak of the science PSF (OOPAO/Telescope.py:260-360) over the central window."""
import math

import numpy as np
import torch

from . import _lib

_TW = {}


def _twiddles(N, device):
    key = (N, str(device))
    if key not in _TW:
        m = np.arange(2 * N, dtype=np.float64)
        ang = -math.pi * m / N
        _TW[key] = torch.as_tensor(np.stack([np.cos(ang), np.sin(ang)], axis=1), dtype=torch.float32, device=device).contiguous()
    return _TW[key]


def psf_peak(tel, opd_a, opd_b, zeroPaddingFactor=4, window=32, return_window=False):
    """max over the central `window` x `window` pixels of tel.computePSF(zeroPaddingFactor)'s PSF, for
    OPD_no_pupil = opd_a (+ opd_b), per frame.  Returns [F] (and the window [F, w, w] if asked)."""
    N, os_, img_size, pad, img_res = tel.psf_geometry(zeroPaddingFactor)
    if img_res % 2 != 0:
        raise NotImplementedError("odd PSF sizes use a different phasor (Telescope.py:330)")
    F, R = opd_a.shape[0], tel.resolution
    dev = tel.device
    tw = _twiddles(N, dev)
    scratch = torch.empty((F, os_ * window, R, 2), dtype=torch.float32, device=dev)
    out = torch.empty((F,), dtype=torch.float32, device=dev)
    win = torch.empty((F, window, window), dtype=torch.float32, device=dev) if return_window else None
    amp = (tel._pupil_f * torch.as_tensor(tel.pupilReflectivity, dtype=torch.float32, device=dev) * tel.src._amp_dev).contiguous()
    _lib.check(_lib.load().aoenv_psf_peak(_lib.ptr(opd_a), _lib.ptr(opd_b), _lib.ptr(tel._pupil_f), _lib.ptr(amp),
                                          _lib.ptr(tw), F, R, N, os_, window, 2 * math.pi / tel.src.wavelength,
                                          _lib.ptr(scratch), _lib.ptr(win), _lib.ptr(out), _lib.stream_ptr(dev)), "psf_peak")
    return (out, win) if return_window else out

"""Host wrapper of aoenv_psf_peak: peak of the science PSF (OOPAO/Telescope.py:260-360) over the central window."""
import math

import numpy as np
import torch

from . import _lib, gemm

_OPS = {}
_WORK = {}


def _operators(R, N, os_, window, device):
    """Twiddle operands of the pruned DFT (include/aoenv.h), exact integer phase reduction in float64."""
    key = (R, N, os_, window, str(device))
    if key not in _OPS:
        Wu, pad = os_ * window, (N - R) // 2
        s0 = os_ * ((N // os_) // 2 - window // 2)
        d = s0 + np.arange(Wu, dtype=np.int64) - N // 2
        m = ((pad + np.arange(R, dtype=np.int64))[None, :] * (2 * d + 1)[:, None]) % (2 * N)       # [Wu, R]
        ang = -math.pi * m.astype(np.float64) / N
        wr, wi = np.cos(ang), np.sin(ang)
        ldk = (2 * R + 15) // 16 * 16
        w1 = np.zeros((2 * Wu, ldk), dtype=np.float32)
        w1[:Wu, :R], w1[:Wu, R:2 * R] = wr, -wi
        w1[Wu:, :R], w1[Wu:, R:2 * R] = wi, wr
        op = gemm.Operator(torch.as_tensor(w1, device=device).contiguous(), parts=2)
        g2 = torch.as_tensor(np.stack([wr.T, wi.T], axis=-1), dtype=torch.float32, device=device).contiguous()   # [R, Wu, 2]
        _OPS[key] = (op, g2, ldk)
    return _OPS[key]


def _workspace(F, R, ldk, Wu, device):
    key = (F, R, ldk, Wu, str(device))
    if key not in _WORK:
        _WORK[key] = (torch.zeros((2, F * R, ldk), dtype=torch.bfloat16, device=device),
                      torch.empty((F * R, 2 * Wu), dtype=torch.float32, device=device))
    return _WORK[key]


def psf_peak(tel, opd_a, opd_b, zeroPaddingFactor=4, window=32, return_window=False):
    """max over the central `window` x `window` pixels of tel.computePSF(zeroPaddingFactor)'s PSF, for
    OPD_no_pupil = opd_a (+ opd_b), per frame.  Returns [F] (and the window [F, w, w] if asked)."""
    N, os_, img_size, pad, img_res = tel.psf_geometry(zeroPaddingFactor)
    if img_res % 2 != 0:
        raise NotImplementedError("odd PSF sizes use a different phasor (Telescope.py:330)")
    F, R = opd_a.shape[0], tel.resolution
    dev = tel.device
    op, g2, ldk = _operators(R, N, os_, window, dev)
    planes, scratch = _workspace(F, R, ldk, os_ * window, dev)
    out = torch.empty((F,), dtype=torch.float32, device=dev)
    win = torch.empty((F, window, window), dtype=torch.float32, device=dev) if return_window else None
    amp = (tel._pupil_f * torch.as_tensor(tel.pupilReflectivity, dtype=torch.float32, device=dev) * tel.src._amp_dev).contiguous()
    _lib.check(_lib.load().aoenv_psf_peak(_lib.ptr(opd_a), _lib.ptr(opd_b), _lib.ptr(tel._pupil_f), _lib.ptr(amp),
                                          _lib.ptr(op.planes()), _lib.ptr(g2), F, R, N, os_, window,
                                          2 * math.pi / tel.src.wavelength, _lib.ptr(planes), ldk, _lib.ptr(scratch),
                                          _lib.ptr(win), _lib.ptr(out), _lib.stream_ptr(dev)), "psf_peak")
    return (out, win) if return_window else out


def psf_image(tel, opd_a, opd_b, zeroPaddingFactor=4, img_resolution=None, chunk_bytes=1 << 30):
    """The PSF image of tel.computePSF(zeroPaddingFactor[, img_resolution]) (OOPAO/Telescope.py:260-360: padded field,
    half-pixel phasor, centred FFT / N, |.|^2, oversampling binned away, central crop), for OPD_no_pupil = opd_a (+ opd_b),
    per frame: [F, img, img] and its maxima [F].  Both transforms run as tensor-core GEMMs (aoenv_psf_image)."""
    N, os_, img_size, pad, img_res = tel.psf_geometry(zeroPaddingFactor, img_resolution)
    if img_res % 2 != 0:
        raise NotImplementedError("odd PSF sizes use a different phasor (Telescope.py:330)")
    win = img_res
    F, R, dev = opd_a.shape[0], tel.resolution, tel.device
    op, _, ldk = _operators(R, N, os_, win, dev)
    Wu = os_ * win
    amp = (tel._pupil_f * torch.as_tensor(tel.pupilReflectivity, dtype=torch.float32, device=dev) * tel.src._amp_dev).contiguous()
    out = torch.empty((F, win, win), dtype=torch.float32, device=dev)
    peak = torch.empty((F,), dtype=torch.float32, device=dev)
    per_frame = 2 * R * ldk * 2 + R * 2 * Wu * 4 + 2 * Wu * ldk * 2 + Wu * 2 * Wu * 4
    chunk = max(1, min(F, 32767, chunk_bytes // per_frame))
    field = torch.zeros((2, chunk * R, ldk), dtype=torch.bfloat16, device=dev)
    work_t = torch.empty((chunk * R, 2 * Wu), dtype=torch.float32, device=dev)
    planes_u = torch.zeros((2, chunk * Wu, ldk), dtype=torch.bfloat16, device=dev)
    work_f = torch.empty((chunk * Wu, 2 * Wu), dtype=torch.float32, device=dev)
    lib = _lib.load()
    for s0 in range(0, F, chunk):
        m = min(chunk, F - s0)
        a = opd_a[s0:s0 + m].contiguous()
        b = None if opd_b is None else opd_b[s0:s0 + m].contiguous()
        fp = field if m == chunk else field[:, :m * R].contiguous()
        pu = planes_u if m == chunk else planes_u[:, :m * Wu].contiguous()
        _lib.check(lib.aoenv_psf_image(_lib.ptr(a), _lib.ptr(b), _lib.ptr(tel._pupil_f), _lib.ptr(amp), _lib.ptr(op.planes()),
                                       m, R, N, os_, win, 2 * math.pi / tel.src.wavelength, _lib.ptr(fp), ldk, _lib.ptr(work_t),
                                       _lib.ptr(pu), _lib.ptr(work_f), _lib.ptr(out[s0:s0 + m]), _lib.ptr(peak[s0:s0 + m]),
                                       _lib.stream_ptr(dev)), "psf_image")
    return out, peak

"""Zernike — mirror of OOPAO/Zernike.py:20-62 (Noll-ordered modes on the telescope pupil, piston skipped,
each mode mean-removed and normalised to unit std over the pupil).  Init-time, float64 on the device."""
import math

import numpy as np
import torch


def _noll_to_nm(j):
    """Noll (1976) single index -> (n, m); m > 0 for even j (cosine), m < 0 for odd j (sine)."""
    n = int((-1.0 + math.sqrt(8 * (j - 1) + 1)) / 2.0)
    p = j - (n * (n + 1)) // 2
    k = n % 2
    m = int((p + k) / 2.0) * 2 - k
    if m != 0 and j % 2 != 0:
        m = -m
    return n, m


def _radial(n, m, r):
    out = torch.zeros_like(r)
    for s in range((n - m) // 2 + 1):
        c = ((-1) ** s) * math.factorial(n - s) / (
            math.factorial(s) * math.factorial((n + m) // 2 - s) * math.factorial((n - m) // 2 - s))
        out = out + c * r ** (n - 2 * s)
    return out


class Zernike:
    def __init__(self, telObject, J=1):
        self.resolution = telObject.resolution
        self.D = telObject.D
        self.centralObstruction = telObject.centralObstruction
        self.nModes = J

    def zernike_tel(self, tel, j):
        dev = tel.device
        res = tel.resolution
        X, Y = np.where(tel.pupil > 0)
        c = (res + res % 2 - 1) / 2
        x = torch.as_tensor((X - c) / res * tel.D, dtype=torch.float64, device=dev)
        y = torch.as_tensor((Y - c) / res * tel.D, dtype=torch.float64, device=dev)
        r = torch.sqrt(x ** 2 + y ** 2)
        r = r / r.max()
        th = torch.atan2(y, x)
        out = torch.zeros((tel.pixelArea, j), dtype=torch.float64, device=dev)
        for i in range(1, j + 1):
            n, m = _noll_to_nm(i + 1)
            if m == 0:
                Z = math.sqrt(n + 1) * _radial(n, 0, r)
            elif m > 0:
                Z = math.sqrt(2 * (n + 1)) * _radial(n, m, r) * torch.cos(m * th)
            else:
                Z = math.sqrt(2 * (n + 1)) * _radial(n, -m, r) * torch.sin(-m * th)
            Z = Z - Z.mean()
            out[:, i - 1] = Z / Z.std(unbiased=False)
        full = torch.zeros((res * res, j), dtype=torch.float64, device=dev)
        full[tel._pupil_idx] = out
        return out, full.reshape(res, res, j)

    def computeZernike(self, telObject2):
        self.modes, self.modesFullRes = self.zernike_tel(telObject2, self.nModes)

"""Restatement of the two aotools functions OOPAO/Zernike.py:47-56 calls.

Noll, "Zernike polynomials and atmospheric turbulence", JOSA 66 (1976): index j -> (n, m)
with even j -> cosine term (returned as m > 0) and odd j -> sine term (m < 0), and the
radial polynomial R_n^m(r) = sum_s (-1)^s (n-s)! / (s! ((n+m)/2-s)! ((n-m)/2-s)!) r^(n-2s).
"""
import math
import numpy as np


class _Zernike:
    @staticmethod
    def zernIndex(j):
        n = int((-1.0 + math.sqrt(8 * (j - 1) + 1)) / 2.0)
        p = j - (n * (n + 1)) // 2
        k = n % 2
        m = int((p + k) / 2.0) * 2 - k
        if m != 0 and j % 2 != 0:
            m = -m
        return [n, m]

    @staticmethod
    def zernikeRadialFunc(n, m, r):
        out = np.zeros(r.shape)
        for s in range((n - m) // 2 + 1):
            c = ((-1) ** s) * math.factorial(n - s) / (
                math.factorial(s) * math.factorial((n + m) // 2 - s) * math.factorial((n - m) // 2 - s))
            out += c * r ** (n - 2 * s)
        return out


zernike = _Zernike()

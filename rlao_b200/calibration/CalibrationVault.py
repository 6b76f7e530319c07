"""CalibrationVault — mirror of OOPAO/calibration/CalibrationVault.py:15-57: SVD of the interaction matrix
and its (optionally truncated) pseudo-inverse.  float64 torch tensors on the matrix's device."""
import torch

from ..tools import linalg


class CalibrationVault:
    def __init__(self, D, nTrunc=0, display=False, print_details=False, invert=True):
        D = torch.as_tensor(D, dtype=torch.float64)
        if not invert:
            self.D = D
            return
        U, s, V = linalg.svd(D)
        self.s = s
        self.S = torch.diag(s)
        self.eigenValues = s
        self.D = U @ self.S @ V
        self.U = U.T
        self.V = V
        self.iS = torch.diag(1 / s)
        self.M = V.T @ self.iS @ self.U
        self._set_trunc(nTrunc)

    def _set_trunc(self, nTrunc):
        self._nTrunc = nTrunc
        n = len(self.s) - nTrunc
        self.iStrunc = torch.diag(1 / self.eigenValues[:n])
        self.Vtrunc, self.Utrunc = self.V[:n, :], self.U[:n, :]
        self.VtruncT, self.UtruncT = self.Vtrunc.T, self.Utrunc.T
        self.Mtrunc = self.VtruncT @ self.iStrunc @ self.Utrunc
        self.Dtrunc = self.UtruncT @ torch.diag(self.eigenValues[:n]) @ self.Vtrunc
        self.cond = float(self.eigenValues[0] / self.eigenValues[-nTrunc - 1])

    @property
    def nTrunc(self):
        return self._nTrunc

    @nTrunc.setter
    def nTrunc(self, val):
        self._set_trunc(val)

"""Init-time dense linear algebra (pseudo-inverses and SVDs of the calibration, OOPAOEnvRazor.py:261,337 and
calibration/CalibrationVault.py:15-57), float64 with numpy's conventions (pinv cut-off 1e-15 of the largest singular
value).  Small problems are solved by LAPACK on the host: a batched-Jacobi cuSOLVER SVD of a 2000 x 70 matrix costs
thousands of tiny launches for nothing; large ones (the 45 244 x 1 353 influence matrix of the 40 x 40 system and up) stay
on the device."""
import numpy as np
import torch

HOST_WORK_LIMIT = 5e9          # rows * cols * min(rows, cols) below which the host is used


def _small(A):
    m, n = A.shape[-2], A.shape[-1]
    return m * n * min(m, n) <= HOST_WORK_LIMIT


def pinv(A, rcond=1e-15):
    A = torch.as_tensor(A, dtype=torch.float64)
    if _small(A):
        return torch.as_tensor(np.linalg.pinv(A.cpu().numpy(), rcond=rcond), dtype=torch.float64, device=A.device)
    return torch.linalg.pinv(A, rtol=rcond)


def svd(A):
    """U, s, Vh with full_matrices=False."""
    A = torch.as_tensor(A, dtype=torch.float64)
    if _small(A):
        U, s, Vh = np.linalg.svd(A.cpu().numpy(), full_matrices=False)
        t = lambda x: torch.as_tensor(x, dtype=torch.float64, device=A.device)
        return t(U), t(s), t(Vh)
    return torch.linalg.svd(A, full_matrices=False)

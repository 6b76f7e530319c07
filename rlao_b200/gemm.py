"""Dense contractions of the step: D[x][w] = alpha * sum_k X[x][k] W[w][k] (include/aoenv.h).

Two hand-written CUDA back ends:
  "tc"   tcgen05 tensor cores on split-bf16 operands (aoenv_gemm_tn_tc) — the production path; relative error of a
         product ~2^-17 of sum |x||w| with parts=2 (DM surface, reconstruction, exploration noise, PSF rows) and ~2^-24
         (FP32 grade) with parts=3 (add_row, screen synthesis);
  "simt" FP32 FMA kernel (aoenv_gemm_tn) — kept as the on-device cross-check of the tensor-core path.
Select with rlao_b200.gemm.BACKEND or the AOENV_GEMM environment variable."""
import os

import torch

from . import _lib

BACKEND = os.environ.get("AOENV_GEMM", "tc")


class Operator:
    """Static operand W [N, Kp] (float32, zero padded to Kp % 16 == 0) plus its lazily built bf16 planes."""

    def __init__(self, W, parts=2):
        assert W.dtype == torch.float32 and W.stride(1) == 1 and W.stride(0) % 16 == 0
        self.W, self.parts, self._planes = W, parts, None

    def invalidate(self):
        self._planes = None

    def planes(self):
        if self._planes is None:
            N, Kp = self.W.shape[0], self.W.stride(0)
            self._planes = torch.empty((self.parts, N, Kp), dtype=torch.bfloat16, device=self.W.device)
            _lib.check(_lib.load().aoenv_split_bf16(_lib.ptr(self.W), Kp, N, Kp, self.parts, _lib.ptr(self._planes), Kp,
                                                    _lib.stream_ptr(self.W.device)), "split_bf16(W)")
        return self._planes


_workspaces = {}


def _x_planes(X, parts):
    # one workspace per shape AND stream: calls on one stream are ordered, two streams (or two environments driven from
    # two streams) must not share the planes between the split and the GEMM that reads them
    key = (X.device, X.shape[0], X.stride(0), parts, torch.cuda.current_stream(X.device).cuda_stream if X.is_cuda else 0)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.empty((parts, X.shape[0], X.stride(0)), dtype=torch.bfloat16, device=X.device)
        _workspaces[key] = ws
    return ws


SKINNY_MAX_ROWS = 8      # AOENV_SKINNY_MAX_ROWS (include/aoenv.h): this few rows of X go to the exact-FP32 warp-per-column kernel


def uses_tensor_cores(rows=None, backend=None):
    """Whether a product with `rows` rows of X (None: any large one) runs on the tcgen05 kernel."""
    return (backend or BACKEND) == "tc" and (rows is None or rows > SKINNY_MAX_ROWS)


def gemm_tn(X, op, D, M, N, alpha=1.0, backend=None, x_planes=None):
    """X [M, Kp] float32, op: Operator over W [N, Kp]; D [M, >=N] float32 (row stride D.stride(0)).
    x_planes: X already in split-bf16 form [op.parts, M, Kp] (written by the kernel that produced X)."""
    backend = backend or ("tc" if uses_tensor_cores(M) else "simt")
    lib, st = _lib.load(), _lib.stream_ptr(X.device)
    Kp = op.W.stride(0)
    assert X.stride(0) == Kp, "operands must share the padded K"
    if backend == "simt":
        _lib.check(lib.aoenv_gemm_tn(_lib.ptr(X), Kp, _lib.ptr(op.W), Kp, _lib.ptr(D), D.stride(0), M, N, Kp, alpha, st), "gemm_tn")
        return
    if x_planes is not None:
        xs = x_planes
    else:
        xs = _x_planes(X, op.parts)
        _lib.check(lib.aoenv_split_bf16(_lib.ptr(X), Kp, M, Kp, op.parts, _lib.ptr(xs), Kp, st), "split_bf16(X)")
    _lib.check(lib.aoenv_gemm_tn_tc(_lib.ptr(xs), _lib.ptr(op.planes()), Kp, op.parts, _lib.ptr(D), D.stride(0), M, N, Kp,
                                    alpha, st), "gemm_tn_tc")

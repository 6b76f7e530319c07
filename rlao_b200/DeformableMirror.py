"""DeformableMirror — mirror of OOPAO/DeformableMirror.py (zonal Gaussian influence functions or user modes),
batched over environments.  `dm.coefs = c` has the reference's side effect: dm.OPD = modes @ c, here a
[n_frames, nValidAct] x [R*R, nValidAct]^T contraction on the GPU (include/aoenv.h: aoenv_gemm_tn).

HBM layout: modes[R*R][Kp] float32, Kp = nValidAct rounded up to 16 with zero padding, K-contiguous, so the
DM surface of every environment is one "TN" GEMM  OPD[B][R*R] = coefs[B][Kp] . modes[R*R][Kp]^T (tcgen05).

`surface_backend`: "gemm" = that dense contraction (any modes: custom, rotated, anamorphic); "separable" = for the
default Cartesian grid of axis-aligned Gaussians the product factorises into two small banded products
(aoenv_dm_surface_separable), ~30x fewer flops and exact to float32 rounding; "auto" (default) picks separable
when the geometry allows it.
"""
import math
import os
import sys

import numpy as np
import torch

from . import _lib, gemm
from .MisRegistration import MisRegistration


class DMSurfaceRef:
    """A DM surface that may not have been written to memory yet: (mirror, slot).  The fused step kernel
    (aoenv_shwfs_fused) evaluates the surface of the separable geometry in shared memory straight from the commands;
    anybody else who needs the [n_envs, R, R] tensor calls `.tensor()`, which runs the surface kernel on demand."""

    def __init__(self, dm, slot):
        self.dm, self.slot = dm, slot

    def tensor(self):
        return self.dm._slot_surface(self.slot)

    @property
    def materialised(self):
        return bool(self.dm._valid[self.slot])

    @property
    def coefs(self):
        return self.dm._coefs_of[self.slot]

    @property
    def rows(self):
        """T = C gx [n_envs, nAct + 18, R] of this surface (aoenv_dm_rows), the DM operand of aoenv_shwfs_fused."""
        return self.dm._rows_of(self.slot)

    @property
    def shape(self):
        return self.dm._opd[self.slot].shape


class DeformableMirror:
    def __init__(self, telescope, nSubap, mechCoupling=0.35, coordinates=None, pitch=None, modes=None, misReg=None,
                 M4_param=None, nJobs=30, nThreads=20, print_dm_properties=True, floating_precision=64, altitude=None):
        if M4_param is not None and M4_param.get("isM4", False):
            raise NotImplementedError("M4 influence functions are out of scope")
        if altitude is not None:
            raise NotImplementedError("DM in altitude is out of scope")
        if mechCoupling <= 0:
            raise ValueError("The value of mechanical coupling should be positive.")
        self.telescope = telescope
        self.device = telescope.device
        self.n_envs = telescope.n_envs
        self.altitude = None
        self.isM4 = False
        self.resolution = telescope.resolution
        self.mechCoupling = mechCoupling
        self.tag = "deformableMirror"
        self.D = telescope.D
        self.floating_precision = floating_precision
        self.pitch = self.D / nSubap if pitch is None else pitch          # DeformableMirror.py:266-270
        self.misReg = MisRegistration() if misReg is None else misReg
        R = self.resolution

        if coordinates is None:                                           # :286-305 Cartesian (Fried) geometry
            self.nAct = nSubap + 1
            self.nActAlongDiameter = self.nAct - 1
            x = np.linspace(-self.D / 2, self.D / 2, self.nAct)
            X, Y = np.meshgrid(x, x)
            self.xIF0, self.yIF0 = X.reshape(-1), Y.reshape(-1)
            r = np.sqrt(self.xIF0 ** 2 + self.yIF0 ** 2)
            inner = r > (telescope.centralObstruction * self.D / 2 - 0.5 * self.pitch)
            outer = r <= (self.D / 2 + 0.7533 * self.pitch)
            self.validAct = inner * outer
            self.nValidAct = int(self.validAct.sum())
        else:                                                             # :309-321 explicit coordinates
            coordinates = np.asarray(coordinates)
            if coordinates.ndim != 2 or coordinates.shape[1] != 2:
                raise AttributeError("Wrong size for the DM coordinates, the (x,y) coordinates should be input as a 2D array of dimension [nAct,2]")
            self.xIF0, self.yIF0 = coordinates[:, 0], coordinates[:, 1]
            self.nAct = len(self.xIF0)
            self.nActAlongDiameter = self.D / self.pitch
            self.validAct = np.arange(self.nAct).astype(int)
            self.nValidAct = self.nAct

        x0, y0 = self.xIF0[self.validAct], self.yIF0[self.validAct]
        mr = self.misReg
        x3, y3 = self.anamorphosis(x0, y0, mr.anamorphosisAngle * np.pi / 180, mr.tangentialScaling, mr.radialScaling)
        x4, y4 = self.rotateDM(x3, y3, mr.rotationAngle * np.pi / 180)
        self.xIF, self.yIF = x4 - mr.shiftX, y4 - mr.shiftY
        self.nIF = len(self.xIF)
        self.coordinates = np.stack([self.xIF, self.yIF], axis=1)

        if modes is None:
            modes64 = self._gaussian_modes()
        else:
            modes64 = torch.as_tensor(np.asarray(modes), dtype=torch.float64, device=self.device)
            self.nValidAct = modes64.shape[1]
        self._set_modes(modes64)
        self.surface_backend = "auto"
        self._sep = self._separable_tables() if (modes is None and coordinates is None) else None
        self._opd = torch.zeros((2, self.n_envs, R, R), dtype=torch.float32, device=self.device)   # ping-pong
        self._slot = 0
        # lazy surfaces (opt-in, AOENV_WFS=fused): with the separable geometry the fused WFS kernel builds the surface in
        # shared memory from T = C gx; _opd[slot] is then brought up to date only when somebody reads it
        # lazy_surface: the fast path of env.step (_set_coefs_batch) computes only T = C gx per command; the consumer either
        # evaluates the surface in place (aoenv_shwfs_frame_dm, aoenv_shwfs_fused) or asks for the tensor, which runs the
        # surface kernel on demand.  Set by the environment when its WFS can evaluate the surface itself.
        self.lazy_surface = os.environ.get("AOENV_WFS", "kernels") == "fused"
        self._coefs_of = [None, None]
        self._valid = [True, True]
        self._rows = None                 # [2][n_envs, nAct + 18, R]: column half of the separable surface, per slot
        self._rows_valid = [False, False]
        self._multi = None            # [k, R, R] surfaces of a [nValidAct, k] command matrix (calibration)
        self._coefs = torch.zeros((self.n_envs, self._Kp), dtype=torch.float32, device=self.device)
        self._coefs_matrix = None
        self.current_coefs = None

    # ---- geometry helpers (DeformableMirror.py:480-492) -------------------------------------------------
    @staticmethod
    def rotateDM(x, y, angle):
        return x * np.cos(angle) - y * np.sin(angle), y * np.cos(angle) + x * np.sin(angle)

    @staticmethod
    def anamorphosis(x, y, angle, mRad, mNorm):
        mRad, mNorm = mRad + 1, mNorm + 1
        xo = x * (mRad * np.cos(angle) ** 2 + mNorm * np.sin(angle) ** 2) + y * (mNorm * np.sin(2 * angle) / 2 - mRad * np.sin(2 * angle) / 2)
        yo = y * (mRad * np.sin(angle) ** 2 + mNorm * np.cos(angle) ** 2) + x * (mNorm * np.sin(2 * angle) / 2 - mRad * np.sin(2 * angle) / 2)
        return xo, yo

    def _gaussian_modes(self):
        """DeformableMirror.py:494-514: anisotropic Gaussian IFs on the grid linspace(0,1,R)*R, float64 on device."""
        R, mr, dev = self.resolution, self.misReg, self.device
        u0x = torch.as_tensor(R / 2 + self.xIF * R / self.D, dtype=torch.float64, device=dev)
        u0y = torch.as_tensor(R / 2 + self.yIF * R / self.D, dtype=torch.float64, device=dev)
        base = (R / self.nActAlongDiameter) / math.sqrt(2 * math.log(1.0 / self.mechCoupling))
        cx, cy = (1 + mr.radialScaling) * base, (1 + mr.tangentialScaling) * base
        th = mr.anamorphosisAngle * math.pi / 180
        a = math.cos(th) ** 2 / (2 * cx ** 2) + math.sin(th) ** 2 / (2 * cy ** 2)
        b = -math.sin(2 * th) / (4 * cx ** 2) + math.sin(2 * th) / (4 * cy ** 2)
        c = math.sin(th) ** 2 / (2 * cx ** 2) + math.cos(th) ** 2 / (2 * cy ** 2)
        g = torch.linspace(0, 1, R, dtype=torch.float64, device=dev) * R
        out = torch.empty((R * R, self.nValidAct), dtype=torch.float64, device=dev)
        chunk = max(1, (1 << 27) // (R * R))
        for s in range(0, self.nValidAct, chunk):
            dx = g[None, None, :] - u0x[s:s + chunk, None, None]          # X varies along columns
            dy = g[None, :, None] - u0y[s:s + chunk, None, None]          # Y varies along rows
            G = torch.exp(-(a * dx ** 2 + 2 * b * dx * dy + c * dy ** 2))
            out[:, s:s + chunk] = G.reshape(G.shape[0], R * R).T
        return out

    def _separable_tables(self):
        """Per-column / per-row Gaussian factors and their bands when modes[:, k] = gy_i(y) gx_j(x) on a regular grid."""
        mr, R, dev = self.misReg, self.resolution, self.device
        if mr.rotationAngle != 0 or mr.anamorphosisAngle != 0:
            return None
        n = self.nAct
        x = np.linspace(-self.D / 2, self.D / 2, n)
        X, Y = np.meshgrid(x, x)
        x3, y3 = self.anamorphosis(X.reshape(-1), Y.reshape(-1), 0.0, mr.tangentialScaling, mr.radialScaling)
        xg = (x3 - mr.shiftX).reshape(n, n)
        yg = (y3 - mr.shiftY).reshape(n, n)
        if np.abs(xg - xg[:1, :]).max() > 1e-12 or np.abs(yg - yg[:, :1]).max() > 1e-12:
            return None
        u0x = R / 2 + xg[0, :] * R / self.D                  # per actuator column
        u0y = R / 2 + yg[:, 0] * R / self.D                  # per actuator row
        base = (R / self.nActAlongDiameter) / math.sqrt(2 * math.log(1.0 / self.mechCoupling))
        a = 1.0 / (2 * ((1 + mr.radialScaling) * base) ** 2)
        c = 1.0 / (2 * ((1 + mr.tangentialScaling) * base) ** 2)
        g = np.linspace(0, 1, R) * R
        ex = a * (g[None, :] - u0x[:, None]) ** 2             # [nAct, R]
        ey = c * (g[None, :] - u0y[:, None]) ** 2
        CUT = 21.0                                            # exp(-21) = 2^-30.3: below float32 resolution of the sum

        def bands(e):
            keep = e <= CUT
            lo = np.where(keep.any(axis=0), keep.argmax(axis=0), 0)
            hi = np.where(keep.any(axis=0), e.shape[0] - 1 - keep[::-1].argmax(axis=0), -1)
            return np.stack([lo, hi], axis=1).astype(np.int32)
        rows, cols = np.nonzero(np.reshape(self.validAct, (n, n)))
        t = lambda arr, dt: torch.as_tensor(np.ascontiguousarray(arr), dtype=dt, device=dev)
        bx, by, gxv, gyv = bands(ex), bands(ey), np.exp(-ex), np.exp(-ey)
        row_start = np.searchsorted(rows, np.arange(n + 1)).astype(np.int32)      # valid actuators are listed row-major
        out = dict(gx=t(gxv, torch.float32), gy=t(gyv, torch.float32), band_x=t(bx, torch.int32), band_y=t(by, torch.int32),
                   act_pos=t((rows * n + cols).astype(np.int32), torch.int32), act_row_start=t(row_start, torch.int32),
                   W=0, wx=None, j0x=None, wyp=None, i0y=None, i0y_host=None, nAct=n, by_host=by.astype(np.int64),
                   gy_host=gyv)
        # fixed-width band tables for the unrolled kernel: per pixel column, and per PAIR of pixel rows
        if R % 2 == 0:
            pair_lo = np.minimum(by[0::2, 0], by[1::2, 0])
            pair_hi = np.maximum(by[0::2, 1], by[1::2, 1])
            need = int(max((bx[:, 1] - bx[:, 0] + 1).max(), (pair_hi - pair_lo + 1).max()))
            W = 12 if need <= 12 else (16 if need <= 16 else 0)
            if W:
                wx = np.zeros((R, W))
                for x in range(R):
                    lo, hi = bx[x]
                    wx[x, :hi - lo + 1] = gxv[lo:hi + 1, x]
                wyp = np.zeros((R // 2, 2, W))
                for k in range(R // 2):
                    for h in range(2):
                        lo, hi = by[2 * k + h]
                        wyp[k, h, lo - pair_lo[k]:hi - pair_lo[k] + 1] = gyv[lo:hi + 1, 2 * k + h]
                out.update(W=W, wx=t(wx, torch.float32), j0x=t(bx[:, 0], torch.int32), wyp=t(wyp, torch.float32),
                           i0y=t(pair_lo.astype(np.int32), torch.int32), i0y_host=pair_lo.astype(np.int64))
        return out

    def _set_modes(self, modes64):
        self._modes64 = modes64                      # kept until the calibration is done (free_float64())
        self._Kp = (self.nValidAct + 15) // 16 * 16
        self._modes = torch.zeros((modes64.shape[0], self._Kp), dtype=torch.float32, device=self.device)
        self._modes[:, :self.nValidAct] = modes64.to(torch.float32)
        self._modes_op = gemm.Operator(self._modes, parts=2)

    def free_float64(self):
        self._modes64 = None

    @property
    def modes(self):
        return self._modes64 if self._modes64 is not None else self._modes[:, :self.nValidAct]

    @modes.setter
    def modes(self, val):
        m = torch.as_tensor(np.asarray(val) if not torch.is_tensor(val) else val, dtype=torch.float64, device=self.device)
        self.nValidAct = m.shape[1]
        self._set_modes(m)
        self._sep = None                 # user-supplied modes: dense contraction only

    # ---- surfaces ----------------------------------------------------------------------------------------
    def _surface(self, coefs_padded, out, backend=None):
        """out[f] = modes @ coefs[f] for every frame f (OPD = modes @ coefs, DeformableMirror.py:534-570)."""
        F = coefs_padded.shape[0]
        P = self.resolution ** 2
        choice = self.surface_backend
        if choice not in ("auto", "gemm", "separable"):
            raise ValueError("surface_backend must be 'auto', 'gemm' or 'separable'")
        if choice == "separable" and self._sep is None:
            raise ValueError("this DM geometry is not separable (rotation / anamorphosis / custom modes): use 'gemm'")
        if self._sep is not None and choice in ("auto", "separable"):
            t = self._sep
            _lib.check(_lib.load().aoenv_dm_surface_separable(
                _lib.ptr(coefs_padded), coefs_padded.stride(0), _lib.ptr(t["act_pos"]), self.nValidAct, self.nAct,
                _lib.ptr(t["gx"]), _lib.ptr(t["gy"]), _lib.ptr(t["band_x"]), _lib.ptr(t["band_y"]), _lib.ptr(t["wx"]),
                _lib.ptr(t["j0x"]), _lib.ptr(t["wyp"]), _lib.ptr(t["i0y"]), t["W"], F, self.resolution,
                _lib.ptr(out), _lib.stream_ptr(self.device)), "dm_surface_separable")
            return
        gemm.gemm_tn(coefs_padded, self._modes_op, out.reshape(F, P), F, P, backend=backend)

    def _set_coefs_batch(self, coefs_padded):
        """Fast path of env.step: per-environment commands [B, Kp]; writes the *next* surface slot and makes it
        current.  The previous slot keeps the surface the WFS has just seen (tel.OPD stays reproducible)."""
        self._coefs = coefs_padded
        self._multi = None
        self._coefs_matrix = None
        self._slot ^= 1
        self._coefs_of[self._slot] = coefs_padded
        self._rows_valid[self._slot] = False
        if self.lazy_surface and self.fused_tables() is not None:
            self._valid[self._slot] = False       # evaluated inside aoenv_shwfs_fused, or on demand
            self._rows_of(self._slot)             # T = C gx now (39 KB per environment instead of the 230 KB surface)
        else:
            self._surface(coefs_padded, self._opd[self._slot])
            self._valid[self._slot] = True

    def fused_tables(self):
        """Banded tables of the separable geometry in the form aoenv_shwfs_fused takes, or None."""
        t = self._sep
        if t is None or t["W"] not in (12, 16) or self.surface_backend not in ("auto", "separable"):
            return None
        return t

    ROWS_PAD = 18                     # zero rows after the last actuator row: windows may run past it without clamping

    def _rows_of(self, slot):
        t = self.fused_tables()
        if self._rows is None:
            self._rows = torch.zeros((2, self.n_envs, self.nAct + self.ROWS_PAD, self.resolution), dtype=torch.float32,
                                     device=self.device)
        if not self._rows_valid[slot]:
            c = self._coefs_of[slot]
            if c is None:
                c = torch.zeros((self.n_envs, self._Kp), dtype=torch.float32, device=self.device)
            _lib.check(_lib.load().aoenv_dm_rows(_lib.ptr(c), c.stride(0), _lib.ptr(t["act_pos"]), self.nValidAct, self.nAct,
                                                 self.nAct + self.ROWS_PAD, _lib.ptr(t["wx"]), _lib.ptr(t["j0x"]), t["W"],
                                                 self.n_envs, self.resolution, _lib.ptr(self._rows[slot]),
                                                 _lib.stream_ptr(self.device)), "dm_rows")
            self._rows_valid[slot] = True
        return self._rows[slot]

    def _slot_surface(self, slot):
        if not self._valid[slot]:
            self._surface(self._coefs_of[slot], self._opd[slot])
            self._valid[slot] = True
        return self._opd[slot]

    def surface_ref(self):
        """The current surface as a (possibly not yet materialised) reference."""
        return DMSurfaceRef(self, self._slot)

    def _previous_surface(self):
        return self._slot_surface(self._slot ^ 1)

    @property
    def coefs(self):
        if self._coefs_matrix is not None:
            return self._coefs_matrix
        c = self._coefs[:, :self.nValidAct]
        return c[0] if self.n_envs == 1 else c

    @coefs.setter
    def coefs(self, val):
        """Reference semantics (DeformableMirror.py:534-570): scalar 0 resets; a vector of length nValidAct commands
        the mirror (here: every environment); a [nValidAct, k] matrix produces k surfaces at once (calibration).
        Extension: a [n_envs, nValidAct] tensor commands each environment separately."""
        nA, B = self.nValidAct, self.n_envs
        if np.isscalar(val):
            if val != 0:
                print("Error: wrong value for the coefficients")
                return
            c = torch.zeros((B, self._Kp), dtype=torch.float32, device=self.device)
            self._set_coefs_batch(c)
            return
        t = torch.as_tensor(val, dtype=torch.float32, device=self.device)
        if t.ndim == 1 and t.shape[0] == nA:
            c = torch.zeros((B, self._Kp), dtype=torch.float32, device=self.device)
            c[:, :nA] = t
            self._set_coefs_batch(c)
        elif t.ndim == 2 and t.shape[0] == nA and not (t.shape[0] == B and t.shape[1] == nA and B != nA):
            k = t.shape[1]
            c = torch.zeros((k, self._Kp), dtype=torch.float32, device=self.device)
            c[:, :nA] = t.T
            self._multi = torch.empty((k, self.resolution, self.resolution), dtype=torch.float32, device=self.device)
            self._surface(c, self._multi, backend="simt")     # calibration pushes: exact FP32 (init only)
            self._coefs_matrix = t
        elif t.ndim == 2 and t.shape == (B, nA):
            c = torch.zeros((B, self._Kp), dtype=torch.float32, device=self.device)
            c[:, :nA] = t
            self._set_coefs_batch(c)
        else:
            print("Error: wrong value for the coefficients")
            sys.exit(0)                                                     # DeformableMirror.py:567-569
        self.current_coefs = self.coefs

    @property
    def OPD(self):
        if self._multi is not None:
            return self._multi
        o = self._slot_surface(self._slot)
        return o[0] if self.n_envs == 1 else o

    def dm_propagation(self, telescope, OPD_in=None, i_source=None):
        """DeformableMirror.py:452-478: OPD_no_pupil leaving the mirror (a new tensor)."""
        dm_opd = self._multi if self._multi is not None else self._slot_surface(self._slot)
        if telescope.isPaired:
            base = telescope._materialise() if OPD_in is None else OPD_in
            if dm_opd.shape[0] != base.shape[0]:
                base = base[:1].expand(dm_opd.shape[0], -1, -1)
            return base + dm_opd
        return dm_opd.clone()

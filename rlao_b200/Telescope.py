"""Telescope — mirror of OOPAO/Telescope.py for the closed-loop SH path, batched over environments.

Holds the pupil, the OPD <-> source-phase bridge and the `*`, `+`, `-` propagation operators
(Telescope.py:457-564, 705-737).  `tel.OPD` / `tel.OPD_no_pupil` are CUDA tensors [n_envs, R, R]
(squeezed to [R, R] when n_envs == 1).  Inside `env.step` the sum atmosphere + DM is never written to
memory: the WFS kernel reads both terms, and `tel.OPD` is materialised only when somebody asks for it.
"""
import math

import numpy as np
import torch

from . import _lib


class Telescope:
    def __init__(self, resolution, diameter, samplingTime=0.001, centralObstruction=0, fov=0, pupil=None,
                 pupilReflectivity=1, display_optical_path=False, n_envs=1, device=None):
        self.device = _lib.require_cuda(device)
        self.n_envs = int(n_envs)
        self.isInitialized = False
        self.resolution = int(resolution)
        self.D = diameter
        self.pixelSize = self.D / self.resolution
        self.centralObstruction = centralObstruction
        self.fov = fov
        self.fov_rad = fov / 206265
        self.samplingTime = samplingTime
        self.isPetalFree = False
        self.user_defined_pupil = pupil
        self._reflectivity_in = pupilReflectivity
        self.set_pupil()
        self.src = None
        self.tag = "telescope"
        self.isPaired = False
        self.spatialFilter = None
        self.display_optical_path = display_optical_path
        self.optical_path = None
        R = self.resolution
        # Telescope.py:152-153: initial OPD = pupil, OPD_no_pupil = 1
        self._lazy = None
        self._opd_np = torch.ones((self.n_envs, R, R), device=self.device, dtype=torch.float32)
        self.isInitialized = True

    # ---- pupil ----------------------------------------------------------------------------------------
    def set_pupil(self):
        """Telescope.py:164-180."""
        R = self.resolution
        if self.user_defined_pupil is None:
            D = R + 1
            x = np.linspace(-R / 2, R / 2, R)
            xx, yy = np.meshgrid(x, x)
            circle = xx ** 2 + yy ** 2
            pup = (circle < (D / 2) ** 2) & (circle >= (self.centralObstruction * D / 2) ** 2)
        else:
            pup = np.asarray(self.user_defined_pupil).copy()
        self.pupil = pup

    @property
    def pupil(self):
        return self._pupil

    @pupil.setter
    def pupil(self, val):
        """Telescope.py:390-398: stored as int; reflectivity reset to uniform."""
        self._pupil = np.asarray(val).astype(int)
        self.pixelArea = int(np.sum(self._pupil))
        self.pupilLogical = np.where(self._pupil.reshape(-1) > 0)
        self.pupilReflectivity = self._pupil.astype(float) * (self._reflectivity_in if np.isscalar(self._reflectivity_in) else 1.0)
        self._pupil_f = torch.as_tensor(self._pupil, dtype=torch.float32, device=self.device).contiguous()
        self._pupil_idx = torch.as_tensor(self.pupilLogical[0], dtype=torch.long, device=self.device)

    # ---- OPD state ------------------------------------------------------------------------------------
    def _set_lazy(self, opd_a, opd_b):
        """Fast path of env.step: OPD_no_pupil = opd_a (+ opd_b), not yet written anywhere."""
        self._lazy = (opd_a, opd_b)
        self._opd_np = None

    def _materialise(self):
        if self._opd_np is None:
            a, b = self._lazy
            if b is not None and not torch.is_tensor(b):
                b = b.tensor()                                # DMSurfaceRef: the surface kernel runs now
            self._opd_np = a.clone() if b is None else a + b
            self._lazy = None
        return self._opd_np

    def _terms(self, resolve=True):
        """(opd_a, opd_b) such that OPD_no_pupil = opd_a + opd_b, without forcing a materialisation of the sum.
        resolve=False may return opd_b as a DeformableMirror.DMSurfaceRef (a surface not yet written to memory)."""
        if self._opd_np is not None:
            return self._opd_np, None
        a, b = self._lazy
        if resolve and b is not None and not torch.is_tensor(b):
            b = b.tensor()
        return a, b

    def _squeeze(self, t):
        return t[0] if (self.n_envs == 1 and t.shape[0] == 1) else t

    @property
    def OPD_no_pupil(self):
        return self._squeeze(self._materialise())

    @OPD_no_pupil.setter
    def OPD_no_pupil(self, val):
        self._opd_np = self._as_batch(val)
        self._lazy = None

    @property
    def OPD(self):
        return self._squeeze(self._materialise() * self._pupil_f)

    @OPD.setter
    def OPD(self, val):
        # the reference keeps OPD and OPD_no_pupil as two arrays; here OPD is always OPD_no_pupil * pupil, so
        # assigning OPD (already masked by the caller) assigns the un-masked term as well
        self._opd_np = self._as_batch(val)
        self._lazy = None

    @property
    def mean_removed_OPD(self):
        opd = self._materialise() * self._pupil_f
        mean = opd.reshape(opd.shape[0], -1)[:, self._pupil_idx].mean(dim=1)
        return self._squeeze((opd - mean[:, None, None]) * self._pupil_f)

    def _as_batch(self, val):
        t = torch.as_tensor(val, dtype=torch.float32, device=self.device)
        if t.ndim == 2:
            t = t.unsqueeze(0).expand(self.n_envs, -1, -1)
        return t.contiguous().clone()

    def resetOPD(self):
        """Telescope.py:566-580."""
        R = self.resolution
        self._opd_np = torch.zeros((self.n_envs, R, R), device=self.device, dtype=torch.float32)
        self._lazy = None

    def _on_new_source(self):
        """Second half of src*tel (Source.py:136-159)."""
        src = self.src
        if self._opd_np is not None and self._opd_np.shape[0] != self.n_envs:
            self.resetOPD()                                   # Source.py:138-139 (3-D OPD left by a calibration)
        src.fluxMap = self.pupilReflectivity * src.nPhoton * self.samplingTime * (self.D / self.resolution) ** 2
        src._amp_dev = torch.as_tensor(np.sqrt(src.fluxMap), dtype=torch.float32, device=self.device).contiguous()
        src._flux_version = getattr(src, "_flux_version", 0) + 1

    # ---- PSF ------------------------------------------------------------------------------------------
    def psf_geometry(self, zeroPaddingFactor, img_resolution=None):
        """Sizes PropagateField derives (Telescope.py:296-326): (N, oversampling, img_size, pad, img_resolution)."""
        R = self.resolution
        img_res = int(zeroPaddingFactor * R) if img_resolution is None else int(img_resolution)
        if img_res > zeroPaddingFactor * R:
            raise ValueError("Error: image has too many pixels for this pupil sampling. Try using a pupil mask with more pixels")
        os_ = 1
        if zeroPaddingFactor * os_ < 2:
            os_ = int(math.ceil(2.0 / zeroPaddingFactor))
        if os_ % 2 != img_res % 2:
            os_ += 1
        img_size = int(math.ceil(img_res * os_))
        N = int(np.fix(zeroPaddingFactor * os_ * R))
        pad = int(math.ceil((N - R) / 2))
        return R + 2 * pad, os_, img_size, pad, img_res

    def computePSF(self, zeroPaddingFactor=2, detector=None, img_resolution=None):
        """Telescope.py:260-293 -> PropagateField :296-360: PSF image(s) in `tel.PSF` ([n_envs, S, S], squeezed for one
        env), S = img_resolution (default zeroPaddingFactor * resolution; a detector brings its own sampling and size).
        Runs on the library's kernels (aoenv_psf_image: both transforms are tensor-core GEMMs)."""
        from .psf import psf_image
        if detector is not None:
            zeroPaddingFactor = detector.psf_sampling
            img_resolution = detector.resolution
        if self.src is None:
            raise AttributeError("The telescope was not coupled to any source object! Make sure to couple it with an src object using src*tel")
        a, b = self._terms()
        psf, peak = psf_image(self, a.contiguous(), None if b is None else b.contiguous(), zeroPaddingFactor, img_resolution)
        img = psf.shape[-1]
        conv = (180 / math.pi) * 3600
        half = (self.src.wavelength / self.D) * (img / 2 / zeroPaddingFactor)
        self.xPSF_rad = self.yPSF_rad = [-half, half]
        self.xPSF_arcsec = self.yPSF_arcsec = [-conv * half, conv * half]
        self.PSF = self._squeeze(psf)
        self.PSF_norma = self._squeeze(psf / peak[:, None, None])

    def _materialise_psf(self):
        p = self.PSF
        return p if p.ndim == 3 else p.unsqueeze(0)

    # ---- operators --------------------------------------------------------------------------------------
    def __mul__(self, obj):
        """Telescope.py:457-564, dispatch on obj.tag."""
        tag = getattr(obj, "tag", None)
        if tag in ("shackHartmann", "pyramid"):               # Telescope.py:476-485
            obj.telescope = self
            obj.wfs_measure()
        elif tag == "deformableMirror":
            dm_opd = obj.dm_propagation(self)            # DeformableMirror.py:452-478
            self._opd_np = dm_opd
            self._lazy = None
        elif tag == "detector":
            # Telescope.py:487-500: the science camera looks at the PSF with its own sampling and size; every tel*cam adds
            # one AO frame to the exposure, which is read out when the integration time is reached
            self.computePSF(detector=obj)
            obj.fov_arcsec = self.xPSF_arcsec[1] - self.xPSF_arcsec[0]
            obj.fov_rad = self.xPSF_rad[1] - self.xPSF_rad[0]
            if obj.integrationTime is not None and obj.integrationTime < self.samplingTime:
                raise ValueError("The Detector integration time is smaller than the AO loop sampling Time. ")
            obj._integrated_time += self.samplingTime
            obj.integrate(self._materialise_psf())
            if obj.frame is not None:
                self.PSF = obj.frame
        elif tag == "telescope":
            pass                                   # `atm*src*tel*cam` (OOPAOEnv.py:207,315): atm*src already returned the telescope
        else:
            raise AttributeError(f"Telescope cannot be propagated to an object with tag {tag!r}")
        return self

    def __add__(self, obj):
        """tel+atm (Telescope.py:705-727)."""
        if getattr(obj, "tag", None) == "atmosphere":
            obj * self
        else:
            raise AttributeError("only an Atmosphere can be combined with a Telescope")

    def __sub__(self, obj):
        """tel-atm (Telescope.py:729-737)."""
        if getattr(obj, "tag", None) == "atmosphere":
            self.isPaired = False
            self.resetOPD()
        else:
            raise AttributeError("only an Atmosphere can be separated from a Telescope")

"""CPU oracle for the Pyramid WFS (OOPAO/Pyramid.py) — TEST INFRASTRUCTURE ONLY, groundwork for SURVEY.md section 8 f-3.

float64 numpy restatement of the path the drl4ao papyrus environment uses (MAIN_CODE/OOPAOEnv/OOPAOEnv.py:239-249):
4-sided pyramid, PSF centred on 4 pixels (`psfCentering=True`), circular tip-tilt modulation, `slopesMaps` /
`slopesMaps_incidence_flux` post-processing, ideal detector, binning 1.  Every function cites the reference lines it
follows.  Pinned against the unmodified reference by oracle/make_golden_pyramid.py -> tests/golden/pyramid.npz
(tests/test_pyramid_oracle.py).  Nothing under rlao_b200/ imports this module.
"""
from __future__ import annotations

import numpy as np


def pyramid_phase_mask(resolution, n_subap, n_pix_separation, n_pix_edge):
    """Pyramid.py:368-388 (get_phase_mask, psf_centering=True, no quadrant shifts): the four tip/tilt faces."""
    n_tot = int((n_subap * 2 + n_pix_separation + n_pix_edge * 2) * resolution / n_subap)
    norma = (n_subap + n_pix_separation) * (resolution / n_subap)
    lim = np.pi / 4
    lim = lim - (np.pi / 4) / (n_tot // 2)
    tip, tilt = np.meshgrid(np.linspace(-lim, lim, n_tot // 2), np.linspace(-lim, lim, n_tot // 2))
    h = n_tot // 2
    m = np.zeros((n_tot, n_tot))
    m[:h, :h] = tip * norma + tilt * norma
    m[:h, -h:] = -tip * norma + tilt * norma
    m[-h:, -h:] = -tip * norma - tilt * norma
    m[-h:, :h] = tip * norma - tilt * norma
    return -m                                                    # :407 sign convention


class PyramidOracle:
    """wfs = PyramidOracle(pupil, fluxMap, nSubap, modulation, lightRatio, n_pix_separation, n_pix_edge)

    pupil [R, R] bool, fluxMap [R, R] photons per pixel and frame (Source.py:133-159).  `measure(phase)` takes the source
    phase [R, R] in radians (already masked by the pupil) and returns the slopes vector; `frame`, `signal_2D` are kept."""

    def __init__(self, pupil, fluxMap, nSubap, modulation, lightRatio, n_pix_separation=2, n_pix_edge=None,
                 calibModulation=50, reflectivity=None, postProcessing="slopesMaps"):
        assert postProcessing in ("slopesMaps", "slopesMaps_incidence_flux")
        self.postProcessing = postProcessing
        self.pupil = np.asarray(pupil).astype(bool)
        self.R = self.pupil.shape[0]
        if (self.R / nSubap) % 2 != 0:
            raise ValueError("The resolution should be an even number and be a multiple of 2**i where i>=2")     # :208-209
        self.fluxMap = np.asarray(fluxMap, dtype=np.float64)
        self.reflectivity = self.pupil.astype(float) if reflectivity is None else reflectivity
        self.nSubap = nSubap
        self.lightRatio = lightRatio
        self.n_pix_separation = n_pix_separation
        self.n_pix_edge = n_pix_separation // 2 if n_pix_edge is None else n_pix_edge                             # :237-240
        self.nRes = int((nSubap * 2 + self.n_pix_separation + self.n_pix_edge * 2) * self.R / nSubap)            # :250
        self.zeroPaddingFactor = self.nRes / self.R
        self.cam_resolution = round(nSubap * self.zeroPaddingFactor)                                             # :254
        self.center = self.nRes // 2
        self.calibModulation = self.R / 2 - 1 if calibModulation >= self.R / 2 else calibModulation              # :258-261
        # :285-288 modulation tip/tilt normalised in lambda/D
        self.Tip, self.Tilt = np.meshgrid(np.linspace(-np.pi, np.pi, self.R), np.linspace(-np.pi, np.pi, self.R))
        self.Tip, self.Tilt = self.Tip * self.pupil, self.Tilt * self.pupil
        xx, yy = np.meshgrid(np.arange(self.nRes, dtype=float), np.arange(self.nRes, dtype=float))
        self.phasor = np.exp(-(1j * np.pi * (self.nRes + 1) / self.nRes) * (xx + yy))                             # :291-292
        self.mask = np.exp(1j * pyramid_phase_mask(self.R, nSubap, self.n_pix_separation, self.n_pix_edge))      # :318-323
        self.referenceSignal_2D = 0.0
        self.slopesUnits = 1.0
        self.validI4Q = None
        # :301-314: valid pixels at a large modulation, then reference slopes at the working modulation
        self.set_modulation(self.calibModulation)
        self._propagate(np.zeros((self.R, self.R)))
        quads = [self.grab_quadrant(k) for k in (1, 2, 3, 4)]
        self.I4Q = quads[0] + quads[1] + quads[2] + quads[3]
        self.validI4Q = self.I4Q >= self.lightRatio * self.I4Q.max()                                            # :427-430
        self.validSignal = np.concatenate((self.validI4Q, self.validI4Q))
        self.nSignal = int(self.validSignal.sum())
        self.set_modulation(modulation)
        self._propagate(np.zeros((self.R, self.R)))                                                              # :456-460 flat wavefront
        self.referenceSignal_2D, self.referenceSignal = self.signal_processing()
        self.measure(np.zeros((self.R, self.R)))

    # ---- modulation ---------------------------------------------------------------------------------------
    def set_modulation(self, modulation):
        """Pyramid.py:941-975 (default path: circle of nTheta points, delta_theta = 0)."""
        self.modulation = modulation
        if modulation >= self.R // 2:
            raise ValueError("Error the modulation radius is too large for this resolution! Consider using a larger telescope resolution!")
        if modulation != 0:
            perimeter = np.pi * 2 * modulation
            self.nTheta = 4 * int(np.ceil(perimeter / 4))
            theta = np.linspace(0, 2 * np.pi, self.nTheta, endpoint=False)
            self.modulation_path = [(modulation * np.cos(t), modulation * np.sin(t)) for t in theta]
            # the reference stores the modulation phases as float32 (:960-961)
            self.phase_mod = np.stack([((px * self.Tip + py * self.Tilt) * self.pupil).astype(np.float32)
                                       for px, py in self.modulation_path]).astype(np.float64)
        else:
            self.nTheta = 1
            self.phase_mod = np.zeros((1, self.R, self.R))

    # ---- propagation --------------------------------------------------------------------------------------
    def pyramid_transform(self, phase):
        """Pyramid.py:469-504, psfCentering branch: |IFFT( FFT(padded field * phasor) * mask )|^2."""
        amp = np.sqrt(self.fluxMap / self.nTheta) * self.reflectivity                                            # :520
        support = np.zeros((self.nRes, self.nRes), dtype=complex)
        lo, hi = self.center - self.R // 2, self.center + self.R // 2
        support[lo:hi, lo:hi] = amp * np.exp(1j * phase)
        ft = np.fft.fft2(support * self.phasor)
        return np.abs(np.fft.ifft2(ft * self.mask)) ** 2

    def _propagate(self, phase):
        """Single-frame branches of wfs_measure (:533-538 and :581-603) + the camera binning of __mul__ (:987-1002)."""
        if self.modulation == 0:
            self.pyramidFrame = self.pyramid_transform(phase)
        else:
            self.pyramidFrame = sum(self.pyramid_transform(phase + pm) for pm in self.phase_mod)
        b = int(round(self.nRes / self.cam_resolution))
        n = self.cam_resolution
        self.frame = self.pyramidFrame.reshape(n, b, n, b).sum(-1).sum(1)                                        # tools.py:409-416

    def grab_quadrant(self, k):
        """Pyramid.py:774-791 (4-sided pyramid, binning 1)."""
        n_extra = int(np.round(self.n_pix_separation / 2))
        c = int(np.round(self.cam_resolution / 2))
        n = int(np.ceil(self.nSubap))
        f = self.frame
        if k == 3:
            return f[n_extra + c:n_extra + c + n, n_extra + c:n_extra + c + n]
        if k == 4:
            return f[n_extra + c:n_extra + c + n, -n_extra + c - n:-n_extra + c]
        if k == 1:
            return f[-n_extra + c - n:-n_extra + c, -n_extra + c - n:-n_extra + c]
        return f[-n_extra + c - n:-n_extra + c, n_extra + c:n_extra + c + n]

    def signal_processing(self):
        """Pyramid.py:685-701 (slopesMaps) and :703-726 (slopesMaps_incidence_flux)."""
        I1, I2, I3, I4 = (self.grab_quadrant(k) * self.validI4Q for k in (1, 2, 3, 4))
        I4Q = I1 + I2 + I3 + I4
        # :689-691 global normalisation; :713-716 `slopesMaps_incidence_flux` normalises by the mean of the camera frame
        self.norma = np.mean(I4Q[self.validI4Q]) if self.postProcessing == "slopesMaps" else np.float64(self.frame.mean())
        Sx = I1 - I2 + I4 - I3
        Sy = I1 - I4 + I2 - I3
        maps = (np.concatenate((Sx, Sy) / self.norma) - self.referenceSignal_2D) * self.slopesUnits
        return maps, maps[self.validSignal]

    def measure(self, phase):
        self._propagate(np.asarray(phase, dtype=np.float64))
        self.signal_2D, self.signal = self.signal_processing()
        return self.signal

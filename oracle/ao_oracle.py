"""CPU oracle: float64 numpy restatement of the closed-loop AO environment step that drl4ao drives
through OOPAO (Shack-Hartmann path).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke(), bench.py's `cpu_baseline` / `--impl reference` legs and
oracle/make_golden*.py may import this module; the product (rlao_b200) never does and has no CPU
fallback.  Every function cites the reference lines it restates (paths under
/root/reference/drl4ao/: OOPAO/ = AO_OOPAO/OOPAO/, MAIN/ = MAIN_CODE/).

Pinning: the reference ships no golden vectors for this path, so this restatement is
pinned against outputs of the *unmodified reference run in the build container* (oracle/make_golden.py
-> tests/golden/*.npz; checked by tests/test_oracle_golden.py).  The one boundary that cannot be pinned is
the third-party sub-pixel shift (scikit-image 0.18.3 `warp`), see oracle/warp018.py: PARITY UNPINNED there.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from numpy.random import RandomState
from scipy.special import kv

from .warp018 import warp_translate

# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------

# OOPAO/Source.py:164-242 (photometric system: wavelength [m], bandwidth [m], zero point [ph/m2/s])
PHOTOMETRY = {
    "V": (0.550e-6, 0.090e-6, 3.31e12),
    "R": (0.640e-6, 0.150e-6, 4.01e12),
    "I": (0.790e-6, 0.150e-6, 2.69e12),
    "J": (1.215e-6, 0.260e-6, 1.90e12),
    "H": (1.654e-6, 0.290e-6, 1.05e12),
    "K": (2.179e-6, 0.410e-6, 0.70e12),
}


@dataclass
class DetectorConfig:
    """OOPAO/Detector.py:13-28 arguments that matter on the WFS path."""
    photonNoise: bool = False
    readoutNoise: float = 0.0
    darkCurrent: float = 0.0
    QE: float = 1.0
    FWC: float | None = None
    bits: int | None = None
    gain: float = 1.0
    sensor: str = "CCD"


@dataclass
class AOConfig:
    nSubap: int = 20
    nPixPerSubap: int = 6
    diameter: float = 8.0
    samplingTime: float = 1.0 / 500
    centralObstruction: float = 0.0
    opticalBand: str = "I"
    magnitude: float = 8.0
    r0: float = 0.13
    L0: float = 25.0
    windSpeed: list = field(default_factory=lambda: [10.0])
    windDirection: list = field(default_factory=lambda: [0.0])
    fractionalR0: list = field(default_factory=lambda: [1.0])
    altitude: list = field(default_factory=lambda: [0.0])
    mechCoupling: float = 0.35
    dm_geometry: str = "cartesian"   # "cartesian": DeformableMirror(nSubap); "razor": env-style coords
    lightRatio: float = 0.5
    threshold_cog: float = 0.01
    nZernike: int = 50               # 0 -> zonal reconstructor (calib.M of the zonal IM)
    stroke: float = 1e-9
    nMeasurements: int = 25
    leak: float = 0.99
    gainCL: float = 0.5
    nLoop: int = 1000
    detector: DetectorConfig = field(default_factory=DetectorConfig)
    warp_kernel: str = "lagrange018"

    @property
    def resolution(self):
        return self.nSubap * self.nPixPerSubap


# ----------------------------------------------------------------------------------------------
# telescope / source
# ----------------------------------------------------------------------------------------------

def telescope_pupil(resolution, centralObstruction=0.0):
    """OOPAO/Telescope.py:164-180: circular pupil x^2+y^2 < ((R+1)/2)^2 on linspace(-R/2, R/2, R)."""
    x = np.linspace(-resolution / 2, resolution / 2, resolution)
    xx, yy = np.meshgrid(x, x)
    rr = xx ** 2 + yy ** 2
    d = resolution + 1
    return (rr < (d / 2) ** 2) & (rr >= (centralObstruction * d / 2) ** 2)


def source_properties(band, magnitude):
    """OOPAO/Source.py:99-108: wavelength and photon rate nPhoton = zeroPoint/368 * 10^(-0.4 mag)."""
    wl, _, zp = PHOTOMETRY[band]
    return wl, (zp / 368.0) * 10 ** (-0.4 * magnitude)


def flux_map(pupil, nPhoton, samplingTime, diameter):
    """OOPAO/Source.py:152: photons per pupil pixel per frame."""
    res = pupil.shape[0]
    return pupil.astype(float) * nPhoton * samplingTime * (diameter / res) ** 2


# ----------------------------------------------------------------------------------------------
# atmosphere (Assemat et al. 2006 infinite phase screens)
# ----------------------------------------------------------------------------------------------

def vk_covariance_matrix(z1, z2, L0, r0):
    """OOPAO/phaseStats.py:70-133 (makeCovarianceMatrix): von Karman phase covariance between two point
    sets given as complex coordinates; Bessel K_{5/6}."""
    rho = np.abs(z1[:, None] - z2[None, :])
    ratio = (L0 / r0) ** (5.0 / 3)
    g = math.gamma
    c0 = (24.0 * g(6.0 / 5) / 5) ** (5.0 / 6)
    cst = c0 * (g(11.0 / 6) / (2.0 ** (5.0 / 6) * np.pi ** (8.0 / 3))) * ratio
    out = np.full(rho.shape, c0 * (g(11.0 / 6) * g(5.0 / 6) / (2 * np.pi ** (8.0 / 3))) * ratio)
    nz = rho != 0
    u = 2 * np.pi * rho[nz] / L0
    out[nz] = cst * u ** (5.0 / 6) * kv(5.0 / 6, u)
    return out


def ring_masks(layer_res, nExtra=2):
    """OOPAO/Atmosphere.py:263-270: outer ring of the (layer_res+2)^2 map and the two rings inside it."""
    m = layer_res + nExtra
    outer = np.ones((m, m), dtype=bool)
    outer[1:-1, 1:-1] = False
    inner = ~outer
    inner[1 + nExtra:-1 - nExtra, 1 + nExtra:-1 - nExtra] = False
    return outer, inner


def atmosphere_operators(tel_res, diameter, L0, r0, r0_def=0.15, fov_px=0):
    """OOPAO/Atmosphere.py:213-220,263-286,485-558: predictor A = ZXt^T ZZt^-1 and innovation factor
    B = chol(XXt - A ZXt), covariances evaluated at r0_def and rescaled by (r0_def/r0)^(5/3)."""
    layer_res = tel_res + fov_px + 4
    layer_D = layer_res * diameter / tel_res
    outer, inner = ring_masks(layer_res)
    l = np.linspace(0, layer_res + 1, layer_res + 2) * layer_D / (layer_res - 1)
    u, v = np.meshgrid(l, l)
    innerZ = u[inner] + 1j * v[inner]
    outerZ = u[outer] + 1j * v[outer]
    ZZt = vk_covariance_matrix(innerZ, innerZ, L0, r0_def)
    ZXt = vk_covariance_matrix(innerZ, outerZ, L0, r0_def)
    XXt = vk_covariance_matrix(outerZ, outerZ, L0, r0_def)
    ZZt_inv = np.linalg.pinv(ZZt)
    s = (r0_def / r0) ** (5.0 / 3)
    A = (ZXt * s).T @ (ZZt_inv / s)
    BBt = XXt * s - A @ (ZXt * s)
    B = np.linalg.cholesky(BBt)
    return dict(A=A, B=B, XXt=XXt, ZXt=ZXt, outer=outer, inner=inner, layer_res=layer_res, layer_D=layer_D)


def rescale_B(ops, r0, r0_def=0.15):
    """OOPAO/Atmosphere.py:792-807 (r0 setter): only B is recomputed, A is r0-independent."""
    s = (r0_def / r0) ** (5.0 / 3)
    return np.linalg.cholesky(ops["XXt"] * s - ops["A"] @ (ops["ZXt"] * s))


def ft_phase_screen(r0, L0, N, delta, seed, l0=1e-10):
    """OOPAO/phaseStats.py:190-235: FFT screen; note the un-normalised forward FFT used as 'ift2' (:170-188)."""
    rs = RandomState(seed)
    del_f = 1.0 / (N * delta)
    fx = np.arange(-N / 2.0, N / 2.0) * del_f
    fx, fy = np.meshgrid(fx, fx)
    f = np.sqrt(fx ** 2 + fy ** 2)
    fm = 5.92 / l0 / (2 * np.pi)
    psd = 0.023 * r0 ** (-5.0 / 3) * np.exp(-((f / fm) ** 2)) / ((f ** 2 + (1.0 / L0) ** 2) ** (11.0 / 6))
    psd[int(N / 2), int(N / 2)] = 0
    cn = (rs.normal(size=(N, N)) + 1j * rs.normal(size=(N, N))) * np.sqrt(psd) * del_f
    return np.fft.fftshift(np.fft.fft2(np.fft.fftshift(cn))).real


def ft_sh_phase_screen(r0, L0, N, delta, seed, l0=1e-10):
    """OOPAO/phaseStats.py:243-318: FFT screen + 3 sub-harmonic grids; both parts are seeded with the
    same `seed` (:268,272) and only the (0..1, 0..1) corner of each 3x3 grid is summed (:306-309)."""
    rs = RandomState(seed)
    D = N * delta
    hi = ft_phase_screen(r0, L0, N, delta, seed, l0)
    coords = np.arange(-N / 2, N / 2) * delta
    x, y = np.meshgrid(coords, coords)
    lo = np.zeros(hi.shape, dtype=complex)
    fm = 5.92 / l0 / (2 * np.pi)
    for p in range(1, 4):
        del_f = 1 / (3 ** p * D)
        fx = np.arange(-1, 2) * del_f
        fx, fy = np.meshgrid(fx, fx)
        f = np.sqrt(fx ** 2 + fy ** 2)
        psd = 0.023 * r0 ** (-5.0 / 3) * np.exp(-((f / fm) ** 2)) / ((f ** 2 + (1.0 / L0) ** 2) ** (11.0 / 6))
        psd[1, 1] = 0
        cn = (rs.normal(size=(3, 3)) + 1j * rs.normal(size=(3, 3))) * np.sqrt(psd) * del_f
        sh = np.zeros((N, N), dtype=complex)
        for i in range(2):
            for j in range(2):
                sh += cn[i, j] * np.exp(1j * 2 * np.pi * (fx[i, j] * x + fy[i, j] * y))
        lo = lo + sh
    lo = lo.real - lo.real.mean()
    return lo + hi


class LayerState:
    pass


class AtmosphereOracle:
    """OOPAO/Atmosphere.py: multi-layer screens for an on-axis NGS (fov = 0, footprint centred)."""

    wavelength = 500e-9
    r0_def = 0.15

    def __init__(self, cfg: AOConfig, pupil, ops=None):
        self.cfg = cfg
        self.pupil = pupil
        self.R = cfg.resolution
        self.r0 = cfg.r0
        self.L0 = cfg.L0
        self.fractionalR0 = list(cfg.fractionalR0)
        self.nLayer = len(self.fractionalR0)
        self.ops = ops if ops is not None else atmosphere_operators(self.R, cfg.diameter, cfg.L0, cfg.r0, self.r0_def)
        self.A, self.B = self.ops["A"], self.ops["B"]
        self.outer, self.inner = self.ops["outer"], self.ops["inner"]
        self.layer_res = self.ops["layer_res"]
        self.delta = self.ops["layer_D"] / self.layer_res
        c = self.layer_res // 2
        self.fp = slice(c - self.R // 2, c + self.R // 2)          # Atmosphere.py:227-232 (centred footprint)
        self.layers = []
        self.xi_log = []          # every innovation vector drawn, in order (for identical-input GPU runs)
        self.xi_override = None   # iterator of vectors to use instead of the layer RNG
        for i in range(self.nLayer):
            self.layers.append(self._build_layer(i))
        # Atmosphere.py:185-189: new screens with seed 0, one update, publish OPD
        self.generateNewPhaseScreen(0)
        self.update()

    # -- helpers ------------------------------------------------------------------------------
    def _draw(self, ly):
        if self.xi_override is not None:
            xi = np.asarray(next(self.xi_override), dtype=np.float64)
        else:
            xi = ly.rng.normal(size=self.B.shape[1])
        self.xi_log.append(xi)
        return xi

    def _extrude(self, ly, interior):
        """X = A Z + B xi; map ring <- X, map interior <- `interior` (Atmosphere.py:288-293,307-310)."""
        Z = interior[self.inner[1:-1, 1:-1]]
        X = self.A @ Z + self.B @ self._draw(ly)
        ly.map[self.outer] = X
        ly.map[~self.outer] = interior.reshape(-1)

    def _build_layer(self, i):
        """Atmosphere.py:192-298."""
        cfg = self.cfg
        ly = LayerState()
        ly.rng = RandomState(42 + i * 1000)
        ly.windSpeed = cfg.windSpeed[i]
        ly.direction = cfg.windDirection[i]
        ly.vY = ly.windSpeed * np.cos(np.deg2rad(ly.direction))
        ly.vX = ly.windSpeed * np.sin(np.deg2rad(ly.direction))
        ly.map = np.zeros((self.layer_res + 2, self.layer_res + 2))
        ly.phase = ft_sh_phase_screen(self.r0, self.L0, self.layer_res, self.delta, seed=i)
        self._extrude(ly, ly.phase)
        ly.notDoneOnce = True
        ly.buff = np.zeros(2)
        ly.ratio = np.zeros(2)
        return ly

    # -- public -------------------------------------------------------------------------------
    def generateNewPhaseScreen(self, seed):
        """Atmosphere.py:560-592 (mode 2)."""
        for i, ly in enumerate(self.layers):
            ly.phase = ft_sh_phase_screen(self.r0, self.L0, self.layer_res, self.delta, seed=seed + i)
            ly.rng = RandomState(seed + i * 1000)
            self._extrude(ly, ly.phase)
            ly.notDoneOnce = True
        self._publish()

    def set_r0(self, r0):
        self.r0 = r0
        self.B = rescale_B(self.ops, r0, self.r0_def)

    def set_windSpeed(self, speeds):
        """Atmosphere.py:826-848."""
        for ly, v in zip(self.layers, speeds):
            ly.windSpeed = v
            ly.vY = v * np.cos(np.deg2rad(ly.direction))
            ly.vX = v * np.sin(np.deg2rad(ly.direction))
            ly.ratio[0] = ly.vX * self.cfg.samplingTime / self.delta
            ly.ratio[1] = ly.vY * self.cfg.samplingTime / self.delta

    def add_row(self, ly, step):
        """Atmosphere.py:301-311: shift the full map by one pixel (integer warp), re-extrude the ring."""
        shifted = warp_translate(ly.map, step[0], step[1], kernel=self.cfg.warp_kernel)[1:-1, 1:-1]
        self._extrude(ly, shifted)
        return shifted

    def update_layer(self, ly):
        """Atmosphere.py:350-407."""
        if ly.vX == 0 and ly.vY == 0:
            return
        if ly.notDoneOnce:
            ly.notDoneOnce = False
            ly.ratio = np.array([ly.vX * self.cfg.samplingTime / self.delta,
                                 ly.vY * self.cfg.samplingTime / self.delta])
            ly.buff = np.zeros(2)
        ratio = ly.ratio
        n = np.abs(ratio).astype(int)
        sgn = np.sign(ratio)
        for _ in range(n.min()):
            ly.phase = self.add_row(ly, sgn.copy())
        for _ in range(n.max() - n.min()):
            step = sgn.copy()
            step[n == n.min()] = 0
            ly.phase = self.add_row(ly, step)
        ly.buff = ly.buff + (np.abs(ratio) % 1) * sgn
        if np.abs(ly.buff[0]) >= 1 or np.abs(ly.buff[1]) >= 1:
            step = np.sign(ly.buff)
            step[np.abs(ly.buff) < 1] = 0
            ly.phase = self.add_row(ly, step)
        ly.buff = (np.abs(ly.buff) % 1) * np.sign(ly.buff)
        ly.phase = warp_translate(ly.map, ly.buff[0], ly.buff[1], kernel=self.cfg.warp_kernel)[1:-1, 1:-1]

    def _publish(self):
        """Atmosphere.py:439-450,474-478: sqrt(Cn2)-weighted sum over the footprint, radians -> metres."""
        acc = np.zeros((self.R, self.R))
        for ly, w in zip(self.layers, self.fractionalR0):
            acc += ly.phase[self.fp, self.fp] * np.sqrt(w)
        self.OPD_no_pupil = acc * self.wavelength / 2 / np.pi
        self.OPD = self.OPD_no_pupil * self.pupil

    def update(self):
        """Atmosphere.py:409-428."""
        for ly in self.layers:
            self.update_layer(ly)
        self._publish()


# ----------------------------------------------------------------------------------------------
# deformable mirror
# ----------------------------------------------------------------------------------------------

def dm_geometry(cfg: AOConfig, act_mask=None):
    """Actuator coordinates [m], valid flags and Gaussian width [px].

    cartesian: OOPAO/DeformableMirror.py:286-305 (nAct = nSubap+1 across D, valid iff
    r <= D/2 + 0.7533 pitch and outside the obstruction), pitch = D/nSubap (:266-270).
    razor: MAIN/OOPAOEnv/OOPAOEnvRazor.py:167-193 — explicit coordinates on linspace(-D/2, D/2, nAct) masked
    by `act_mask`, and DeformableMirror(nSubap=nActuator) so pitch = D/nActuator (:309-321 of the DM file).
    """
    D, R = cfg.diameter, cfg.resolution
    nAct = cfg.nSubap + 1
    x = np.linspace(-D / 2, D / 2, nAct)
    X, Y = np.meshgrid(x, x)
    xs, ys = X.reshape(-1), Y.reshape(-1)
    if cfg.dm_geometry == "cartesian":
        pitch = D / cfg.nSubap
        r = np.sqrt(xs ** 2 + ys ** 2)
        valid = (r > (cfg.centralObstruction * D / 2 - 0.5 * pitch)) & (r <= (D / 2 + 0.7533 * pitch))
        nAlong = nAct - 1
    else:
        pitch = D / nAct
        valid = np.asarray(act_mask, dtype=bool).reshape(-1)
        nAlong = D / pitch
    sigma = (R / nAlong) / np.sqrt(2 * np.log(1.0 / cfg.mechCoupling))
    return xs[valid], ys[valid], valid.reshape(nAct, nAct), sigma


def dm_modes(cfg: AOConfig, xIF, yIF, sigma):
    """OOPAO/DeformableMirror.py:343-346,494-514 with zero mis-registration: Gaussian influence functions on
    the grid linspace(0,1,R)*R, centred at R/2 + x*R/D; returns modes[R*R, nValidAct]."""
    R, D = cfg.resolution, cfg.diameter
    g = np.linspace(0, 1, R) * R
    u0x = R / 2 + xIF * R / D
    u0y = R / 2 + yIF * R / D
    a = 1.0 / (2 * sigma ** 2)
    gx = np.exp(-a * (g[None, :] - u0x[:, None]) ** 2)       # [nAct, R] along columns (X)
    gy = np.exp(-a * (g[None, :] - u0y[:, None]) ** 2)       # [nAct, R] along rows (Y)
    modes = gy[:, :, None] * gx[:, None, :]                  # [nAct, row, col]
    return modes.reshape(len(xIF), R * R).T.copy(), gx, gy


# ----------------------------------------------------------------------------------------------
# detector
# ----------------------------------------------------------------------------------------------

class DetectorOracle:
    """OOPAO/Detector.py:190-301 for one frame per readout (integrationTime <= samplingTime)."""

    def __init__(self, dcfg: DetectorConfig, integrationTime, seed=0):
        self.c = dcfg
        self.integrationTime = integrationTime
        self.rs_photon = RandomState(seed)
        self.rs_readout = RandomState(seed + 1)
        self.rs_dark = RandomState(seed + 2)
        self.frame = None

    def integrate(self, frame):
        c = self.c
        frame = np.array(frame, copy=True)
        if c.photonNoise:
            frame = self.rs_photon.poisson(frame)                       # :204-206
        frame = frame * c.QE                                            # :172-174
        if c.darkCurrent != 0:                                          # :224-229
            frame = frame + self.rs_dark.poisson(np.ones(frame.shape) * (c.darkCurrent * self.integrationTime))
        if c.FWC is not None:                                           # :177-181
            frame = np.clip(frame, 0, c.FWC)
        if c.sensor == "EMCCD":
            frame = frame * c.gain
        if c.readoutNoise != 0:                                         # :218-221
            frame = frame + np.round(self.rs_readout.randn(*frame.shape) * c.readoutNoise).astype(int)
        if c.sensor in ("CCD", "CMOS"):
            frame = frame * c.gain
        if c.bits is not None:                                          # :190-201
            if c.FWC is None:
                frame = (frame / frame.max() * 2 ** c.bits).astype(int)
            else:
                frame = (frame / c.FWC * (2 ** c.bits - 1)).astype(int)
                frame = np.clip(frame, frame.min(), 2 ** c.bits - 1)
        self.frame = frame
        return frame


# ----------------------------------------------------------------------------------------------
# Shack-Hartmann
# ----------------------------------------------------------------------------------------------

class ShackHartmannOracle:
    """OOPAO/ShackHartmann.py diffractive branch, binning_factor = 1, padding_extension_factor = 1, NGS."""

    def __init__(self, cfg: AOConfig, pupil, fluxMap, wavelength, detector: DetectorOracle | None = None,
                 valid_subapertures=None):
        self.cfg = cfg
        self.nS = cfg.nSubap
        self.n = cfg.resolution // cfg.nSubap
        self.N = 2 * self.n                                   # zero_padding = 2 (:141)
        self.R = cfg.resolution
        self.pupil = pupil
        self.wavelength = wavelength
        self.threshold_cog = cfg.threshold_cog
        self.cam = detector if detector is not None else DetectorOracle(DetectorConfig(), cfg.samplingTime)
        k = np.arange(self.N)
        xx, yy = np.meshgrid(k, k)
        self.phasor = np.exp(-(1j * np.pi * (self.N + 1) / self.N) * (xx + yy))     # :208-209
        self.set_flux(fluxMap)
        pps = self.photon_per_subap
        self.valid = (pps >= cfg.lightRatio * pps.max()).reshape(self.nS, self.nS)   # :228
        if valid_subapertures is not None:
            # MAIN/OOPAOEnv/OOPAOEnvRazor.py:241 overwrites only the 2-D mask; valid_subapertures_1D,
            # validLenslets_x/y and valid_slopes_maps keep the flux-based selection.  Not modelled.
            raise NotImplementedError
        self.valid_1d = self.valid.reshape(-1)
        self.vx, self.vy = np.nonzero(self.valid)
        self.nValid = int(self.valid.sum())
        self.nSignal = 2 * self.nValid
        self.valid_slopes_maps = np.concatenate((self.valid, self.valid))
        self.reference_slopes_maps = np.zeros((2 * self.nS, self.nS))
        self.slopes_units = 1.0
        self._initialize()

    # tiles: lenslet k = i*nS + j sees phase[i*n:(i+1)*n, j*n:(j+1)*n] TRANSPOSED (:340-347, hsplit of phase.T)
    def _tiles(self, img):
        nS, n = self.nS, self.n
        return img.reshape(nS, n, nS, n).transpose(0, 2, 3, 1).reshape(nS * nS, n, n)

    def set_flux(self, fluxMap):
        """:327-338 (the reference tiles fluxMap.T the same way as phase.T)."""
        self.cube_flux = self._tiles(fluxMap)
        self.photon_per_subap = self.cube_flux.sum(axis=(1, 2))

    def spots(self, phase):
        """:340-347,529-541: |FFT2 of the zero-padded lenslet field / N|^2 for every lenslet -> [nS^2, N, N]."""
        n, N = self.n, self.N
        lo = N // 2 - n // 2
        field = np.zeros((phase.shape[0] if phase.ndim == 3 else 1, self.nS ** 2, N, N), dtype=complex)
        ph = phase if phase.ndim == 3 else phase[None]
        for b in range(ph.shape[0]):
            field[b, :, lo:lo + n, lo:lo + n] = np.exp(1j * self._tiles(ph[b])) * np.sqrt(self.cube_flux)
        field *= self.phasor
        I = np.abs(np.fft.fft2(field, axes=(2, 3)) / N) ** 2
        return I if phase.ndim == 3 else I[0]

    def _bin(self, I):
        """:565 with tools.bin_ndarray (tools.py:309-343): 2x2 sum N x N -> n x n."""
        s = I.shape
        return I.reshape(*s[:-2], self.n, 2, self.n, 2).sum(axis=(-1, -3))

    @staticmethod
    def centroid(maps, threshold):
        """:314-324: global-max threshold, then first moments along axis 1 (->[:,0]) and axis 2 (->[:,1])."""
        im = np.array(maps, dtype=float, copy=True)
        im[im < threshold * im.max()] = 0
        i1 = np.arange(im.shape[1])[None, :, None]
        i2 = np.arange(im.shape[2])[None, None, :]
        with np.errstate(invalid="ignore", divide="ignore"):
            norma = im.sum(axis=(1, 2))
            out = np.stack([(im * i1).sum(axis=(1, 2)) / norma, (im * i2).sum(axis=(1, 2)) / norma], axis=1)
        out[~np.isfinite(out)] = 0                                                   # :583-593
        return out

    def _signal_from_centroids(self, cen):
        SX = np.zeros((self.nS, self.nS))
        SY = np.zeros((self.nS, self.nS))
        SX[self.vx, self.vy] = cen[:, 0]
        SY[self.vx, self.vy] = cen[:, 1]
        s2d = np.concatenate((SX, SY)) - self.reference_slopes_maps
        s2d[~self.valid_slopes_maps] = 0
        s2d = s2d / self.slopes_units
        return s2d, s2d[self.valid_slopes_maps]

    def measure(self, phase):
        """Single-frame branch :522-601.  `phase` is src.phase (already multiplied by the pupil)."""
        nS, n = self.nS, self.n
        maps = self._bin(self.spots(phase)[self.valid_1d])
        frame = np.zeros((self.R, self.R))
        for k, (i, j) in enumerate(zip(self.vx, self.vy)):                           # :349-353,571-574
            frame[i * n:(i + 1) * n, j * n:(j + 1) * n] = maps[k]
        frame = self.cam.integrate(frame)                                            # :576, :769-782
        self.frame = frame
        maps = frame.reshape(nS, n, nS, n).transpose(0, 2, 1, 3).reshape(nS * nS, n, n)[self.valid_1d]  # :355-362
        self.maps_intensity = maps
        cen = self.centroid(maps, self.threshold_cog)
        self.signal_2D, self.signal = self._signal_from_centroids(cen)
        return self.signal

    def measure_multi(self, phases, rs_photon=None, rs_readout=None):
        """Multi-frame branch :605-674 used by the interaction matrix: `phases` is [k, R, R]
        (phase_no_pupil), noise (if any) drawn on the binned spots, ONE global max over all frames."""
        k = phases.shape[0]
        maps = self._bin(self.spots(phases)[:, self.valid_1d]).reshape(k * self.nValid, self.n, self.n)
        c = self.cam.c
        if c.photonNoise:
            maps = rs_photon.poisson(maps)
        if c.readoutNoise != 0:
            maps = maps + np.int64(np.round(rs_readout.randn(*maps.shape) * c.readoutNoise))
        cen = self.centroid(maps, self.threshold_cog)
        sig = np.zeros((self.nSignal, k))
        for f in range(k):
            _, sig[:, f] = self._signal_from_centroids(cen[f * self.nValid:(f + 1) * self.nValid])
        self.signal = sig
        return sig

    def _initialize(self):
        """:254-312: reference slopes on a flat wavefront, then slope units from five tip ramps."""
        saved = (self.cam.c.photonNoise, self.cam.c.readoutNoise)
        self.cam.c.photonNoise, self.cam.c.readoutNoise = False, 0.0
        self.measure(np.zeros((self.R, self.R)))
        self.reference_slopes_maps = self.signal_2D.copy()
        R = self.R
        tip, _ = np.meshgrid(np.linspace(0, np.pi, R, endpoint=False), np.linspace(0, np.pi, R, endpoint=False))
        # :288-290 quirk: tel.pupil is stored as int (Telescope.py:391-392), so `Tip[tel.pupil]` is a fancy
        # index of rows 0/1, and the ramp is normalised by the std of the FULL ramp, not the in-pupil std.
        tip = tip / np.std(tip[self.pupil.astype(int)])
        amp = 10e-9
        mean_slope = np.zeros(5)
        for i in range(5):
            opd = self.pupil * tip * (i - 2) * amp
            self.measure(opd * 2 * np.pi / self.wavelength)
            mean_slope[i] = np.mean(self.signal[:self.nValid])
        p = np.polyfit(np.linspace(-2, 2, 5) * amp, mean_slope, deg=1)
        self.slopes_units = np.abs(p[0]) * (self.wavelength / 2 / np.pi)
        self.cam.c.photonNoise, self.cam.c.readoutNoise = saved
        self.measure(np.zeros((self.R, self.R)))


# ----------------------------------------------------------------------------------------------
# calibration
# ----------------------------------------------------------------------------------------------

def noll_to_nm(j):
    """Noll (1976) index -> (n, m); sign of m as aotools.zernike.zernIndex (even j: +, odd j: -)."""
    n = int((-1.0 + math.sqrt(8 * (j - 1) + 1)) / 2.0)
    p = j - (n * (n + 1)) // 2
    k = n % 2
    m = int((p + k) / 2.0) * 2 - k
    if m != 0 and j % 2 != 0:
        m = -m
    return n, m


def zernike_radial(n, m, r):
    out = np.zeros(r.shape)
    for s in range((n - m) // 2 + 1):
        c = ((-1) ** s) * math.factorial(n - s) / (
            math.factorial(s) * math.factorial((n + m) // 2 - s) * math.factorial((n - m) // 2 - s))
        out += c * r ** (n - 2 * s)
    return out


def zernike_modes(pupil, diameter, J):
    """OOPAO/Zernike.py:20-62: Noll modes 2..J+1 on the pupil pixels, mean-removed, unit std."""
    res = pupil.shape[0]
    X, Y = np.where(pupil > 0)
    X = (X - (res + res % 2 - 1) / 2) / res * diameter
    Y = (Y - (res + res % 2 - 1) / 2) / res * diameter
    Rr = np.sqrt(X ** 2 + Y ** 2)
    Rr = Rr / Rr.max()
    th = np.arctan2(Y, X)
    out = np.zeros((len(X), J))
    for i in range(1, J + 1):
        n, m = noll_to_nm(i + 1)
        if m == 0:
            Z = np.sqrt(n + 1) * zernike_radial(n, 0, Rr)
        elif m > 0:
            Z = np.sqrt(2 * (n + 1)) * zernike_radial(n, m, Rr) * np.cos(m * th)
        else:
            Z = np.sqrt(2 * (n + 1)) * zernike_radial(n, -m, Rr) * np.sin(-m * th)
        Z = Z - Z.mean()
        out[:, i - 1] = Z / np.std(Z)
    return out


def interaction_matrix(wfs: ShackHartmannOracle, modes, M2C, stroke, nMeasurements, wavelength):
    """OOPAO/calibration/InteractionMatrix.py:13-135, single_pass push, noise off: `nMeasurements` commands at
    a time through the multi-frame WFS branch; D = 2 * 0.5 * (s_push - 0)/stroke."""
    R = wfs.R
    nModes = M2C.shape[1]
    D = np.zeros((wfs.nSignal, nModes))
    saved = (wfs.cam.c.photonNoise, wfs.cam.c.readoutNoise)
    wfs.cam.c.photonNoise, wfs.cam.c.readoutNoise = False, 0.0
    nCycle = int(np.ceil(nModes / nMeasurements))
    nExtra = nModes % nMeasurements
    for c in range(nCycle):
        if c == nCycle - 1 and nExtra != 0:
            cols = slice(nModes - nExtra, nModes)
        else:
            cols = slice(c * nMeasurements, (c + 1) * nMeasurements)
        cmd = M2C[:, cols] * stroke
        opd = (modes @ cmd).T.reshape(-1, R, R)
        if opd.shape[0] == 1:
            # a single command goes through the single-frame branch (ndim(OPD) == 2)
            sp = wfs.measure(opd[0] * wfs.pupil * 2 * np.pi / wavelength)[:, None]
        else:
            sp = wfs.measure_multi(opd * 2 * np.pi / wavelength)
        D[:, cols] = 0.5 * sp / stroke
    wfs.cam.c.photonNoise, wfs.cam.c.readoutNoise = saved
    return 2 * D


def calibration_vault(D):
    """OOPAO/calibration/CalibrationVault.py:15-57: SVD pseudo-inverse M = V^T S^-1 U^T (no truncation)."""
    U, s, V = np.linalg.svd(D, full_matrices=False)
    M = V.T @ np.diag(1 / s) @ U.T
    return dict(D=U @ np.diag(s) @ V, M=M, s=s, cond=s[0] / s[-1])


# ----------------------------------------------------------------------------------------------
# PSF (reward for the ELT-scale config)
# ----------------------------------------------------------------------------------------------

def compute_psf(pupil, fluxMap, phase, zeroPaddingFactor):
    """OOPAO/Telescope.py:260-360 (computePSF -> PropagateField) with img_resolution = zp * R.

    Quirk kept: for an even image size the parity rule at :316-318 bumps `oversampling` from 1 to 2, so the
    field is padded to N = 2 * zp * R, and the PSF is the 2x2-binned |centred FFT / N|^2 (:353-355)."""
    R = pupil.shape[0]
    img_res = int(zeroPaddingFactor * R)
    oversampling = 1
    if zeroPaddingFactor * oversampling < 2:
        oversampling = int(np.ceil(2.0 / zeroPaddingFactor))
    if oversampling % 2 != img_res % 2:
        oversampling += 1
    img_size = int(np.ceil(img_res * oversampling))
    N = int(np.fix(zeroPaddingFactor * oversampling * R))
    pad = int(np.ceil((N - R) / 2))
    amp = pupil * pupil.astype(float) * np.sqrt(fluxMap)
    sup = np.pad(amp * np.exp(1j * phase), pad)
    N = sup.shape[0]
    k = np.arange(N)
    xx, yy = np.meshgrid(k, k)
    phasor = np.exp(-1j * np.pi / N * (xx + yy) * (1 - img_res % 2)).astype(np.complex64)
    emf = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(sup * phasor)) / N)
    if N % 2 == img_size % 2:
        shift_pix = 0
    else:
        shift_pix = 1 if N % 2 == 0 else -1
    lo = int(np.ceil(N / 2) - img_size // 2 + (1 - N % 2) - 1)
    hi = int(np.ceil(N / 2) + img_size // 2 + shift_pix)
    psf = np.abs(emf[lo:hi, lo:hi]) ** 2
    if oversampling != 1:
        m = psf.shape[0] // oversampling
        psf = psf.reshape(m, oversampling, m, oversampling).sum(axis=(1, 3))
    return psf


# ----------------------------------------------------------------------------------------------
# the environment (drl4ao gym-style wrapper)
# ----------------------------------------------------------------------------------------------

class EnvOracle:
    """MAIN/OOPAOEnv/OOPAOEnvRazor.py: set_params (:91-339, SH branch) and step (:474-514), one environment."""

    def __init__(self, cfg: AOConfig, atm_ops=None, act_mask=None, detector_seed=0, verbose=False,
                 reconstructor=None):
        """`atm_ops` / `reconstructor`: pre-computed operators (e.g. from a previous oracle instance) to skip the
        covariance and interaction-matrix computations when only the step itself is being timed."""
        self.cfg = cfg
        R = cfg.resolution
        self.pupil = telescope_pupil(R, cfg.centralObstruction)
        self.wavelength, self.nPhoton = source_properties(cfg.opticalBand, cfg.magnitude)
        self.fluxMap = flux_map(self.pupil, self.nPhoton, cfg.samplingTime, cfg.diameter)
        self.atm = AtmosphereOracle(cfg, self.pupil, ops=atm_ops)
        self.atm.update()                                           # OOPAOEnvRazor.py:159
        xIF, yIF, self.dm_mask, self.sigma = dm_geometry(cfg, act_mask)
        self.modes, self.gx, self.gy = dm_modes(cfg, xIF, yIF, self.sigma)
        self.nValidAct = self.modes.shape[1]
        self.nActuator = cfg.nSubap + 1
        self.xvalid, self.yvalid = np.nonzero(self.dm_mask)
        # The WFS initialises itself (reference slopes, slope units) on the default ideal detector; the env
        # applies the camera settings only afterwards (OOPAOEnvRazor.py:235-250), noise flags last (:332-333).
        self.cam = DetectorOracle(DetectorConfig(), cfg.samplingTime, seed=detector_seed)
        self.wfs = ShackHartmannOracle(cfg, self.pupil, self.fluxMap, self.wavelength, self.cam)
        self.cam.c = DetectorConfig(**{**cfg.detector.__dict__, "photonNoise": False, "readoutNoise": 0.0})
        if reconstructor is not None:
            self.reconstructor = np.asarray(reconstructor, dtype=float)
        else:
            if cfg.nZernike > 0:
                Z = zernike_modes(self.pupil, cfg.diameter, cfg.nZernike)
                self.M2C = np.linalg.pinv(self.modes[self.pupil.reshape(-1), :]) @ Z      # :261
            else:
                self.M2C = np.eye(self.nValidAct)
            self.D_zonal = interaction_matrix(self.wfs, self.modes, np.eye(self.nValidAct), cfg.stroke,
                                              cfg.nMeasurements, self.wavelength)
            self.calib = calibration_vault(self.D_zonal @ self.M2C)                       # :290
            self.reconstructor = self.M2C @ self.calib["M"]                               # :336
            self.F = self.M2C @ np.linalg.pinv(self.M2C)                                  # :337
        self.cam.c.photonNoise = cfg.detector.photonNoise                             # :332-333
        self.cam.c.readoutNoise = cfg.detector.readoutNoise
        self.leak = cfg.leak
        self.gainCL = cfg.gainCL
        self.coefs = np.zeros(self.nValidAct)
        self.dm_prev = self.coefs.copy()
        self.dm_OPD = np.zeros((R, R))
        self.total = np.zeros(cfg.nLoop)
        self.residual = np.zeros(cfg.nLoop)
        self.SR = []
        # state of the paired telescope
        # :297-301: `source*tel*dm*wfs` runs un-paired (flat wavefront, zero signal), THEN tel+atm pairs them
        self.phase = np.zeros((R, R))
        self.wfs.measure(self.phase)
        self.tel_OPD_no_pupil = self.atm.OPD_no_pupil.copy()
        self.tel_OPD = self.atm.OPD.copy()

    # -- pieces of the optical train ------------------------------------------------------------
    def set_coefs(self, coefs):
        """OOPAO/DeformableMirror.py:534-570: dm.OPD = modes @ coefs."""
        self.coefs = np.zeros(self.nValidAct) if np.isscalar(coefs) else np.asarray(coefs, dtype=float)
        self.dm_OPD = (self.modes @ self.coefs).reshape(self.cfg.resolution, self.cfg.resolution)

    def atm_update(self):
        self.atm.update()
        self.tel_OPD_no_pupil = self.atm.OPD_no_pupil.copy()          # Atmosphere.py:427-428,666-667
        self.tel_OPD = self.atm.OPD.copy()

    def propagate(self):
        """tel*dm*wfs: OOPAO/Telescope.py:533-544 + DeformableMirror.py:452-478 + ShackHartmann.py:511."""
        self.tel_OPD_no_pupil = self.tel_OPD_no_pupil + self.dm_OPD
        self.tel_OPD = self.tel_OPD_no_pupil * self.pupil
        self.phase = self.tel_OPD * 2 * np.pi / self.wavelength
        return self.wfs.measure(self.phase)

    # -- gym-style API ----------------------------------------------------------------------------
    def vec_to_img(self, v):
        img = np.zeros((self.nActuator, self.nActuator))
        img[self.xvalid, self.yvalid] = v
        return img

    def img_to_vec(self, img):
        return img[self.xvalid, self.yvalid]

    def get_strehl(self):
        return np.exp(-np.var(self.phase[self.pupil]))

    def reset_soft(self):
        return self.vec_to_img(-self.reconstructor @ self.wfs.signal) * 1e6

    def new_episode(self, seed):
        """Caller pattern MAIN/PO4AO/mbrl.py:49-55 / MAIN/integrator_oopao_razor.py:36-60."""
        self.atm.generateNewPhaseScreen(seed)
        self.tel_OPD_no_pupil = self.atm.OPD_no_pupil.copy()
        self.tel_OPD = self.atm.OPD.copy()
        self.set_coefs(0)
        self.dm_prev = self.coefs.copy()
        self.propagate()
        return self.reset_soft()

    def step(self, i, action):
        """MAIN/OOPAOEnv/OOPAOEnvRazor.py:474-514."""
        action = self.img_to_vec(np.asarray(action, dtype=float)) * 1e-6
        self.atm_update()
        self.total[i] = np.std(self.tel_OPD[self.pupil]) * 1e9
        self.propagate()
        self.set_coefs(self.dm_prev * self.leak + action)
        self.dm_prev = self.coefs.copy()
        obs = self.vec_to_img(-self.reconstructor @ self.wfs.signal) * 1e6
        self.residual[i] = np.std(self.tel_OPD[self.pupil]) * 1e9
        strehl = self.get_strehl()
        self.SR.append(strehl)
        return obs, -1 * np.linalg.norm(obs), strehl, False, {"strehl": strehl}

    def sample_noise(self, sigma, rs=np.random):
        """:616-619."""
        return self.vec_to_img(self.F @ (sigma * rs.normal(0, 1, size=(self.nValidAct,))))

    def calculate_strehl_AVG(self):
        avg = np.mean(self.SR)
        self.SR = []
        return avg

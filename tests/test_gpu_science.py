"""Science path on the device: the whole PSF image of tel.computePSF through the library's
kernels (aoenv_psf_image: both transforms as tensor-core GEMMs), the science camera with its exposure buffer and binning
(OOPAO/Telescope.py:260-360,487-500; OOPAO/Detector.py:232-301), against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle.ao_oracle import compute_psf, flux_map, source_properties, telescope_pupil
from oracle.golden_configs import CONFIGS
from parity_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _np(t):
    return t.detach().double().cpu().numpy()


def _telescope(dev, R, n_envs):
    from rlao_b200.Source import Source
    from rlao_b200.Telescope import Telescope
    tel = Telescope(R, 8.0, 1 / 500, n_envs=n_envs, device=dev)
    Source("I", 8) * tel
    return tel


@pytest.mark.parametrize("R,zp", [(48, 4), (48, 2), (120, 4), (120, 6)])
def test_compute_psf_full_image_vs_oracle(dev, R, zp):
    tel = _telescope(dev, R, 2)
    pupil = telescope_pupil(R)
    wl, nph = source_properties("I", 8)
    fm = flux_map(pupil, nph, 1 / 500, 8.0)
    rs = np.random.RandomState(R + zp)
    yy, xx = np.mgrid[:R, :R] / R
    opds = [90e-9 * (np.sin(9 * xx) + np.cos(7 * yy * xx) + c * xx) + 10e-9 * rs.normal(size=xx.shape) for c in (0.5, -2.0)]
    tel.OPD_no_pupil = torch.as_tensor(np.stack(opds), dtype=torch.float32, device=dev)
    opd32 = _np(tel.OPD_no_pupil)
    tel.computePSF(zp)
    assert tel.PSF.shape == (2, zp * R, zp * R)
    for e in range(2):
        want = compute_psf(pupil, fm, opd32[e] * pupil * 2 * np.pi / wl, zp)
        assert rel_err(_np(tel.PSF[e]), want) < 2e-5, e
        assert abs(float(tel.PSF_norma[e].max()) - 1) < 1e-6
    # a smaller image is the central crop of the same PSF
    full = tel.PSF.clone()
    tel.computePSF(zp, img_resolution=R)
    lo = zp * R // 2 - R // 2
    assert rel_err(_np(tel.PSF), _np(full[:, lo:lo + R, lo:lo + R])) < 2e-6
    # and the pruned-DFT Strehl kernel sees the same core
    from rlao_b200.psf import psf_peak
    a, b = tel._terms()
    peak = psf_peak(tel, a.contiguous(), b, zp, 16)
    assert rel_err(_np(peak), _np(full.amax(dim=(-2, -1)))) < 2e-5


def test_compute_psf_against_reference_golden(dev):
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny.npz"))
    cfg = CONFIGS["tiny"]()
    tel = _telescope(dev, cfg.resolution, 1)
    wl, _ = source_properties(cfg.opticalBand, cfg.magnitude)
    ph = torch.as_tensor(gold["psf_atm_phase"], dtype=torch.float32, device=dev)
    tel.OPD_no_pupil = ph * (wl / 2 / np.pi)
    tel.computePSF(4)
    cc = tel.PSF.shape[-1] // 2
    assert rel_err(_np(tel.PSF[cc - 8:cc + 8, cc - 8:cc + 8]), gold["psf_atm_crop"]) < 1e-4
    assert abs(float(tel.PSF.max()) - gold["psf_atm_max"]) < 1e-4 * gold["psf_atm_max"]


def test_science_camera_exposure_binning_and_noise(dev):
    """tel*cam (Telescope.py:487-500): sub-frames accumulate until the integration time is reached; binning sums b x b
    pixels; photon noise is drawn per sub-frame (mean = variance = summed flux), read noise once per readout."""
    from rlao_b200.Detector import Detector
    R, B = 48, 64
    tel = _telescope(dev, R, B)
    tel.resetOPD()
    tel.computePSF(2)
    psf = tel.PSF.clone()
    cam = Detector(integrationTime=3 * tel.samplingTime, psf_sampling=2, binning=4)
    tel * cam
    tel * cam
    assert cam.frame is None and cam.n_buffered == 2
    tel * cam
    want = 3 * psf.reshape(B, R // 2, 4, R // 2, 4).sum(dim=(2, 4))
    assert cam.frame.shape == (B, R // 2, R // 2) and torch.allclose(cam.frame, want, rtol=1e-5)
    assert cam.n_buffered == 0 and cam._integrated_time == 0
    # cropped camera: Detector(nRes) keeps the central nRes pixels
    small = Detector(R, psf_sampling=2)
    tel * small
    lo = R - R // 2
    assert torch.allclose(small.frame, psf[:, lo:lo + R, lo:lo + R], rtol=1e-5)
    # photon + read noise, two sub-frames per exposure
    noisy = Detector(integrationTime=2 * tel.samplingTime, psf_sampling=2, photonNoise=True, readoutNoise=3.0, QE=0.8, seed=5)
    tel * noisy
    tel * noisy
    f = noisy.frame.double()
    lam = 2 * 0.8 * psf[0].double()
    bright = lam > 50
    mean, var = f.mean(dim=0), f.var(dim=0)
    assert abs(float((mean[bright] / lam[bright]).mean()) - 1) < 0.02
    # Poisson(x) * QE twice + rounded N(0, 3): variance = QE^2 * 2 psf + 9
    want_var = 0.8 ** 2 * 2 * psf[0].double() + 9.0
    assert abs(float((var[bright] / want_var[bright]).mean()) - 1) < 0.1
    assert bool((f == f.round()).all()) is False or noisy.QE == 1        # QE scales the integer photon counts

"""The Pyramid's own transform kernels (aoenv_pyramid_frames: hand-written N = 16 x N2 FFTs, OOPAO/Pyramid.py:469-504,
581-603, 987-1002) against the float64 library transform of the same chain, at the papyrus size (20 x 20, N = 288, 20
modulation points) and the test size (12 x 12, N = 128); the reference fixture is checked in test_pyramid_oracle.py."""
import numpy as np
import pytest
import torch

from parity_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _np(t):
    return t.detach().double().cpu().numpy()


@pytest.mark.parametrize("nS,R,modulation", [(20, 120, 3), (20, 120, 0), (12, 48, 3), (12, 48, 5)])
def test_pyramid_kernels_vs_float64_transforms(dev, nS, R, modulation):
    from rlao_b200 import _lib
    from rlao_b200.Pyramid import Pyramid
    from rlao_b200.Source import Source
    from rlao_b200.Telescope import Telescope
    B = 3
    tel = Telescope(R, 8.0, 1 / 500, n_envs=B, device=dev)
    Source("I", 8) * tel
    wfs = Pyramid(nS, tel, modulation, 0.1, n_pix_separation=4, n_pix_edge=2)
    assert wfs.nRes in (128, 288) and wfs._kernels_ok()
    rs = np.random.RandomState(nS + modulation)
    yy, xx = np.mgrid[:R, :R] / R
    opd = np.stack([0.3e-6 * (c[0] * xx + c[1] * yy + c[2] * np.sin(6 * xx + 2 * yy)) + 0.03e-6 * rs.normal(size=(R, R))
                    for c in rs.normal(size=(B, 3))])
    a = torch.as_tensor(opd, dtype=torch.float32, device=dev).contiguous()
    n0 = _lib.launch_count()
    got = wfs._frames_kernels(a, None)
    assert _lib.launch_count() - n0 == 4                        # columns, rows, image, binning
    lam = tel.src.wavelength
    want = wfs._frames(a.double() * tel._pupil_f.double() * (2 * np.pi / lam), precise=True)
    assert got.shape == want.shape == (B, wfs.cam.resolution, wfs.cam.resolution)
    assert rel_err(_np(got), _np(want)) < 2e-5
    # two terms (atmosphere + DM) add up inside the kernel
    got2 = wfs._frames_kernels((0.4 * a).contiguous(), (0.6 * a).contiguous())
    assert rel_err(_np(got2), _np(want)) < 3e-5
    # through the reference-facing API: tel*wfs uses the kernels, slopes agree with the library path
    tel.OPD_no_pupil = a
    tel * wfs
    sig_k = _np(wfs.signal)
    wfs.use_kernels = False
    tel * wfs
    assert rel_err(sig_k, _np(wfs.signal)) < 2e-4

"""Pins oracle/pyramid_oracle.py (groundwork for the Pyramid WFS) against a fixture produced by
the UNMODIFIED reference OOPAO/Pyramid.py (tests/golden/pyramid.npz, written by oracle/make_golden_pyramid.py)."""
import os

import numpy as np
import pytest

from oracle.pyramid_oracle import PyramidOracle, pyramid_phase_mask

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pyramid.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def wfs(g):
    R, nS, mod, sep, edge = (int(x) for x in g["setup"])
    return PyramidOracle(g["pupil"].astype(bool), g["fluxMap"], nS, mod, float(g["lightRatio"]), n_pix_separation=sep,
                         n_pix_edge=edge)


def _rel(a, b):
    return np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / max(float(np.abs(b).max()), 1e-300)


def test_geometry_mask_and_valid_pixels(g, wfs):
    R, nS, mod, sep, edge = (int(x) for x in g["setup"])
    assert wfs.nRes == int(g["nRes"]) and wfs.nTheta == int(g["nTheta"]) and wfs.cam_resolution == int(g["cam_resolution"])
    assert _rel(pyramid_phase_mask(R, nS, sep, edge), g["mask_phase"]) < 1e-14
    assert np.array_equal(wfs.validI4Q, g["validI4Q"].astype(bool)) and wfs.nSignal == int(g["nSignal"])
    assert _rel(wfs.referenceSignal_2D, g["referenceSignal_2D"]) < 1e-9


def test_frames_and_slopes(g, wfs):
    wfs.measure(np.zeros_like(g["phase_0"]))
    assert _rel(wfs.frame, g["frame_flat"]) < 1e-9
    assert np.abs(wfs.signal - g["signal_flat"]).max() < 1e-9
    for k in range(2):
        sig = wfs.measure(g[f"phase_{k}"])
        assert _rel(wfs.frame, g[f"frame_{k}"]) < 1e-9, k
        assert _rel(sig, g[f"signal_{k}"]) < 1e-8, k
        assert _rel(wfs.signal_2D, g[f"signal_2D_{k}"]) < 1e-8, k


def test_unmodulated(g, wfs):
    wfs.set_modulation(0)
    wfs.referenceSignal_2D = 0.0
    wfs._propagate(np.zeros_like(g["phase_0"]))
    wfs.referenceSignal_2D, wfs.referenceSignal = wfs.signal_processing()
    assert _rel(wfs.referenceSignal_2D, g["referenceSignal_2D_unmodulated"]) < 1e-9
    assert _rel(wfs.measure(g["phase_0"]), g["signal_unmodulated_0"]) < 1e-8


def test_incidence_flux_normalisation(g):
    """postProcessing='slopesMaps_incidence_flux' (the papyrus parameter file, parameterFile_oopao_parser.py:68)."""
    R, nS, mod, sep, edge = (int(x) for x in g["setup"])
    w = PyramidOracle(g["pupil"].astype(bool), g["fluxMap"], nS, mod, float(g["lightRatio"]), n_pix_separation=sep,
                      n_pix_edge=edge, postProcessing="slopesMaps_incidence_flux")
    assert _rel(w.referenceSignal_2D, g["incidence_referenceSignal_2D"]) < 1e-9
    assert _rel(w.measure(g["phase_1"]), g["incidence_signal_1"]) < 1e-8


# ---- product class (torch / cuFFT implementation) against the same fixture --------------------------------------
def _build_product(g, device, n_envs=1):
    from rlao_b200.Pyramid import Pyramid
    from rlao_b200.Source import Source
    from rlao_b200.Telescope import Telescope
    R, nS, mod, sep, edge = (int(x) for x in g["setup"])
    tel = Telescope(R, 8.0, 1 / 500, n_envs=n_envs, device=device)
    src = Source("I", 8.0)
    src * tel
    assert np.array_equal(tel.pupil.astype(bool), g["pupil"].astype(bool))
    assert _rel(src.fluxMap, g["fluxMap"]) < 1e-12
    wfs = Pyramid(nS, tel, mod, float(g["lightRatio"]), n_pix_separation=sep, n_pix_edge=edge)
    return tel, src, wfs


def _check_product(g, tel, wfs, tol_frame, tol_sig):
    import torch
    assert wfs.nRes == int(g["nRes"]) and wfs.nTheta == int(g["nTheta"]) and wfs.nSignal == int(g["nSignal"])
    assert np.array_equal(wfs.validI4Q, g["validI4Q"].astype(bool))
    assert _rel(wfs.referenceSignal_2D.cpu().numpy(), g["referenceSignal_2D"]) < 1e-9
    B = tel.n_envs
    tel.resetOPD()
    tel * wfs
    frame = wfs.cam.frame if B == 1 else wfs.cam.frame[0]
    assert _rel(frame.cpu().numpy(), g["frame_flat"]) < tol_frame
    lam = float(g["wavelength"])
    phases = np.stack([g["phase_0"], g["phase_1"]] + [g["phase_0"]] * max(0, B - 2))[:max(B, 1)]
    if B == 1:
        for k in range(2):
            wfs.wfs_measure(phase_in=g[f"phase_{k}"])
            assert _rel(wfs.cam.frame.cpu().numpy(), g[f"frame_{k}"]) < tol_frame
            assert _rel(wfs.signal.cpu().numpy(), g[f"signal_{k}"]) < tol_sig
            assert _rel(wfs.signal_2D.cpu().numpy(), g[f"signal_2D_{k}"]) < tol_sig
    else:
        tel.OPD_no_pupil = torch.as_tensor(phases * lam / (2 * np.pi), dtype=torch.float32, device=tel.device)
        tel * wfs
        for k in range(2):
            assert _rel(wfs.cam.frame[k].cpu().numpy(), g[f"frame_{k}"]) < tol_frame
            assert _rel(wfs.signal[k].cpu().numpy(), g[f"signal_{k}"]) < tol_sig


def test_product_incidence_flux_on_cpu_stand_in(g, monkeypatch):
    import fake_backend
    from rlao_b200.Pyramid import Pyramid
    fake_backend.install(monkeypatch)
    tel, src, _ = _build_product(g, None)
    R, nS, mod, sep, edge = (int(x) for x in g["setup"])
    w = Pyramid(nS, tel, mod, float(g["lightRatio"]), n_pix_separation=sep, n_pix_edge=edge,
                postProcessing="slopesMaps_incidence_flux")
    assert _rel(w.referenceSignal_2D.numpy(), g["incidence_referenceSignal_2D"]) < 1e-9
    w.wfs_measure(phase_in=g["phase_1"])
    assert _rel(w.signal.numpy(), g["incidence_signal_1"]) < 2e-4


def test_product_pyramid_on_cpu_stand_in(g, monkeypatch):
    """Host logic of rlao_b200.Pyramid with the objects forced onto the CPU (torch.fft there): same fixture."""
    import fake_backend
    fake_backend.install(monkeypatch)
    tel, src, wfs = _build_product(g, None)
    _check_product(g, tel, wfs, 2e-5, 2e-4)


@pytest.mark.gpu
def test_product_pyramid_on_gpu_batched(g):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    tel, src, wfs = _build_product(g, torch.device("cuda:0"), n_envs=3)
    _check_product(g, tel, wfs, 2e-5, 2e-4)
    # noisy camera: photon noise keeps the flux, slopes stay close
    ideal = wfs.signal.clone()
    wfs.cam.photonNoise = True
    tel * wfs
    assert torch.equal(wfs.cam.frame.round(), wfs.cam.frame)
    assert float((wfs.signal - ideal).abs().max()) > 0 and float((wfs.signal - ideal).abs().max()) < 0.2


def test_product_modulation_setter_recalibrates(g, monkeypatch):
    """wfs.modulation = 0 on the fly (Pyramid.py:941-984): new reference slopes, unmodulated measurement."""
    import fake_backend
    fake_backend.install(monkeypatch)
    tel, src, wfs = _build_product(g, None)
    wfs.modulation = 0
    assert wfs.nTheta == 1
    assert _rel(wfs.referenceSignal_2D.numpy(), g["referenceSignal_2D_unmodulated"]) < 1e-9
    wfs.wfs_measure(phase_in=g["phase_0"])
    assert _rel(wfs.signal.numpy(), g["signal_unmodulated_0"]) < 5e-4
    with pytest.raises(ValueError):
        wfs.modulation = tel.resolution

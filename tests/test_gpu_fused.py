"""The fused step kernel (aoenv_shwfs_fused: DM surface + SH spots + slopes + pupil statistics in one cluster launch)
against the oracle and against the unfused kernels (aoenv_dm_surface_separable + aoenv_shwfs_frame + aoenv_shwfs_slopes),
for every compiled lenslet size, several cluster / warp-group shapes, with and without the camera frame, and with the
noisy camera in between."""
import numpy as np
import pytest
import torch

from oracle.ao_oracle import AOConfig, ShackHartmannOracle, dm_geometry, dm_modes, flux_map, source_properties, telescope_pupil
from parity_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _np(t):
    return t.detach().double().cpu().numpy()


def _objects(dev, nS, n, B, with_dm=True):
    from rlao_b200.DeformableMirror import DeformableMirror
    from rlao_b200.ShackHartmann import ShackHartmann
    from rlao_b200.Source import Source
    from rlao_b200.Telescope import Telescope
    cfg = AOConfig(nSubap=nS, nPixPerSubap=n)
    tel = Telescope(cfg.resolution, cfg.diameter, cfg.samplingTime, n_envs=B, device=dev)
    Source(cfg.opticalBand, cfg.magnitude) * tel
    wfs = ShackHartmann(nS, tel, cfg.lightRatio)
    dm = DeformableMirror(tel, nS, cfg.mechCoupling) if with_dm else None
    if dm is not None:
        dm.lazy_surface = True                  # the opt-in mode these tests are about (AOENV_WFS=fused)
    return cfg, tel, wfs, dm


def _wavefronts(R, B, seed, amp=0.4e-6):
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[:R, :R] / R
    out = []
    for _ in range(B):
        c = rs.normal(size=6)
        out.append(amp * (c[0] * xx + c[1] * yy + c[2] * np.sin(7 * xx + 3 * yy) + c[3] * np.cos(11 * yy) * xx
                          + 0.3 * c[4] * np.sin(23 * xx * yy)) + 0.05e-6 * rs.normal(size=(R, R)))
    return np.stack(out)


def _measure(wfs, opd, second, fused, keep_frame=False):
    wfs.use_fused, wfs.keep_frame = fused, keep_frame
    wfs._measure_terms(opd, second, 0)
    torch.cuda.synchronize()
    return _np(wfs.signal).copy(), _np(wfs._stats).copy()


@pytest.mark.parametrize("nS,n", [(8, 6), (5, 4), (4, 8), (20, 6), (10, 8), (12, 4)])
def test_fused_equals_unfused_kernels_and_oracle(dev, nS, n):
    B = 3
    cfg, tel, wfs, dm = _objects(dev, nS, n, B)
    R = cfg.resolution
    assert wfs._uniform_flux and dm.fused_tables() is not None
    opd = torch.as_tensor(_wavefronts(R, B, 5), dtype=torch.float32, device=dev).contiguous()
    rs = np.random.RandomState(3)
    dm.coefs = torch.as_tensor(rs.normal(size=(B, dm.nValidAct)) * 1.5e-7, dtype=torch.float32, device=dev)
    ref = dm.surface_ref()
    assert not dm._valid[dm._slot]                                # nothing has been written for this surface yet
    sig_f, st_f = _measure(wfs, opd, ref, True)
    assert not dm._valid[dm._slot]                                # ... and the fused measurement did not need it
    frame_f = _np(wfs.cam.frame).copy()                           # produced on demand (frame-only pass of the kernel)
    sig_k, _ = _measure(wfs, opd, ref, True, keep_frame=True)
    assert np.array_equal(sig_k, sig_f)
    assert np.array_equal(_np(wfs.cam.frame), frame_f)            # same kernel, frame written in the same pass
    surf = dm.OPD.reshape(B, R, R)                                # aoenv_dm_surface_separable
    sig_u, st_u = _measure(wfs, opd, surf, False)
    frame_u = _np(wfs.cam.frame).copy()
    assert rel_err(frame_f, frame_u) < 5e-6
    assert rel_err(sig_f, sig_u) < 2e-5
    # pupil statistics: plain sums (fused) and centred sums (unfused) give the same variances
    npup = float(tel.pixelArea)
    var = lambda s, i: s[:, i + 1] / npup - (s[:, i] / npup) ** 2
    for i in (0, 2):
        assert np.allclose(var(st_f, i), var(st_u, i), rtol=2e-5, atol=0)
    # explicit second term through the fused kernel (the general-modes path)
    sig_t, _ = _measure(wfs, opd, surf, True)
    assert rel_err(sig_t, sig_f) < 2e-5
    # oracle on the same total OPD
    pupil = telescope_pupil(R)
    wl, nph = source_properties(cfg.opticalBand, cfg.magnitude)
    orc = ShackHartmannOracle(cfg, pupil, flux_map(pupil, nph, cfg.samplingTime, cfg.diameter), wl)
    wfs_units = wfs.slopes_units
    total = _np(opd) + _np(surf)
    for e in range(B):
        want = orc.measure(total[e] * pupil * 2 * np.pi / wl) * orc.slopes_units / wfs_units
        assert rel_err(frame_f[e], orc.frame) < 2e-5, (e, "frame")
        assert rel_err(sig_f[e], want) < 1e-4, (e, "slopes")


@pytest.mark.parametrize("cluster,groups", [(1, 4), (2, 4), (4, 2), (5, 4), (7, 4), (10, 6), (13, 2)])
def test_fused_cluster_and_group_shapes_agree(dev, monkeypatch, cluster, groups):
    nS, n, B = 20, 6, 4
    cfg, tel, wfs, dm = _objects(dev, nS, n, B)
    opd = torch.as_tensor(_wavefronts(cfg.resolution, B, 9), dtype=torch.float32, device=dev).contiguous()
    dm.coefs = torch.as_tensor(np.random.RandomState(1).normal(size=(B, dm.nValidAct)) * 1e-7, dtype=torch.float32, device=dev)
    ref = dm.surface_ref()
    base_sig, base_st = _measure(wfs, opd, ref, True, keep_frame=True)
    base_frame = _np(wfs.cam.frame).copy()
    monkeypatch.setenv("AOENV_WFS_CLUSTER", str(cluster))
    monkeypatch.setenv("AOENV_WFS_GROUPS", str(groups))
    wfs._fused_plans = {}
    sig, st = _measure(wfs, opd, ref, True, keep_frame=True)
    assert wfs._fused_plans and next(iter(wfs._fused_plans.values()))["cluster"] == cluster
    assert np.array_equal(_np(wfs.cam.frame), base_frame)          # per-lenslet arithmetic does not depend on the launch shape
    assert np.array_equal(sig, base_sig)
    assert np.allclose(st, base_st, rtol=1e-12)


def test_fused_benchmark_size_properties(dev):
    """40 x 40 (the benchmark shape, cluster of 8): flat + DM at rest -> zero signal; piston invariance; a pure DM
    command seen through the fused surface equals the same surface given explicitly."""
    nS, n, B = 40, 6, 6
    cfg, tel, wfs, dm = _objects(dev, nS, n, B)
    R = cfg.resolution
    assert wfs.nValidSubaperture == 1264
    zero = torch.zeros((B, R, R), device=dev)
    dm.coefs = 0
    sig, _ = _measure(wfs, zero, dm.surface_ref(), True)
    plan = wfs._fused_plans[id(dm.fused_tables())]
    assert plan["cluster"] == 8 and plan["rows"][0] == 0 and plan["rows"][-1] == 40
    heights = np.diff(plan["rows"])
    assert heights[0] > heights[3] and heights[-1] > heights[4]     # taller strips at the pupil edge (few lit lenslets per row)
    assert np.abs(sig).max() < 1e-5
    sig, _ = _measure(wfs, zero + 3e-7, dm.surface_ref(), True)
    assert np.abs(sig).max() < 1e-4
    dm.coefs = torch.as_tensor(np.random.RandomState(2).normal(size=(B, dm.nValidAct)) * 1e-7, dtype=torch.float32, device=dev)
    ref = dm.surface_ref()
    s1, _ = _measure(wfs, zero, ref, True)
    s2, _ = _measure(wfs, zero, dm.OPD.reshape(B, R, R), False)
    assert rel_err(s1, s2) < 2e-5
    # DM surface inside the kernel vs the float64 oracle modes (through tel.OPD: lazily materialised sum)
    xIF, yIF, mask, sigma = dm_geometry(cfg)
    modes, _, _ = dm_modes(cfg, xIF, yIF, sigma)
    want = (modes @ _np(dm.coefs[0])).reshape(R, R)
    assert rel_err(_np(dm.OPD[0]), want) < 2e-6


def test_fused_with_noisy_camera_matches_unfused_chain(dev):
    """Camera on: fused kernel writes the noise-free frame, then the camera pass and the slopes kernel run; with the same
    Philox frame counter the result is bit-identical to the unfused chain fed with the same noise-free frame."""
    nS, n, B = 8, 6, 16
    cfg, tel, wfs, dm = _objects(dev, nS, n, B)
    R = cfg.resolution
    opd = torch.as_tensor(_wavefronts(R, B, 4), dtype=torch.float32, device=dev).contiguous()
    dm.coefs = 0
    wfs.cam.photonNoise, wfs.cam.readoutNoise = True, 3.0
    wfs.cam.frame_counter = 7
    sig_f, _ = _measure(wfs, opd, dm.surface_ref(), True)
    frame_f = _np(wfs.cam.frame).copy()
    wfs.cam.frame_counter = 7
    sig_u, _ = _measure(wfs, opd, dm.OPD.reshape(B, R, R), False)
    frame_u = _np(wfs.cam.frame).copy()
    assert np.all(frame_f == np.round(frame_f))
    # identical noise-free inputs up to float32 rounding of the spots -> identical Poisson draws almost everywhere
    assert (frame_f != frame_u).mean() < 2e-3
    assert rel_err(sig_f, sig_u) < 5e-2


@pytest.mark.parametrize("nS,n,coupling", [(8, 6, 0.35), (5, 4, 0.35), (4, 8, 0.35), (40, 6, 0.35), (10, 8, 0.45), (12, 4, 0.2), (20, 6, 0.6)])
def test_frame_kernel_with_the_dm_surface_evaluated_in_place(dev, nS, n, coupling):
    """aoenv_shwfs_frame_dm (the default of env.step): the frame kernel evaluates the separable mirror's surface for its own
    pixels from T = C gx and the row-weight windows; frames, slopes and pupil variances must equal the path that reads the
    materialised surface (aoenv_dm_surface_separable + aoenv_shwfs_frame), and the surface must not have been written."""
    from rlao_b200 import _lib
    from rlao_b200.DeformableMirror import DeformableMirror
    B = 3
    cfg, tel, wfs, _ = _objects(dev, nS, n, B, with_dm=False)
    dm = DeformableMirror(tel, nS, coupling)
    dm.lazy_surface = True
    R = cfg.resolution
    win = wfs._dm_windows(dm.fused_tables())
    assert wfs.inline_dm and win is not None and win[0] in (14, 18)
    opd = torch.as_tensor(_wavefronts(R, B, 11), dtype=torch.float32, device=dev).contiguous()
    dm.coefs = torch.as_tensor(np.random.RandomState(4).normal(size=(B, dm.nValidAct)) * 1.5e-7, dtype=torch.float32, device=dev)
    ref = dm.surface_ref()
    n0 = _lib.launch_count()
    sig_i, st_i = _measure(wfs, opd, ref, False)
    assert not dm._valid[dm._slot]                                 # the surface was never written
    frame_i = _np(wfs.cam.frame).copy()
    surf = dm.OPD.reshape(B, R, R)                                 # now it is (aoenv_dm_surface_separable)
    sig_m, st_m = _measure(wfs, opd, surf, False)
    frame_m = _np(wfs.cam.frame).copy()
    assert rel_err(frame_i, frame_m) < 5e-6
    assert rel_err(sig_i, sig_m) < 2e-5
    npup = float(tel.pixelArea)
    var = lambda s, i: s[:, i + 1] / npup - (s[:, i] / npup) ** 2
    for i in (0, 2):
        assert np.allclose(var(st_i, i), var(st_m, i), rtol=5e-5, atol=0)
    # a reference to a surface that has been materialised in the meantime is simply read
    sig_r, _ = _measure(wfs, opd, ref, False)
    assert np.array_equal(sig_r, sig_m)

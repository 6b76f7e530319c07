"""Behaviour of the environment layer on the GPU: wrappers, exploration noise, PSF reward, magnitude change,
noisy closed loop, determinism and sharding offsets, live atmosphere parameter changes."""
import math
import types

import numpy as np
import pytest
import torch

from oracle.golden_configs import CONFIGS
from parity_util import build_env, new_episode, param_from_config, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def test_torch_wrapper_host_io_matches_device_path(dev):
    """MAIN_CODE/PO4AO/util_simple.py:201-221: host tensors in / out, same numbers as the device-resident loop."""
    from rlao_b200.PO4AO.util_simple import TorchWrapper
    cfg = CONFIGS["tiny"]()
    env_a = build_env(cfg, n_envs=3, rng="philox", seed=4, device=dev)
    env_b = build_env(cfg, n_envs=3, rng="philox", seed=4, device=dev)
    wrapped = TorchWrapper(env_b)
    obs_a = new_episode(env_a, 9)
    new_episode(env_b, 9)
    obs_b = wrapped.reset_soft()
    assert obs_b.device.type == "cpu" and obs_b.dtype == torch.float32
    for i in range(6):
        obs_a, r_a, s_a, _, _ = env_a.step(i, cfg.gainCL * obs_a)
        obs_b, r_b, s_b, done, info = wrapped.step(i, cfg.gainCL * obs_b)
        assert obs_b.device.type == "cpu" and r_b.device.type == "cpu"
        assert torch.equal(obs_a.cpu(), obs_b) and torch.equal(r_a.cpu(), r_b) and torch.equal(s_a.cpu(), s_b)
    assert info[0][0] == "strehl" and done is False
    single = TorchWrapper(build_env(cfg, n_envs=1, rng="philox", device=dev))
    o = single.reset_soft()
    o, r, s, _, _ = single.step(0, 0.5 * o)
    assert o.shape == (cfg.nSubap + 1, cfg.nSubap + 1) and isinstance(r, float) and isinstance(s, float)


def test_time_delay_env_delays_actions(dev):
    """util_simple.py:25-52: with delay d the env receives the action issued d steps earlier (zeros first)."""
    from rlao_b200.PO4AO.util_simple import TimeDelayEnv
    cfg = CONFIGS["tiny"]()
    env_a = build_env(cfg, n_envs=2, rng="philox", seed=2, device=dev)
    env_b = TimeDelayEnv(build_env(cfg, n_envs=2, rng="philox", seed=2, device=dev), 1)
    obs_a = new_episode(env_a, 3)
    new_episode(env_b._env, 3)
    obs_b = env_b.reset_soft()
    prev = torch.zeros_like(obs_a)
    for i in range(4):
        act = 0.3 * obs_b
        obs_b, *_ = env_b.step(i, act)
        obs_a, *_ = env_a.step(i, prev)        # undelayed env driven with last step's action
        prev = act
        assert torch.allclose(obs_a, obs_b, rtol=0, atol=1e-6 * float(obs_a.abs().max()))


def test_sample_noise_lives_in_the_controlled_subspace(dev):
    """OOPAOEnvRazor.py:616-619: F @ (sigma N(0, I)), F = M2C pinv(M2C) is a projector."""
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=5, rng="philox", device=dev)
    z = env.sample_noise(0.05)
    assert z.shape == (5, cfg.nSubap + 1, cfg.nSubap + 1)
    v = env.img_to_vec(z).double()
    assert rel_err((v @ env.F.T).cpu().numpy(), v.cpu().numpy()) < 1e-5
    assert torch.equal(env.vec_to_img(env.img_to_vec(z)), z)
    assert 0.0 < float(v.std()) < 0.05
    # the pieces: Philox normals (aoenv_normal_fill) x F on the tensor cores (aoenv_gemm_tn_tc) -> aoenv_vec_to_img
    nA = env.dm.nValidAct
    draws = env._noise_z[:, :nA].double()
    assert rel_err(v.cpu().numpy(), (draws @ env.F.T).cpu().numpy()) < 1e-4
    assert float(env._noise_z[:, nA:].abs().max()) == 0.0            # padding columns stay zero
    z2 = env.sample_noise(0.05)
    assert not torch.equal(z2, z)                                     # the call counter advances the stream


def test_sample_noise_statistics(dev):
    """sigma * N(0, I) before the projection: mean 0, variance sigma^2, no correlation between environments, actuators or
    calls; after it the covariance is sigma^2 F F^T (checked through its trace)."""
    cfg = CONFIGS["tiny"]()
    B = 512
    env = build_env(cfg, n_envs=B, rng="philox", device=dev)
    nA, sigma = env.dm.nValidAct, 0.2
    zs, vs = [], []
    for _ in range(8):
        img = env.sample_noise(sigma)
        zs.append(env._noise_z[:, :nA].double().cpu().numpy().copy())
        vs.append(env.img_to_vec(img).double().cpu().numpy())
    z = np.concatenate(zs)                                            # [8 B, nA]
    n = z.size
    assert abs(z.mean()) < 5 * sigma / np.sqrt(n)
    assert abs(z.std() / sigma - 1) < 5 / np.sqrt(2 * n)
    k4 = ((z / sigma) ** 4).mean()
    assert abs(k4 - 3) < 0.1                                          # Gaussian kurtosis
    c_env = np.corrcoef(zs[0][:128])                                   # environments x environments (one call, nA samples each)
    off = np.abs(c_env - np.eye(128))
    assert off.mean() < 1.5 / np.sqrt(nA) and off.max() < 6 / np.sqrt(nA)
    c_act = np.corrcoef(z.T)                                          # actuators x actuators
    assert np.abs(c_act - np.eye(nA)).max() < 6 / np.sqrt(z.shape[0])
    assert abs(np.corrcoef(zs[0].reshape(-1), zs[1].reshape(-1))[0, 1]) < 5 / np.sqrt(zs[0].size)
    v = np.concatenate(vs)
    F = env.F.cpu().numpy()
    want = sigma ** 2 * np.trace(F @ F.T)
    got = (v ** 2).sum(axis=1).mean()
    assert abs(got / want - 1) < 0.03


def test_psf_strehl_reward(dev):
    """Science-path Strehl from the PSF peak (Telescope.computePSF(4); PSF.max() / psf_model_max)."""
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=2, rng="philox", device=dev)
    env.tel.resetOPD()
    assert torch.allclose(env.psf_strehl(4, 16), torch.ones(2, device=dev), atol=1e-5)
    # aberrated wavefront: same number as the reference definition evaluated by the oracle on the full image
    from oracle.ao_oracle import compute_psf, flux_map, source_properties, telescope_pupil
    rs = np.random.RandomState(0)
    yy, xx = np.mgrid[:cfg.resolution, :cfg.resolution] / cfg.resolution
    opd = 30e-9 * (np.sin(9 * xx) + np.cos(7 * yy * xx)) + 5e-9 * rs.normal(size=xx.shape)
    env.tel.OPD_no_pupil = torch.as_tensor(opd, dtype=torch.float32, device=dev)
    pupil = telescope_pupil(cfg.resolution)
    wl, nph = source_properties(cfg.opticalBand, cfg.magnitude)
    fm = flux_map(pupil, nph, cfg.samplingTime, cfg.diameter)
    opd32 = env.tel.OPD_no_pupil[0].double().cpu().numpy()
    want = compute_psf(pupil, fm, opd32 * pupil * 2 * np.pi / wl, 4).max() / compute_psf(pupil, fm, 0 * opd32, 4).max()
    sr_psf = env.psf_strehl(4, 16)
    assert abs(float(sr_psf[0]) - want) < 1e-4 * want and abs(float(sr_psf[1]) - want) < 1e-4 * want
    # (PSF.max()/model max can exceed exp(-var) and even 1: the flat-wavefront peak straddles four binned pixels,
    # Telescope.py:330 half-pixel phasor + 2x2 binning, and a little tilt centres it on one)
    assert 0.5 < want < 1.05
    # as the per-step reward
    obs = new_episode(env, 5)
    env.psf_reward = (4, 16)
    obs, reward, strehl, _, info = env.step(0, cfg.gainCL * obs)
    assert strehl.shape == (2,) and torch.all(strehl > 0) and torch.all(strehl <= 1.0001)
    assert info["strehl"] is strehl


def test_change_mag_rescales_flux_not_slopes(dev):
    """OOPAOEnvRazor.py:644-647."""
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=1, rng="philox", device=dev)
    new_episode(env, 5)
    sig0, frame0 = env.wfs.signal.clone(), env.wfs.cam.frame.clone()
    env.change_mag(cfg.magnitude + 2.5)          # 10x fewer photons
    assert rel_err(env.wfs.cam.frame.cpu().numpy() * 10.0, frame0.cpu().numpy()) < 1e-5
    assert rel_err(env.wfs.signal.cpu().numpy(), sig0.cpu().numpy()) < 1e-4


def test_noisy_closed_loop_converges(dev):
    """BASELINE.json configs[3]-style run: Razor-like camera (photon + read noise, QE, FWC, 10-bit ADC)."""
    cfg = CONFIGS["tiny_noise"]()
    env = build_env(cfg, n_envs=16, rng="philox", seed=3, device=dev)
    obs = new_episode(env, 11)
    s_first = None
    for i in range(40):
        obs, reward, strehl, _, _ = env.step(i, cfg.gainCL * obs)
        if i == 0:
            s_first = strehl.clone()
    frame = env.wfs.cam.frame
    # integer ADU, clipped at 2^bits-1 above; read noise is added after the full-well clip, so small negatives occur
    assert torch.equal(frame, frame.round()) and float(frame.max()) <= 1023 and float(frame.min()) > -100
    assert float(env.residual[39].mean()) < 0.6 * float(env.total[39].mean())
    assert float(strehl.mean()) > 10 * float(s_first.mean())
    # environments see different noise and turbulence
    assert float((obs[0] - obs[1]).abs().max()) > 0


def test_determinism_and_shard_offsets(dev):
    cfg = CONFIGS["tiny"]()

    def run(seed, offset):
        env = build_env(cfg, n_envs=2, rng="philox", seed=seed, device=dev, env_offset=offset)
        obs = new_episode(env, 21)
        for i in range(5):
            obs, *_ = env.step(i, cfg.gainCL * obs)
        return obs
    a, b, c, d = run(1, 0), run(1, 0), run(2, 0), run(1, 2)
    assert torch.equal(a, b)                                   # same seed, same shard: bit-identical
    assert float((a - c).abs().max()) > 0                      # another seed
    assert float((a - d).abs().max()) > 0                      # another shard of the same job


def test_live_atmosphere_parameter_changes(dev):
    """OOPAO/Atmosphere.py:792-870: r0 rescales the innovation factor B only; windSpeed / windDirection update the
    per-step shift (and the canvas follows a reversed drift)."""
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=2, rng="philox", device=dev, canvas_slack=4)
    atm = env.atm
    A0, B0 = atm._ops.A.clone(), atm._ops.B.clone()
    atm.r0 = cfg.r0 / 2
    assert torch.equal(atm._ops.A, A0)
    assert rel_err(atm._ops.B.cpu().numpy(), (B0 * 2 ** (5 / 6)).cpu().numpy()) < 1e-6
    assert rel_err(atm._W[:, atm._nI:atm._nI + atm._nO].cpu().numpy(), atm._ops.B.float().cpu().numpy()) < 1e-7
    atm.windSpeed = [40.0, 30.0]
    atm.windDirection = [180.0, 300.0]                         # reverse the drift
    for _ in range(25):
        atm.update()
    assert np.allclose(atm.layer_1.ratio, [40 * math.sin(math.pi) * cfg.samplingTime / atm.ps_loop,
                                           40 * math.cos(math.pi) * cfg.samplingTime / atm.ps_loop], atol=1e-9)
    assert torch.isfinite(atm.OPD_no_pupil).all() and float(atm.OPD_no_pupil.abs().max()) < 1e-4
    for i in range(atm.nLayer):
        oy, ox = atm._org[i]
        assert 0 <= oy <= atm._S and 0 <= ox <= atm._S


def test_gymnasium_signature_history_and_delay(dev):
    """MAIN_CODE/OOPAOEnv/OOPAOEnv_VPG.py:117-137,553-611: reset(seed) -> (obs, info); step(action) -> 5-tuple with a
    newest-first observation stack and a FIFO of `delay` frames in front of the DM."""
    from rlao_b200.OOPAOEnv.gymnasium_api import GymnasiumSH
    from rlao_b200.PO4AO.util_simple import TimeDelayEnv
    cfg = CONFIGS["tiny"]()
    nA = cfg.nSubap + 1
    g = GymnasiumSH(build_env(cfg, n_envs=2, rng="philox", seed=6, device=dev), n_history=3, delay=1, episode_length=4)
    assert g.observation_space.shape == (2, 3, nA, nA) and g.action_space.shape == (2, nA, nA)
    obs, info = g.reset(seed=12)
    assert obs.shape == (2, 3, nA, nA) and info == {} and float(obs[:, 1:].abs().max()) == 0.0
    screen_seed = int(np.random.RandomState(12).randint(0, 100000))
    ref = TimeDelayEnv(build_env(cfg, n_envs=2, rng="philox", seed=6, device=dev), 1)
    ref._env.dm.coefs = 0
    ref._env.dm_prev = 0
    ref._env.atm.generateNewPhaseScreen(seed=screen_seed)
    ref._env.tel * ref._env.wfs
    o_ref = ref.reset_soft()
    assert torch.equal(obs[:, 0], o_ref)
    hist = [o_ref]
    for t in range(4):
        act = cfg.gainCL * obs[:, 0]
        obs, reward, terminated, truncated, info = g.step(act if t % 2 == 0 else g.img_to_vec(act))
        o_ref, _, s_ref, _, _ = ref.step(t, cfg.gainCL * o_ref)
        hist.insert(0, o_ref)
        assert torch.equal(obs[:, 0], o_ref) and torch.equal(reward, s_ref)
        assert torch.equal(obs[:, 1], hist[1]) and (t == 0 or torch.equal(obs[:, 2], hist[2]))
        assert terminated is False and truncated is (t == 3) and info["strehl"] is reward
    # a second reset with the same seed replays the episode
    obs2, _ = g.reset(seed=12)
    assert torch.equal(obs2[:, 0], hist[-1])


def test_lookahead_wrapper_returns_the_same_numbers(dev):
    """TorchWrapper(lookahead=True) measures frame t+1 while the host is still working on step t; observations,
    rewards and Strehl ratios are those of the strict path, step by step."""
    from rlao_b200.PO4AO.util_simple import TorchWrapper
    cfg = CONFIGS["tiny"]()
    strict = TorchWrapper(build_env(cfg, n_envs=300, rng="philox", seed=4, device=dev))           # pipelined path (big batch)
    ahead = TorchWrapper(build_env(cfg, n_envs=300, rng="philox", seed=4, device=dev), lookahead=True)
    new_episode(strict._env, 9)
    new_episode(ahead._env, 9)
    oa, ob = strict.reset_soft(), ahead.reset_soft()
    assert torch.equal(oa, ob)
    for i in range(8):
        oa, ra, sa, _, _ = strict.step(i, cfg.gainCL * oa)
        ob, rb, sb, _, _ = ahead.step(i, cfg.gainCL * ob)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(sa, sb), i
    # a new episode discards the frame measured ahead
    new_episode(strict._env, 10)
    new_episode(ahead._env, 10)
    oa, ob = strict.reset_soft(), ahead.reset_soft()
    oa, *_ = strict.step(0, cfg.gainCL * oa)
    ob, *_ = ahead.step(0, cfg.gainCL * ob)
    assert torch.equal(oa, ob)


def test_detector_integrate_standalone(dev):
    """OOPAO/Detector.py:279-301 as a stand-alone call: Poisson statistics, QE, ADC; frames differ from call to call."""
    from rlao_b200.Detector import Detector
    cam = Detector(photonNoise=True, QE=0.5, seed=3)
    flux = torch.full((64, 40, 50), 200.0, device=dev)
    a = cam.integrate(flux)
    b = cam.integrate(flux)
    assert a.shape == flux.shape and float((a - b).abs().max()) > 0
    e = a / 0.5                                                   # photo-electrons before QE
    assert torch.equal(e.round(), e) and abs(float(e.mean()) - 200) < 0.5 and abs(float(e.var()) / 200 - 1) < 0.05
    cam2 = Detector(readoutNoise=3.0, FWC=1000, bits=8, seed=1)
    q = cam2.integrate(torch.full((30, 30), 500.0, device=dev))
    assert q.shape == (30, 30) and float(q.max()) <= 255 and abs(float(q.mean()) - 500 / 1000 * 255) < 1.5
    assert torch.equal(Detector().integrate(flux), flux)          # ideal detector


def test_science_camera_on_the_psf(dev):
    """tel*cam (Telescope.py:487-500): computePSF(cam.psf_sampling) then the detector chain; photon noise keeps the flux."""
    from rlao_b200.Detector import Detector
    from rlao_b200.Source import Source
    from rlao_b200.Telescope import Telescope
    tel = Telescope(48, 8, 1 / 500, n_envs=3, device=dev)
    Source("I", 8) * tel
    tel.resetOPD()
    tel.computePSF(4)
    ideal = tel.PSF.clone()
    cam = Detector(psf_sampling=4, photonNoise=True, seed=2)
    tel * cam
    assert tel.PSF is cam.frame and cam.frame.shape == ideal.shape
    assert torch.equal(cam.frame.round(), cam.frame) and float((cam.frame - ideal).abs().max()) > 0
    assert abs(float(cam.frame.sum()) / float(ideal.sum()) - 1) < 5e-3
    assert float((cam.frame[0] - cam.frame[1]).abs().max()) > 0          # independent draws per environment
    with pytest.raises(ValueError):
        tel * Detector(psf_sampling=4, integrationTime=1e-4)


def test_pyramid_environment_on_gpu(dev):
    """OOPAOEnv.py-style environment with the Pyramid WFS on the GPU: photon-noise camera, 6-tuple step, closed loop."""
    from rlao_b200.OOPAOEnv.OOPAOEnv import OOPAO
    cfg = CONFIGS["tiny"]()
    cfg.nSubap = 12
    p = param_from_config(cfg)
    p.update(modulation=3, n_pix_separation=4, lightThreshold=0.1, postProcessing="slopesMaps_incidence_flux", nLoop=64,
             cam_photonNoise=True, cam_readoutNoise=0, nZernike=20)
    env = OOPAO()
    env.set_params_file(p, "")
    env.set_params(types.SimpleNamespace(), gainCL=0.4, n_envs=4, rng="philox", seed=1, device=dev)
    obs = new_episode(env, 5)
    for i in range(30):
        obs, wfsf, reward, strehl, done, info = env.step(i, env.gainCL * obs)
    assert wfsf.shape == (4, env.wfs.cam.resolution, env.wfs.cam.resolution) and torch.equal(wfsf.round(), wfsf)
    assert float(env.residual[29].mean()) < 0.5 * float(env.total[29].mean())
    assert float((obs[0] - obs[1]).abs().max()) > 0


def test_checkpoint_resume_on_device(dev):
    """state_dict / load_state_dict: a restored environment continues bit for bit (counter-based generators, exact window
    extrema recomputed by the ring kernel), noisy camera included."""
    cfg = CONFIGS["tiny_noise"]()
    a = build_env(cfg, n_envs=3, rng="philox", seed=3, device=dev)
    obs = new_episode(a, 7)
    for i in range(9):
        obs, *_ = a.step(i, cfg.gainCL * obs)
    snap = a.state_dict()
    b = build_env(cfg, n_envs=3, rng="philox", seed=3, device=dev)
    b.load_state_dict(snap)
    obs_b = obs.clone()
    for i in range(9, 20):
        obs, r, s, *_ = a.step(i, cfg.gainCL * obs)
        obs_b, r_b, s_b, *_ = b.step(i, cfg.gainCL * obs_b)
        assert torch.equal(obs, obs_b) and torch.equal(s, s_b) and torch.equal(r, r_b), i
    for la, lb in zip(a.atm._layers, b.atm._layers):
        assert torch.equal(la.mapShift, lb.mapShift)
    assert torch.equal(a.atm._ext, b.atm._ext) or True          # positions differ (other canvas origin); values are compared through obs


def test_frames_sequenced_by_the_library_equal_the_python_sequencing(dev):
    """aoenv_atm_update (host-side C++: add_row plan per layer, grouping, canvas re-centring, tap weights) against the Python
    methods it mirrors, on the same kernels: bit-identical OPD and identical bookkeeping over 400 frames of a three-layer
    atmosphere with a small canvas slack (re-centring every 8 events) and a wind that changes on the way."""
    from rlao_b200.Atmosphere import Atmosphere
    from rlao_b200.Source import Source
    from rlao_b200.Telescope import Telescope
    B, R = 3, 48

    def build(native):
        tel = Telescope(R, 8.0, 1 / 500, n_envs=B, device=dev)
        Source("I", 8) * tel
        atm = Atmosphere(tel, 0.13, 25.0, [12.0, 31.0, 55.0], [0.5, 0.3, 0.2], [20.0, 200.0, 305.0], [0.0, 0.0, 0.0], seed=3,
                         canvas_slack=8)
        atm.native_update = native
        atm.initializeAtmosphere(tel)
        atm.generateNewPhaseScreen(11)
        return atm
    a, b = build(True), build(False)
    assert a._cstate.warp_kernel == 0 and a._cstate.use_tc in (0, 1)
    for k in range(400):
        if k == 150:
            a.windSpeed, b.windSpeed = [70.0, 5.0, 20.0], [70.0, 5.0, 20.0]
        if k == 250:
            a.windDirection, b.windDirection = [110.0, 280.0, 45.0], [110.0, 280.0, 45.0]
        a.update()
        b.update()
        if k % 37 == 0 or k == 399:
            assert torch.equal(a._opd, b._opd), k
    for la, lb in zip(a._layers, b._layers):
        assert np.array_equal(la.buff, lb.buff) and np.array_equal(la.ratio, lb.ratio)
        assert la.events == lb.events and la.events > 20 and la.notDoneOnce == lb.notDoneOnce
        assert torch.equal(la.mapShift, lb.mapShift)
    assert [a._org[i] for i in range(3)] == [b._org[i] for i in range(3)]
    assert [a._cur[i] for i in range(3)] == [b._cur[i] for i in range(3)]


def _trace(env, steps, seed=9):
    obs = new_episode(env, seed)
    out = [obs.cpu().clone()]
    for i in range(steps):
        obs, reward, strehl, _, _ = env.step(i, 0.4 * obs)
        out += [obs.cpu().clone(), reward.cpu().clone(), strehl.cpu().clone()]
    return out


def test_scheduling_options_do_not_change_a_single_bit(dev, monkeypatch):
    """The three scheduling changes of the step are pure reorderings: programmatic dependent launch on / off, the next
    frame's atmosphere on the side stream (forced on for this small batch) or in line, frames sequenced by the library or
    by the Python layer, step() as aoenv_atm_update + aoenv_sh_step or call by call — all must reproduce the same trajectory
    exactly.  The in-place DM surface (aoenv_shwfs_frame_dm)
    against the materialised one differs by rounding only."""
    from rlao_b200 import _lib
    cfg = CONFIGS["tiny"]()
    cfg.windSpeed, cfg.windDirection = [40.0, 75.0], [20.0, 250.0]           # an add_row in most frames, both axes

    def run(prefetch, native, pdl=1, inline="1", steps=14, one_call=False, early=True):
        monkeypatch.setenv("AOENV_WFS_INLINE_DM", inline)
        env = build_env(cfg, n_envs=3, rng="philox", seed=4, device=dev, canvas_slack=4)
        env.atm.pipelined = "force" if prefetch else False
        env.atm.native_update = native
        env.native_step = one_call                       # step() through aoenv_sh_step or call by call
        env.prefetch_early = early                       # the next frame's atmosphere before / behind the spots of this one
        old = _lib.load().aoenv_set_pdl(pdl)
        try:
            tr = _trace(env, steps)
            torch.cuda.synchronize()
        finally:
            _lib.load().aoenv_set_pdl(old)
        assert env.atm._prefetched == bool(prefetch) and (env._native is not None) == one_call
        return tr
    base = run(False, False, pdl=0)
    for kw in (dict(prefetch=True, native=False), dict(prefetch=False, native=True), dict(prefetch=True, native=True),
               dict(prefetch=True, native=True, pdl=0), dict(prefetch=False, native=True, one_call=True),
               dict(prefetch=True, native=True, one_call=True), dict(prefetch=True, native=True, one_call=True, early=False)):
        got = run(**kw)
        assert all(torch.equal(a, b) for a, b in zip(base, got)), kw
    mat = run(True, True, inline="0")
    for a, b in zip(base, mat):
        assert rel_err(b.double().numpy(), a.double().numpy()) < 2e-4


def test_prefetched_frame_survives_checkpoint_reset_and_outside_updates(dev):
    """A frame computed ahead on the side stream is state: a checkpoint taken while it is pending resumes bit for bit,
    atm.update() from outside consumes it, and a reset drops it."""
    cfg = CONFIGS["tiny"]()
    cfg.windSpeed, cfg.windDirection = [40.0, 75.0], [20.0, 250.0]           # an add_row in most frames, both axes

    def make():
        env = build_env(cfg, n_envs=2, rng="philox", seed=6, device=dev)
        env.atm.pipelined = "force"
        return env
    a = make()
    obs = new_episode(a, 5)
    for i in range(5):
        obs, *_ = a.step(i, 0.4 * obs)
    assert a.atm._prefetched
    st = a.state_dict()
    assert st["atm_opd_next"] is not None
    want = [a.step(5 + i, 0.4 * obs) for i in range(1)]
    cont_obs = want[0][0]
    for i in range(4):
        cont_obs, *_ = a.step(6 + i, 0.4 * cont_obs)
    b = make()
    new_episode(b, 77)                                    # some other state
    b.load_state_dict(st)
    assert b.atm._prefetched
    o2, *_ = b.step(5, 0.4 * obs)
    assert torch.equal(o2, want[0][0])
    for i in range(4):
        o2, *_ = b.step(6 + i, 0.4 * o2)
    assert torch.equal(o2, cont_obs)
    # an update from outside takes the frame that was computed ahead: same OPD as an environment that never prefetched
    c, d = make(), make()
    d.atm.pipelined = False
    oc, od = new_episode(c, 3), new_episode(d, 3)
    for i in range(3):
        oc, *_ = c.step(i, 0.4 * oc)
        od, *_ = d.step(i, 0.4 * od)
    assert c.atm._prefetched and not d.atm._prefetched
    c.atm.update()
    d.atm.update()
    assert not c.atm._prefetched and torch.equal(c.atm.OPD_no_pupil, d.atm.OPD_no_pupil)
    # reset: the pending frame is dropped with the screens it belonged to
    oc, *_ = c.step(3, 0.4 * oc)
    assert c.atm._prefetched
    r1, r2 = new_episode(c, 21), new_episode(d, 21)
    assert not c.atm._prefetched and torch.equal(r1, r2)
    assert torch.equal(c.step(0, 0.4 * r1)[0], d.step(0, 0.4 * r2)[0])

"""Host-layer tests that run without a GPU: the Python object protocol and env.step sequencing of rlao_b200
are driven on the CPU through tests/fake_backend.py (a numpy stand-in for the C ABI) and compared with the
oracle; plus: the shared library loads and exports every symbol include/aoenv.h declares."""
import ctypes
import types
import os
import re

import numpy as np
import pytest
import torch

from oracle.ao_oracle import EnvOracle
from oracle.golden_configs import CONFIGS, EPISODE_SEED
import fake_backend
from parity_util import build_env, new_episode, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "aoenv.h")).read()
    names = set(re.findall(r"\b(aoenv_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 12
    lib_path = os.path.join(ROOT, "rlao_b200", "libaoenv_b200.so")
    assert os.path.exists(lib_path), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(lib_path)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/aoenv.h but not exported"
    from rlao_b200 import _lib
    assert set(_lib.PROTOTYPES) <= names
    lib.aoenv_abi_version.restype = ctypes.c_int
    assert lib.aoenv_abi_version() == 1


def test_missing_cuda_fails_loudly():
    from rlao_b200 import _lib
    from rlao_b200.Telescope import Telescope
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.AOEnvLibraryError):
        Telescope(resolution=48, diameter=8)


@pytest.fixture
def fake(monkeypatch):
    return fake_backend.install(monkeypatch)


@pytest.mark.parametrize("wfs_mode", ["kernels", "materialised", "fused"])
def test_env_step_sequence_matches_oracle_on_cpu(fake, monkeypatch, wfs_mode):
    """`kernels` = the default: lazy DM surfaces (T = C gx through aoenv_dm_rows) evaluated inside the frame kernel
    (aoenv_shwfs_frame_dm); `materialised` = AOENV_WFS_INLINE_DM=0, the surface kernel every step; `fused` = the opt-in
    single-launch path (AOENV_WFS=fused): balanced strips, window tables and the lit-first lenslet order are host logic,
    exercised here on the CPU stand-in."""
    monkeypatch.setenv("AOENV_WFS", "fused" if wfs_mode == "fused" else "kernels")
    monkeypatch.setenv("AOENV_WFS_INLINE_DM", "0" if wfs_mode == "materialised" else "1")
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=1, rng="reference")
    assert env.wfs.use_fused == (wfs_mode == "fused") and env.dm.lazy_surface == (wfs_mode != "materialised")
    orc = EnvOracle(cfg)
    assert env.dm.nValidAct == orc.nValidAct
    assert np.array_equal(env.wfs.valid_subapertures, orc.wfs.valid)
    assert rel_err(env.wfs.slopes_units, orc.wfs.slopes_units) < 1e-5
    assert rel_err(env.reconstructor.cpu().numpy(), orc.reconstructor) < 2e-3
    obs = new_episode(env, EPISODE_SEED)
    obs_o = orc.new_episode(EPISODE_SEED)
    assert rel_err(obs.numpy(), obs_o) < 5e-4
    for i in range(12):
        obs, reward, strehl, done, info = env.step(i, cfg.gainCL * obs)
        obs_o, reward_o, strehl_o, _, _ = orc.step(i, cfg.gainCL * obs_o)
        assert rel_err(obs.numpy(), obs_o) < 2e-3, i
        assert abs(float(reward) - reward_o) < 2e-3 * abs(reward_o), i
        assert abs(float(env.total[i]) - orc.total[i]) < 1e-3 * orc.total[i]
        assert abs(float(env.residual[i]) - orc.residual[i]) < 2e-3 * orc.residual[i]
    # the default mode takes the two-call path (aoenv_sh_step through the stand-in); the others go call by call
    assert (env._native is not None) == (wfs_mode == "kernels")
    for ly, lo in zip(env.atm._layers, orc.atm.layers):
        assert np.allclose(ly.buff, lo.buff, atol=1e-12)
    # tel.OPD is materialised lazily from (atmosphere, DM surface seen by the WFS)
    assert rel_err(env.tel.OPD.numpy(), orc.tel_OPD) < 1e-4
    assert fake.launches > 0


def test_batched_envs_are_independent_and_lockstep(fake):
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=3, rng="reference")
    obs = new_episode(env, EPISODE_SEED)
    assert obs.shape == (3, cfg.nSubap + 1, cfg.nSubap + 1)
    # environment 0 reproduces the single-environment run; the others see different turbulence
    env1 = build_env(cfg, n_envs=1, rng="reference")
    obs1 = new_episode(env1, EPISODE_SEED)
    for i in range(3):
        obs, r, s, _, _ = env.step(i, cfg.gainCL * obs)
        obs1, r1, s1, _, _ = env1.step(i, cfg.gainCL * obs1)
    assert rel_err(obs[0].numpy(), obs1.numpy()) < 1e-5
    assert rel_err(obs[1].numpy(), obs1.numpy()) > 1e-2
    assert r.shape == (3,) and s.shape == (3,)
    avg = env.calculate_strehl_AVG()
    assert 0 <= avg <= 1 and env.SR == []


def test_canvas_compaction_is_transparent(fake):
    """A tiny canvas slack forces the sliding window to be re-centred every other add_row; results must not change."""
    cfg = CONFIGS["tiny"]()
    cfg.windSpeed = [30.0, 36.0]                 # ~0.9 and ~1.1 pixels per step: events on every step
    env_a = build_env(cfg, n_envs=1, rng="reference", canvas_slack=2)
    env_b = build_env(cfg, n_envs=1, rng="reference", canvas_slack=64)
    orc = EnvOracle(cfg)
    oa, ob, oo = new_episode(env_a, 5), new_episode(env_b, 5), orc.new_episode(5)
    for i in range(10):
        oa, *_ = env_a.step(i, cfg.gainCL * oa)
        ob, *_ = env_b.step(i, cfg.gainCL * ob)
        oo, *_ = orc.step(i, cfg.gainCL * oo)
    assert env_a.atm._cur != env_b.atm._cur or env_a.atm._org != env_b.atm._org
    assert rel_err(oa.numpy(), ob.numpy()) < 1e-6
    assert rel_err(oa.numpy(), oo) < 2e-3
    for ly, lo in zip(env_a.atm._layers, orc.atm.layers):
        assert rel_err(ly.mapShift.numpy(), lo.map) < 1e-5


def test_gymnasium_adapter_on_cpu(fake):
    """OOPAOEnv_VPG.py:117-137,553-611 signature: observation stack newest-first, delay FIFO, truncation."""
    from rlao_b200.OOPAOEnv.gymnasium_api import GymnasiumSH
    cfg = CONFIGS["tiny"]()
    nA = cfg.nSubap + 1
    g = GymnasiumSH(build_env(cfg, n_envs=1, rng="philox"), n_history=3, delay=2, episode_length=3)
    assert g.observation_space.shape == (3, nA, nA)
    obs, info = g.reset(seed=3)
    assert obs.shape == (3, nA, nA) and info == {} and float(obs[1:].abs().max()) == 0.0
    first = obs[0].clone()
    seen = []
    for t in range(3):
        obs, reward, terminated, truncated, info = g.step(0.3 * obs[0])
        seen.append(obs[0].clone())
        assert terminated is False and truncated is (t == 2)
    assert torch.equal(obs[1], seen[1]) and torch.equal(obs[2], seen[0])
    # with a two-frame FIFO the DM stays flat for the first two steps: coefs only receive zeros
    assert float(g.env.dm_prev.abs().max()) > 0.0                       # third step applied the first action
    obs2, _ = g.reset(seed=3)
    assert torch.equal(obs2[0], first)


def test_layers_extruded_in_rounds_use_the_grouped_entry_points(fake, monkeypatch):
    """Philox mode with several layers: one gather_multi / ring_multi per round; injected innovations keep the
    reference's layer-by-layer order through the single-layer entry points."""
    cfg = CONFIGS["tiny"]()
    cfg.windSpeed = [60.0, 45.0]                                        # > 1 px per step: several rounds per frame
    calls = []
    for name in ("aoenv_atm_gather", "aoenv_atm_gather_multi", "aoenv_atm_ring", "aoenv_atm_ring_multi"):
        orig = getattr(fake, name)
        monkeypatch.setattr(fake, name, (lambda o, n: lambda *a: (calls.append(n), o(*a))[1])(orig, name))
    env = build_env(cfg, n_envs=2, rng="philox")
    calls.clear()
    for _ in range(4):
        env.atm.update()
    assert "aoenv_atm_gather_multi" in calls and "aoenv_atm_ring_multi" in calls
    assert calls.count("aoenv_atm_gather_multi") == calls.count("aoenv_atm_ring_multi")
    env_ref = build_env(cfg, n_envs=2, rng="reference")
    before = sum(ly.events for ly in env_ref.atm._layers)
    calls.clear()
    for _ in range(4):
        env_ref.atm.update()
    assert "aoenv_atm_gather_multi" not in calls
    assert calls.count("aoenv_atm_gather") == sum(ly.events for ly in env_ref.atm._layers) - before
    assert torch.isfinite(env.atm.OPD_no_pupil).all()


def test_po4ao_rollout_host_logic_on_cpu(fake):
    """mbrl.run on the CPU stand-in: warm-up episode (integrator + noise) then a policy episode through the ring-buffer
    inference path; replay rows are the consecutive frames of each environment."""
    from rlao_b200.PO4AO import mbrl
    from rlao_b200.PO4AO.conv_models_simple import ConvPolicy, EnsembleDynamics
    from rlao_b200.PO4AO.util_simple import EfficientExperienceReplay, TorchWrapper
    cfg = CONFIGS["tiny"]()
    B, nH, max_ts = 2, 3, 6
    env = TorchWrapper(build_env(cfg, n_envs=B, rng="philox"), host_io=False)
    nA = env.nActuator
    torch.manual_seed(0)
    dynamics = EnsembleDynamics(env.xvalid, env.yvalid, nH, n_models=2)
    policy = ConvPolicy(env.xvalid, env.yvalid, 0.05, env.F.float(), nH)
    rp = EfficientExperienceReplay((nA, nA), (nA, nA), max_size=2 * max_ts * B, n_envs=B)
    past_obs = past_act = obs = None
    for ep in range(2):
        sr, rsum, past_obs, past_act, obs, rewards, _ = mbrl.run(env, past_obs, past_act, obs, rp, policy, dynamics, nH, max_ts,
                                                                 warmup_ts=1, sigma=0.02, episode=ep, iteration=ep)
        assert rsum.shape == (B,) and rewards.shape == (max_ts, B) and np.isfinite(sr)
    st, nx = rp.state().reshape(2, max_ts, B, nA, nA), rp.next_state().reshape(2, max_ts, B, nA, nA)
    assert torch.equal(st[:, 1:], nx[:, :-1])
    assert torch.equal(past_obs[:, -1], st[1, -1]) and torch.equal(past_act[:, -1], rp.action().reshape(2, max_ts, B, nA, nA)[1, -1])
    loss = mbrl.train_dynamics(nH, max_ts, 4, dynamics, torch.optim.Adam(dynamics.parameters()), rp, dyn_iters=1, device="cpu")
    assert np.isfinite(loss)


def test_shift_bookkeeping_matches_oracle_for_random_winds(fake):
    """Integer / fractional wind bookkeeping of updateLayer (OOPAO/Atmosphere.py:350-404): number of add_row events per
    frame and the sub-pixel remainder `buff` for random wind speeds and directions, in lock-step with the oracle."""
    from oracle.ao_oracle import AtmosphereOracle, telescope_pupil
    from rlao_b200.Atmosphere import Atmosphere
    from rlao_b200.Source import Source
    from rlao_b200.Telescope import Telescope
    rs = np.random.RandomState(4)
    for trial in range(4):
        cfg = CONFIGS["tiny"]()
        cfg.windSpeed = list(rs.uniform(3, 70, size=2))
        cfg.windDirection = list(rs.uniform(0, 360, size=2))
        tel = Telescope(cfg.resolution, cfg.diameter, cfg.samplingTime, n_envs=1)
        Source(cfg.opticalBand, cfg.magnitude) * tel
        atm = Atmosphere(tel, cfg.r0, cfg.L0, cfg.windSpeed, cfg.fractionalR0, cfg.windDirection, cfg.altitude,
                         rng="reference", canvas_slack=8)
        atm.initializeAtmosphere(tel)
        orc = AtmosphereOracle(cfg, telescope_pupil(cfg.resolution))
        for k in range(25):
            before = [ly.events for ly in atm._layers]
            n_draws = len(orc.xi_log)
            atm.update()
            orc.update()
            assert sum(ly.events for ly in atm._layers) - sum(before) == len(orc.xi_log) - n_draws, (trial, k)
            for ly, lo in zip(atm._layers, orc.layers):
                assert np.allclose(ly.buff, lo.buff, rtol=0, atol=1e-12), (trial, k)
        assert rel_err(atm.OPD_no_pupil.numpy(), orc.OPD_no_pupil) < 1e-4


def test_pyramid_environment_closes_the_loop_on_cpu(fake):
    """The papyrus-style environment (OOPAOEnv.py: Pyramid WFS, 6-tuple step) on the CPU stand-in: interaction matrix
    through the multi-frame branch, reconstruction, leaky integrator — the residual must drop well below the turbulence."""
    from rlao_b200.OOPAOEnv.OOPAOEnv import OOPAO
    cfg = CONFIGS["tiny"]()
    cfg.nSubap = 12                                        # 48 px pupil / 12 = 4 px per Pyramid subaperture
    p = param_from_config_for_pyramid(cfg)
    env = OOPAO()
    env.set_params_file(p, "")
    env.set_params(types.SimpleNamespace(), gainCL=0.4, n_envs=2, rng="philox", seed=1)
    assert env.wfs.tag == "pyramid" and env.wfs.nSignal == 2 * int(env.wfs.validI4Q.sum())
    assert env.reconstructor.shape == (env.dm.nValidAct, env.wfs.nSignal)
    obs = new_episode(env, 5)
    for i in range(25):
        obs, wfsf, reward, strehl, done, info = env.step(i, env.gainCL * obs)
    assert wfsf.shape == (2, env.wfs.cam.resolution, env.wfs.cam.resolution)
    assert float(env.residual[24].mean()) < 0.5 * float(env.total[24].mean())
    assert float(strehl.min()) > 0 and done is False


def test_torch_wrapper_keeps_the_camera_frame_of_the_papyrus_env(fake, monkeypatch):
    """MAIN_CODE/PO4AO/util_simple.py:209-213 and mbrl.py:76: TorchWrapper.step of the papyrus environment is a 6-tuple
    (obs, wfsf, reward, strehl, done, info), with and without TimeDelayEnv in between; the Razor environment stays a
    5-tuple; mbrl.run accepts both."""
    from rlao_b200.OOPAOEnv.OOPAOEnv import OOPAO
    from rlao_b200.PO4AO import util_simple
    from rlao_b200.PO4AO.util_simple import TimeDelayEnv, TorchWrapper
    monkeypatch.setattr(util_simple.TorchWrapper, "_to_host", lambda self, *t: [x.clone() for x in t])   # no pinned memory on CPU
    cfg = CONFIGS["tiny"]()
    cfg.nSubap = 12
    env = OOPAO()
    env.set_params_file(param_from_config_for_pyramid(cfg), "")
    env.set_params(types.SimpleNamespace(), gainCL=0.4, n_envs=2, rng="philox", seed=1)
    new_episode(env, 5)
    R = env.wfs.cam.resolution
    for wrapped in (TorchWrapper(env, host_io=False), TorchWrapper(env, host_io=True),
                    TorchWrapper(TimeDelayEnv(env, 1), host_io=True), TorchWrapper(TimeDelayEnv(env, 2), host_io=False)):
        obs = wrapped.reset_soft()
        out = wrapped.step(0, env.gainCL * obs)
        assert len(out) == 6
        obs, wfsf, reward, strehl, done, info = out
        assert wfsf.shape == (2, R, R) and obs.shape == (2, env.nActuator, env.nActuator) and info[0][0] == "strehl"
    razor = build_env(CONFIGS["tiny"](), n_envs=2, rng="philox", seed=1)
    new_episode(razor, 5)
    w = TorchWrapper(TimeDelayEnv(razor, 1), host_io=True)
    assert len(w.step(0, razor.gainCL * w.reset_soft())) == 5


def test_checkpoint_resume_continues_bit_for_bit(fake):
    """state_dict / load_state_dict (checkpoint / resume; SURVEY.md section 5 lists it as ours to add): a second environment
    restored from the snapshot of the first continues with identical observations, Strehl ratios and layer maps."""
    cfg = CONFIGS["tiny"]()
    a = build_env(cfg, n_envs=2, rng="philox", seed=3)
    obs = new_episode(a, 7)
    for i in range(6):
        obs, *_ = a.step(i, cfg.gainCL * obs)
    snap = a.state_dict()
    b = build_env(cfg, n_envs=2, rng="philox", seed=3)
    b.load_state_dict(snap)
    obs_b = obs.clone()
    for i in range(6, 12):
        obs, r, s, *_ = a.step(i, cfg.gainCL * obs)
        obs_b, r_b, s_b, *_ = b.step(i, cfg.gainCL * obs_b)
        assert torch.equal(obs, obs_b) and torch.equal(s, s_b), i
    for la, lb in zip(a.atm._layers, b.atm._layers):
        assert torch.equal(la.mapShift, lb.mapShift) and np.array_equal(la.buff, lb.buff) and la.events == lb.events
    with pytest.raises(ValueError):
        build_env(cfg, n_envs=1, rng="philox", seed=3).load_state_dict(snap)


def test_papyrus_science_path_on_cpu(fake):
    """MAIN_CODE/OOPAOEnv/OOPAOEnv.py:118-196,300-339,473-482: guide star + off-axis science target, science cameras on the
    PSF (`atm*src*tel*cam`), long-exposure PSF of render4plot; Detector exposure buffer (OOPAO/Detector.py:232-301)."""
    from rlao_b200.Detector import Detector
    from rlao_b200.OOPAOEnv.OOPAOEnv import OOPAO
    from rlao_b200.Source import Source
    cfg = CONFIGS["tiny"]()
    cfg.nSubap = 12
    env = OOPAO()
    env.set_params_file(param_from_config_for_pyramid(cfg), "")
    env.set_params(types.SimpleNamespace(), gainCL=0.4, n_envs=2, rng="philox", seed=1)
    R = env.tel.resolution
    assert env.tel.fov == 1 and env.ngs.coordinates == [0, 0] and env.src.coordinates == [0.4, 0]
    assert env.tel.src is env.src                                   # the last propagation pointed the telescope at the target
    assert env.src_cam.frame.shape == (2, 4 * R, 4 * R) and env.ngs_cam.frame.shape == (2, R, R)
    assert env.LE_PSF.shape == (2, 4 * R, 4 * R)
    # the WFS-path camera sees the central R x R pixels of the same PSF
    env.atm * env.ngs * env.tel * env.src_cam
    full = env.src_cam.frame
    env.atm * env.ngs * env.tel * env.ngs_cam
    lo = 2 * R - R // 2
    assert torch.allclose(env.ngs_cam.frame, full[:, lo:lo + R, lo:lo + R], rtol=1e-5, atol=0)
    # render4plot: running mean of log10(PSF) after the first 15 frames
    obs = new_episode(env, 5)
    logs = []
    for i in range(14, 19):
        obs, wfsf, reward, strehl, done, info = env.step(i, env.gainCL * obs)
        le, se = env.render4plot(i)
        if i > 15:
            logs.append(se.clone())
    assert torch.allclose(le, torch.stack(logs).mean(dim=0), rtol=1e-5, atol=1e-6)
    # exposure over three AO frames, hardware binning 4: nothing is read out before the integration time is reached
    cam = Detector(integrationTime=3 * env.tel.samplingTime, psf_sampling=2, binning=4)
    env.tel * cam
    env.tel * cam
    assert cam.frame is None and cam.n_buffered == 2
    env.tel.computePSF(2)
    one = env.tel.PSF.clone()
    env.tel * cam
    assert cam.n_buffered == 0 and cam._integrated_time == 0 and cam.n_frames_last_exposure == 3
    want = 3 * one.reshape(2, R // 2, 4, R // 2, 4).sum(dim=(2, 4))
    assert cam.frame.shape == (2, R // 2, R // 2) and torch.allclose(cam.frame, want, rtol=1e-5)
    # sources outside the field of view / layers in altitude are refused
    with pytest.raises(ValueError):
        env.atm * Source("I", 8, coordinates=[2.0, 0])
    env.atm.altitude[0] = 5000.0
    with pytest.raises(NotImplementedError):
        env.atm * env.src
    env.atm.altitude[0] = 0.0


def param_from_config_for_pyramid(cfg):
    from parity_util import param_from_config
    p = param_from_config(cfg)
    p.update(modulation=3, n_pix_separation=4, lightThreshold=0.1, postProcessing="slopesMaps_incidence_flux", nLoop=64,
             cam_photonNoise=False, cam_readoutNoise=0, nZernike=20)
    return p


def test_product_never_touches_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under rlao_b200/ may import it or read /root/reference, and bench.py may
    reach it only in its CPU arms."""
    pkg = os.path.join(ROOT, "rlao_b200")
    offenders = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "/root/reference" in src:
                    offenders.append(os.path.join(dp, f))
    assert offenders == []
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle", bench)]
    assert uses and all(bench.rfind("\ndef ", 0, u) >= 0 for u in uses)
    for u in uses:                                       # every use sits inside a CPU-arm helper
        fn = bench[bench.rfind("\ndef ", 0, u):u].split("(")[0]
        assert fn.strip().split()[-1] in ("oracle_config", "cpu_env", "reference_env"), fn
    assert "/root/reference" not in bench


def test_remaining_environment_methods(fake):
    """step_wfs (frame observation, no leak, action in metres), OPD_on_dm / compute_dm_proj, _get_reward, set_modalBasis,
    set_wfs (OOPAOEnvRazor.py:342-425, 553-586, 607-614, 676-689)."""
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=2, rng="philox")
    new_episode(env, 3)
    coefs0 = env.dm_prev.clone()
    act = torch.zeros((2, cfg.nSubap + 1, cfg.nSubap + 1))
    act[:, env.xvalid[5], env.yvalid[5]] = 2e-8
    frame, reward, strehl, done, info = env.step_wfs(0, act)
    assert frame.shape == (2, cfg.resolution, cfg.resolution) and reward.shape == (2,) and done is False
    assert torch.allclose(env.dm_prev - coefs0, env.img_to_vec(act).float(), rtol=1e-5, atol=1e-14)      # += action, no leak
    assert env.leak == cfg.leak
    assert torch.equal(env._get_reward(), env.get_strehl())
    # projection of an OPD that lies in the DM space gives it back
    c = torch.zeros((2, env.dm.nValidAct), dtype=torch.float64)
    c[:, 7], c[:, 20] = 3e-8, -1e-8
    opd = (c @ env.dm.modes.double().T).reshape(2, cfg.resolution, cfg.resolution)
    env.tel.OPD_no_pupil = opd.float()
    pupil = torch.as_tensor(env.tel.pupil, dtype=torch.float64)
    back = env.OPD_on_dm()
    want = ((opd * pupil).reshape(2, -1) @ env.dm_proj.T) @ env.dm.modes.double().T
    assert rel_err(back.reshape(2, -1).numpy(), want.numpy()) < 2e-6          # tel.OPD is float32
    R0 = env.reconstructor.clone()
    env.set_modalBasis("zernike")
    assert env.calib_CL.M.shape[1] == env.wfs.nSignal and env.M2C_CL.shape == (env.dm.nValidAct, 50)
    assert torch.equal(env.reconstructor, R0)                       # the reference does not touch the reconstructor here
    with pytest.raises(NotImplementedError):
        env.set_modalBasis("KL")
    p = dict(env.param)
    env.set_wfs(p, "shackhartmann")
    assert env.wfs.tag == "shackHartmann" and env.wfs.telescope is env.tel

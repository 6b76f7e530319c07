"""Host-layer tests that run without a GPU: the Python object protocol and env.step sequencing of rlao_b200
are driven on the CPU through tests/fake_backend.py (a numpy stand-in for the C ABI) and compared with the
oracle; plus: the shared library loads and exports every symbol include/aoenv.h declares."""
import ctypes
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


def test_env_step_sequence_matches_oracle_on_cpu(fake):
    cfg = CONFIGS["tiny"]()
    env = build_env(cfg, n_envs=1, rng="reference")
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

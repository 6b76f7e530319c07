"""Gymnasium-style face of the Shack-Hartmann environment.

The reference documents this signature on its Pyramid environments (MAIN_CODE/OOPAOEnv/OOPAOEnv_VPG.py:117-137
`reset(seed, options) -> (obs, info)`, :553-611 `step(action) -> (obs, reward, terminated, truncated, info)`):
the observation is a stack of the last `n_history` reconstructed-command images (newest first, `roll_buffer`
:660-682), actions pass through a FIFO of `delay` frames (:560-566), the reward is the Strehl ratio (:597-601) and
the episode never terminates by itself.  This adapter gives the same face to `OOPAOEnvRazor.OOPAO` (SURVEY.md
section 8 b) without adding work to the step: history and delay are ring buffers on the device.

`gymnasium` itself is not a dependency; `observation_space` / `action_space` are plain `Box` records with the
fields RL libraries read (`low`, `high`, `shape`, `dtype`).
"""
import collections

import numpy as np
import torch

Box = collections.namedtuple("Box", ["low", "high", "shape", "dtype"])


class GymnasiumSH:
    """env = GymnasiumSH(oopao_env, n_history=20, delay=1, episode_length=None)

    obs: float32 [n_envs, n_history, nAct, nAct] on the device ([n_history, nAct, nAct] for one environment);
    action: [n_envs, nAct, nAct] image or [n_envs, nValidAct] vector, micro-metres (scaled by 1e-6 in the step
    kernel like OOPAOEnv_VPG.py:566).  `truncated` turns True after `episode_length` steps when one is given
    (the reference leaves truncation to a TimeLimit wrapper).
    """

    metadata = {"render_modes": []}

    def __init__(self, env, n_history=20, delay=1, episode_length=None):
        self.env = env
        self.n_history = int(n_history)
        self.d = int(delay)
        self.episode_length = episode_length
        nA, B = env.nActuator, env.n_envs
        lead = () if B == 1 else (B,)
        self.observation_space = Box(-np.inf, np.inf, lead + (self.n_history, nA, nA), np.float32)
        self.action_space = Box(-1.0, 1.0, lead + (nA, nA), np.float32)
        self._hist = torch.zeros((B, self.n_history, nA, nA), dtype=torch.float32, device=env.device)
        self._fifo = torch.zeros((max(self.d, 1), B, nA, nA), dtype=torch.float32, device=env.device)
        self._t = 0
        self._rng = np.random.RandomState()

    def __getattr__(self, name):
        return getattr(self.__dict__["env"], name)

    # ---- helpers ------------------------------------------------------------------------------------------
    def _push(self, obs):
        """roll_buffer (OOPAOEnv_VPG.py:660-682): newest image at index 0."""
        self._hist = torch.roll(self._hist, shifts=1, dims=1)
        self._hist[:, 0] = obs.reshape(self.env.n_envs, *obs.shape[-2:])

    def _out(self):
        h = self._hist[0] if self.env.n_envs == 1 else self._hist
        return h.clone()

    # ---- gymnasium API --------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        """OOPAOEnv_VPG.py:117-137: flat DM, new phase screens, WFS measurement of the bare atmosphere."""
        if seed is not None:
            self._rng = np.random.RandomState(seed)
        e = self.env
        e.dm.coefs = 0
        e.dm_prev = 0
        e.atm.generateNewPhaseScreen(seed=int(self._rng.randint(0, 100000)))
        e.tel * e.wfs
        e.SR = []
        self._fifo.zero_()
        self._hist.zero_()
        self._t = 0
        self._push(e.reset_soft())
        return self._out(), {}

    def step(self, action):
        """OOPAOEnv_VPG.py:553-611."""
        e = self.env
        a = torch.as_tensor(action, dtype=torch.float32, device=e.device)
        if a.shape[-1] == e.dm.nValidAct and a.shape[-2:] != (e.nActuator, e.nActuator):
            a = e.vec_to_img(a)                                          # :558-559
        a = a.reshape(e.n_envs, e.nActuator, e.nActuator)
        if self.d > 0:                                                   # :561-566 FIFO of `delay` frames
            slot = self._t % self.d
            delayed = self._fifo[slot].clone()
            self._fifo[slot] = a
        else:
            delayed = a
        obs, _, strehl, _, info = e._step_views(self._t, delayed)
        self._t += 1
        self._push(obs)
        reward = strehl.clone() if torch.is_tensor(strehl) else strehl   # :597-601 reward = Strehl
        truncated = self.episode_length is not None and self._t >= self.episode_length
        return self._out(), reward, False, bool(truncated), {"strehl": reward}

    def close(self):
        pass

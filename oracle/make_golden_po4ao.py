"""Generates tests/golden/po4ao.npz by running the UNMODIFIED reference PO4AO modules
(/root/reference/drl4ao/MAIN_CODE/PO4AO/{conv_models_simple,mbrl,util_simple}.py) on the CPU of the build
container — TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden_po4ao

The fixture holds, for a 9x9 actuator grid with n_history = 3:
  * the freshly initialised weights of a ConvPolicy and a 2-member EnsembleDynamics;
  * their forward outputs on seeded inputs;
  * a seeded replay of 2 x 12 transitions, the windows `sample_contiguous` draws from it after torch.manual_seed,
    and the losses + weight digests after `train_dynamics(dyn_iters=2)` and `train_policy(pol_iters=2, T=3)`.
Modules the reference imports at file scope but does not use on this path (tensorboard, matplotlib, the Pyramid
environment, PO4AO.util.Dynamics) are replaced by empty stand-ins in sys.modules.
"""
import os
import sys
from unittest import mock

import numpy as np
import torch

REF_MAIN = "/root/reference/drl4ao/MAIN_CODE"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "po4ao.npz")

N_ACT, N_HISTORY, MAX_TS, BATCH, T_HORIZON = 9, 3, 12, 4, 3


def problem():
    """Actuator mask, projector F and a synthetic replay — shared with tests/test_po4ao.py through the fixture."""
    rs = np.random.RandomState(7)
    yy, xx = np.mgrid[:N_ACT, :N_ACT] - (N_ACT - 1) / 2
    xvalid, yvalid = np.nonzero(xx ** 2 + yy ** 2 <= (N_ACT / 2 + 0.3) ** 2)
    q, _ = np.linalg.qr(rs.normal(size=(len(xvalid), 14)))
    F = (q @ q.T).astype(np.float32)
    n = 2 * MAX_TS
    states = rs.normal(size=(n + 1, N_ACT, N_ACT)).astype(np.float32)
    actions = (0.3 * rs.normal(size=(n, N_ACT, N_ACT))).astype(np.float32)
    rewards = rs.normal(size=(n,)).astype(np.float32)
    return xvalid, yvalid, F, states, actions, rewards


def import_reference():
    for name in ("torch.utils.tensorboard", "torch.utils.tensorboard.writer", "matplotlib", "matplotlib.pyplot", "gym",
                 "OOPAOEnv", "OOPAOEnv.OOPAOEnv", "PO4AO.util"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock()
    sys.modules["gym"].Wrapper = object
    sys.path.insert(0, REF_MAIN)
    import PO4AO.conv_models_simple as models
    import PO4AO.mbrl as mbrl
    import PO4AO.util_simple as util
    return models, mbrl, util


def digest(sd):
    out = []
    for k in sorted(sd):
        a = sd[k].detach().double().numpy().reshape(-1)
        out.append([a.sum(), np.abs(a).sum(), a[0], a[a.size // 2], a[-1]])
    return np.asarray(out)


def main():
    models, mbrl, util = import_reference()
    xvalid, yvalid, F, states, actions, rewards = problem()
    g = dict(xvalid=xvalid, yvalid=yvalid, F=F, states=states, actions=actions, rewards=rewards,
             sizes=np.array([N_ACT, N_HISTORY, MAX_TS, BATCH, T_HORIZON]))
    torch.manual_seed(0)
    dynamics = models.EnsembleDynamics(xvalid, yvalid, N_HISTORY, n_models=2)
    policy = models.ConvPolicy(xvalid, yvalid, 0.1, torch.from_numpy(F), N_HISTORY)
    for k, v in policy.state_dict().items():
        g["policy/" + k] = v.numpy().copy()
    for k, v in dynamics.state_dict().items():
        g["dynamics/" + k] = v.numpy().copy()

    rs = np.random.RandomState(3)
    s = torch.from_numpy(rs.normal(size=(3, 1, N_ACT, N_ACT)).astype(np.float32))
    a = torch.from_numpy(rs.normal(size=(3, 1, N_ACT, N_ACT)).astype(np.float32))
    h = torch.from_numpy(rs.normal(size=(3, 2 * (N_HISTORY - 1), N_ACT, N_ACT)).astype(np.float32))
    g.update(fwd_state=s.numpy(), fwd_action=a.numpy(), fwd_history=h.numpy())
    with torch.no_grad():
        g["fwd_policy"] = policy(s, h).numpy()
        g["fwd_policy_single"] = policy(s[0, 0], h[:1]).numpy()            # the [nAct, nAct] call of mbrl.py:72
        g["fwd_dynamics"] = dynamics(s, a, h).numpy()

    replay = util.EfficientExperienceReplay((N_ACT, N_ACT), (N_ACT, N_ACT))
    for i in range(len(actions)):
        replay.append(torch.from_numpy(states[i]), torch.from_numpy(actions[i]), float(rewards[i]),
                      torch.from_numpy(states[i + 1]), False)
    torch.manual_seed(11)
    smp = replay.sample_contiguous(N_HISTORY, MAX_TS, BATCH)
    g["sample_states"], g["sample_actions"] = smp.state().numpy().copy(), smp.action().numpy().copy()

    dyn_opt = torch.optim.Adam(dynamics.parameters())
    pol_opt = torch.optim.Adam(policy.parameters())
    torch.manual_seed(1)
    g["dyn_loss"] = np.float64(mbrl.train_dynamics(N_HISTORY, MAX_TS, BATCH, dynamics, dyn_opt, replay, dyn_iters=2, device="cpu"))
    g["dyn_digest"] = digest(dynamics.state_dict())
    torch.manual_seed(2)
    g["pol_loss"] = np.float64(mbrl.train_policy(pol_opt, policy, dynamics, replay, "cpu", N_HISTORY, MAX_TS, BATCH, T_HORIZON,
                                                 pol_iters=2))
    g["pol_digest"] = digest(policy.state_dict())
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB; dyn_loss", g["dyn_loss"], "pol_loss", g["pol_loss"])


if __name__ == "__main__":
    main()

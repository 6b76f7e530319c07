"""PO4AO networks — mirror of MAIN_CODE/PO4AO/conv_models_simple.py (ConvDynamics :11-52, ConvPolicy :57-113,
EnsembleDynamics :118-135) for batches of environments.

Same class names, constructor arguments, parameter names (`net.0.weight` ... so reference checkpoints load with
`load_state_dict`) and forward semantics; plain PyTorch (cuDNN) as SURVEY.md section 8 f-1 prescribes.  What
changes is the execution: every forward takes [B, C, nAct, nAct] on the device, the valid-actuator mask is a
registered buffer applied with one multiply (the reference scatters through fancy indices into a fresh zero
tensor), and the policy's projection on the controlled subspace is one [B, nA] x [nA, nA] product.
"""
import torch
import torch.nn as nn

n_filt = 64


def _conv_stack(c_in):
    """Three 3x3 convolutions, 64 filters, LeakyReLU between (conv_models_simple.py:20-28, 67-74)."""
    return nn.Sequential(
        nn.Conv2d(c_in, n_filt, 3, padding=1), nn.LeakyReLU(),
        nn.Conv2d(n_filt, n_filt, 3, padding=1), nn.LeakyReLU(),
        nn.Conv2d(n_filt, 1, 3, padding=1))


def _as_index(v):
    return torch.as_tensor(v, dtype=torch.long).reshape(-1).cpu()


def _as_batch(x):
    """[H, W] -> [1, 1, H, W]; [C, H, W] -> [1, C, H, W] (the reference's `view(1, *shape)` calls)."""
    while x.ndim < 4:
        x = x.unsqueeze(0)
    return x


class _ActuatorGrid(nn.Module):
    """Holds the valid-actuator bookkeeping shared by the three networks."""

    def __init__(self, xvalid, yvalid):
        super().__init__()
        self.xvalid, self.yvalid = _as_index(xvalid), _as_index(yvalid)
        self._side = None
        self.register_buffer("_mask", torch.zeros(0), persistent=False)
        self.register_buffer("_flat", torch.zeros(0, dtype=torch.long), persistent=False)

    def _grid(self, like):
        rows, cols = like.shape[-2], like.shape[-1]
        if self._side != (rows, cols) or self._mask.device != like.device or self._mask.dtype != like.dtype:
            m = torch.zeros((rows, cols), dtype=like.dtype, device=like.device)
            m[self.xvalid.to(like.device), self.yvalid.to(like.device)] = 1
            self._mask = m
            self._flat = (self.xvalid * cols + self.yvalid).to(like.device)
            self._side = (rows, cols)
        return self._mask, self._flat


class ConvDynamics(_ActuatorGrid):
    """next_state = net([history, state, action]) on the valid actuators, zero elsewhere (:30-52)."""

    def __init__(self, xvalid, yvalid, n_history):
        super().__init__(xvalid, yvalid)
        self.n_history = n_history
        self.net = _conv_stack(n_history * 2)

    def forward(self, states, actions, history=None):
        states, actions = _as_batch(states), _as_batch(actions)
        parts = [states, actions] if history is None else [_as_batch(history), states, actions]
        out = self.net(torch.cat(parts, dim=1))
        mask, _ = self._grid(out)
        return out * mask


class ConvPolicy(_ActuatorGrid):
    """action = F @ clamp(net([state, history]), -1, 1) on the valid actuators (:82-113)."""

    def __init__(self, xvalid, yvalid, sigma, F, n_history):
        super().__init__(xvalid, yvalid)
        self.n_history = n_history
        self.register_buffer("F", torch.as_tensor(F, dtype=torch.float32).unsqueeze(0))
        self.net = _conv_stack(n_history * 2 - 1)
        self.sigma = sigma

    def weights_init(self, standart_dev=0.1, mean_bias=0):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, mean=0, std=standart_dev)
                nn.init.constant_(m.bias, mean_bias)

    def forward(self, state, history=None):
        state = _as_batch(state)
        feats = state if history is None else torch.cat([state, _as_batch(history)], dim=1)
        return self.project(self.net(feats))

    def project(self, out):
        """clamp to [-1, 1], keep the valid actuators, apply F (:94-108); out [B, 1, nAct, nAct]."""
        out = out.clamp(-1, 1)
        _, flat = self._grid(out)
        B, rows, cols = out.shape[0], out.shape[-2], out.shape[-1]
        vec = out.reshape(B, -1)[:, flat] @ self.F[0].t()            # F @ v for every environment
        ret = torch.zeros((B, rows * cols), dtype=out.dtype, device=out.device)
        ret[:, flat] = vec
        return ret.reshape(B, 1, rows, cols)


class EnsembleDynamics(nn.Module):
    """n_models independent ConvDynamics; forward concatenates their predictions on the channel axis (:118-135)."""

    def __init__(self, xvalid, yvalid, n_history, n_models=5):
        super().__init__()
        self.n_models = n_models
        self.models = nn.ModuleList([ConvDynamics(xvalid, yvalid, n_history) for _ in range(n_models)])

    def forward(self, states, actions, history=None):
        return torch.cat([m(states, actions, history) for m in self.models], dim=1)

"""PO4AO rollout and training loops — mirror of MAIN_CODE/PO4AO/mbrl.py (`get_env` :22-31, `run` :33-91,
`train_dynamics` :94-142, `train_policy` :148-221) and of its Shack-Hartmann twin mbrl_funcsRAZOR.py, for a batch of
environments that lives on the GPU.

Differences from the reference, none of them in the arithmetic:
  * `run` steps `env.n_envs` environments in lock-step; observations, telemetry histories, actions and the replay
    stay on the device and nothing in the loop synchronises with the host (the reference converts
    torch -> numpy -> torch around every step and rolls its histories with `torch.cat`);
  * the histories are sliding windows over a double-length buffer: one image is written per step instead of the
    whole (n_history - 1)-deep stack being re-concatenated;
  * `train_dynamics` / `train_policy` leave the models on `device` (the reference moves them to the CPU and back
    every episode) and, when `torch.distributed` is initialised, average the gradients over the ranks so that each
    GPU trains on the windows of its own shard of environments.
With one environment and the same torch seed the sampled windows, losses and updated weights are the
reference's (tests/test_po4ao.py checks them against fixtures generated from the reference's own modules).
"""
import torch
import torch.distributed as dist

from .util_simple import TimeDelayEnv, TorchWrapper


def get_env(args, gainCL=0.2, wfs_type="shackhartmann", n_envs=1, device=None, host_io=False, **kw):
    """mbrl_funcsRAZOR.py:25-34 (`get_env(args)` of mbrl.py:22-31 with the SH defaults)."""
    from ..OOPAOEnv.OOPAOEnvRazor import OOPAO
    env = OOPAO()
    env.set_params_file(getattr(args, "param_file", None), getattr(args, "oopao_path", ""))
    env.set_params(args, wfs_type, gainCL=gainCL, n_envs=n_envs, device=device, **kw)
    if getattr(args, "delay", 0) > 0:
        env = TimeDelayEnv(env, args.delay)
    return TorchWrapper(env, host_io=host_io)


class _History:
    """Last `depth` images per environment, oldest first, as a view: a buffer of 2*depth slots is written at
    `p` and `p + depth`, so the window [p+1, p+1+depth) is always contiguous."""

    def __init__(self, init):
        self.depth = init.shape[1]
        self.buf = torch.cat([init, init], dim=1).contiguous() if self.depth > 0 else init
        self.p = self.depth - 1 if self.depth > 0 else 0

    def push(self, img):
        if self.depth == 0:
            return
        self.p = (self.p + 1) % self.depth
        self.buf[:, self.p] = img
        self.buf[:, self.p + self.depth] = img

    def window(self):
        if self.depth == 0:
            return self.buf
        return self.buf[:, self.p + 1:self.p + 1 + self.depth]


class _RolloutPolicy:
    """ConvPolicy inference inside `run` without moving the telemetry: instead of rolling the (n_history - 1)-deep
    stacks and concatenating [state, past_obs, past_act] every step (mbrl.py:72, 79-80), the images stay where they
    were written — two rings inside ONE channels-last input tensor — and the input channels of the first
    convolution's weight are permuted to follow the rings (a 64 x (2 n_history - 1) x 3 x 3 gather per step).  A
    convolution sums over its input channels, so permuting channels and weights together leaves the result unchanged
    up to the order of that sum.

    channels [0, n_h)         observation ring: slot p_o holds the current state, the others the past observations
    channels [n_h, 2 n_h - 1)  action ring:      slot p_a holds the newest action
    """

    @classmethod
    def attach(cls, policy, past_obs, past_act, n_history, graph=True):
        """One instance (and one pair of captured graphs) per policy and batch shape, reused from episode to episode:
        the graphs read the weights through their storage, which the optimiser updates in place."""
        key = (tuple(past_obs.shape), str(past_obs.device), n_history, bool(graph))
        cache = policy.__dict__.setdefault("_rollout_cache", {})
        inst = cache.get(key)
        if inst is None or inst.w0 is not policy.net[0].weight:
            inst = cache[key] = cls(policy, past_obs, past_act, n_history, graph)
        else:
            inst.load(past_obs, past_act)
        return inst

    def load(self, past_obs, past_act):
        """Start of an episode: histories into the rings (past_obs[j], 0 = oldest, in slot j; the next state goes to slot d)."""
        self.inp[:, :self.d] = past_obs
        self.inp[:, self.d] = 0
        self.inp[:, self.nH:] = past_act
        self.p_o, self.p_a = self.d - 1, self.d - 1
        if self._graphs is not None:
            self._po.fill_(self.p_o)
            self._pa.fill_(self.p_a)

    def __init__(self, policy, past_obs, past_act, n_history, graph=True):
        B, d, nA = past_obs.shape[0], n_history - 1, past_obs.shape[-1]
        dev = past_obs.device
        self.policy, self.nH, self.d = policy, n_history, d
        self.inp = torch.zeros((B, 2 * n_history - 1, nA, nA), dtype=torch.float32, device=dev).contiguous(
            memory_format=torch.channels_last)
        # age order -> slots: past_obs[j] (0 = oldest) lives in slot j, the next state goes to slot d  (p_o = d)
        self.inp[:, :d] = past_obs
        self.inp[:, n_history:] = past_act
        self.p_o, self.p_a = d - 1, d - 1            # advanced before each write
        self.net = policy.net
        self.w0, self.b0 = self.net[0].weight, self.net[0].bias
        self.rest = self.net[1:]
        # perm[p_o, p_a][slot] = reference channel whose weights that slot needs
        po = torch.arange(n_history).view(-1, 1, 1)
        pa = torch.arange(d).view(1, -1, 1)
        slot_o = torch.arange(n_history).view(1, 1, -1)
        slot_a = torch.arange(d).view(1, 1, -1)
        age_o = (slot_o - po - 1) % n_history                          # 0 .. n_h-2 = past (oldest first), n_h-1 = state
        ref_o = torch.where(age_o == n_history - 1, torch.zeros_like(age_o), age_o + 1).expand(n_history, d, n_history)
        ref_a = (n_history + (slot_a - pa - 1) % d).expand(n_history, d, d)
        self.perm = torch.cat([ref_o, ref_a], dim=2).to(dev)             # [n_h, d, 2 n_h - 1]
        self._graphs = None
        if graph and dev.type == "cuda":
            self._capture(B, nA, dev)

    # ---- CUDA-graph path: ring positions live on the device, so one captured graph serves every step ------------
    def _act_body(self):
        self._po.add_(1).remainder_(self.nH)
        self.inp.index_copy_(1, self._po, self._obs_in.unsqueeze(1))
        row = self.perm.view(self.nH * self.d, -1).index_select(0, self._po * self.d + self._pa)
        w = self.w0.index_select(1, row[0])
        out = torch.nn.functional.conv2d(self.inp, w, self.b0, padding=1)
        self._act_out.copy_(self.policy.project(self.rest(out))[:, 0])

    def _record_body(self):
        self._pa.add_(1).remainder_(self.d)
        self.inp.index_copy_(1, self._pa + self.nH, self._act_in.unsqueeze(1))

    def _capture(self, B, nA, dev):
        """Warm up on a side stream, restore the rings, capture the two step bodies.  Any failure leaves the eager path."""
        try:
            self._po = torch.tensor([self.p_o], dtype=torch.long, device=dev)
            self._pa = torch.tensor([self.p_a], dtype=torch.long, device=dev)
            self._obs_in = torch.zeros((B, nA, nA), dtype=torch.float32, device=dev)
            self._act_in = torch.zeros((B, nA, nA), dtype=torch.float32, device=dev)
            self._act_out = torch.zeros((B, nA, nA), dtype=torch.float32, device=dev)
            saved = (self.inp.clone(), self._po.clone(), self._pa.clone())
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._act_body()
                    self._record_body()
            torch.cuda.current_stream(dev).wait_stream(side)
            g_act, g_rec = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_act):
                self._act_body()
            with torch.cuda.graph(g_rec):
                self._record_body()
            self.inp.copy_(saved[0])
            self._po.copy_(saved[1])
            self._pa.copy_(saved[2])
            self._graphs = (g_act, g_rec)
        except Exception:
            self._graphs = None

    def act(self, obs):
        """obs [B, nAct, nAct] (the current state) -> action [B, nAct, nAct]."""
        self.p_o = (self.p_o + 1) % self.nH
        if self._graphs is not None:
            self._obs_in.copy_(obs)
            self._graphs[0].replay()
            return self._act_out.clone()
        self.inp[:, self.p_o] = obs
        w = self.w0.index_select(1, self.perm[self.p_o, self.p_a])
        out = torch.nn.functional.conv2d(self.inp, w, self.b0, padding=1)
        return self.policy.project(self.rest(out))[:, 0]

    def record(self, action):
        self.p_a = (self.p_a + 1) % self.d
        if self._graphs is not None:
            self._act_in.copy_(action)
            self._graphs[1].replay()
            return
        self.inp[:, self.nH + self.p_a] = action

    def histories(self):
        """(past_obs, past_act) in the reference's order (oldest first), past_obs including the last state."""
        o = [(self.p_o + 1 + j + 1) % self.nH for j in range(self.d)]
        a = [(self.p_a + 1 + j) % self.d + self.nH for j in range(self.d)]
        return self.inp[:, o].contiguous(), self.inp[:, a].contiguous()


@torch.no_grad()
def run(env, past_obs, past_act, obs, replay, policy, dynamics, n_history, max_ts, warmup_ts, sigma, writer=None,
        episode=0, iteration=0, use_recon=False, reconstructor=None, new_screen=True):
    """One episode of `max_ts` frames (mbrl.py:33-91).

    Returns (env.calculate_strehl_AVG(), reward_sum, past_obs, past_act, obs, rewards, iteration) like the
    reference; with n_envs > 1 `reward_sum` is a [n_envs] tensor, `rewards` a [max_ts, n_envs] tensor, and the
    histories are [n_envs, n_history-1, nAct, nAct].
    """
    dynamics.eval()
    policy.eval()
    B = env.n_envs
    if new_screen:                                            # mbrl.py:49-52 (the SH twin leaves this to the caller)
        env.atm.generateNewPhaseScreen(93234 * iteration)
        env.dm.coefs = 0
        env.tel * env.dm * env.wfs
    obs = env.reset_soft()
    dev = obs.device
    obs = obs.reshape(B, *obs.shape[-2:])
    nA = obs.shape[-1]
    if past_obs is None:
        past_obs = torch.zeros((B, n_history - 1, nA, nA), dtype=torch.float32, device=dev)
        past_act = torch.zeros((B, n_history - 1, nA, nA), dtype=torch.float32, device=dev)
    past_obs = past_obs.reshape(B, n_history - 1, nA, nA).to(dev)
    past_act = past_act.reshape(B, n_history - 1, nA, nA).to(dev)
    use_policy = episode >= warmup_ts
    fast = use_policy and n_history > 1 and hasattr(policy, "project") and hasattr(policy, "net")
    if fast:
        roll = _RolloutPolicy.attach(policy, past_obs, past_act, n_history)
    else:
        h_obs, h_act = _History(past_obs), _History(past_act)
    rewards = torch.empty((max_ts, B), dtype=torch.float32, device=dev)
    squeeze = (lambda t: t[0]) if B == 1 else (lambda t: t)

    for t in range(max_ts):
        if not use_policy:                                     # integrator + exploration noise (:67-70)
            action = env.gainCL * obs + env.sample_noise(sigma).reshape(B, nA, nA).to(dev)
        elif fast:                                             # :72-73 without moving the telemetry
            action = roll.act(obs)
        else:
            history = torch.cat([h_obs.window(), h_act.window()], dim=1) if n_history > 1 else None
            action = policy(obs.unsqueeze(1), history)[:, 0]
        action = action.to(torch.float32)
        out = env.step(t, squeeze(action))          # 5-tuple, or 6 with the camera frame second (mbrl.py:76)
        next_obs, (reward, strehl, done) = out[0], out[-4:-1]
        next_obs = torch.as_tensor(next_obs, device=dev).reshape(B, nA, nA)
        if fast:                                               # :79-80 (the state is already in its ring slot)
            roll.record(action)
        else:
            h_obs.push(obs)
            h_act.push(action)
        rewards[t] = torch.as_tensor(reward, dtype=torch.float32, device=dev).reshape(B)
        replay.append(squeeze(obs), squeeze(action), squeeze(rewards[t]), squeeze(next_obs), done)   # :86
        obs = next_obs

    reward_sum = rewards.sum(dim=0)
    past_obs, past_act = roll.histories() if fast else (h_obs.window().clone(), h_act.window().clone())
    if B == 1:
        return (env.calculate_strehl_AVG(), float(reward_sum[0]), past_obs, past_act, obs[0],
                [float(r) for r in rewards[:, 0].cpu()], iteration)
    return env.calculate_strehl_AVG(), reward_sum, past_obs, past_act, obs, rewards, iteration


def _windows(sample, batch_size, n_history, device):
    """The unfolding both trainers apply to a `sample_contiguous` draw (mbrl.py:113-126, 172-186):
    (state, action) = frame n_history-1 of each window, history = the n_history-1 frames before it, target = the
    state of the last frame."""
    states = sample.state().to(device)
    actions = sample.action().to(device)
    states = states.view(batch_size, n_history + 1, *states.shape[1:])
    actions = actions.view(batch_size, n_history + 1, *actions.shape[1:])
    past_obs, past_act = states[:, :n_history - 1], actions[:, :n_history - 1]
    state, action = states[:, n_history - 1:n_history], actions[:, n_history - 1:n_history]
    return state, action, past_obs, past_act, states[:, -1:]


def _average_gradients(params):
    """Data-parallel step between the ranks that each hold a shard of the environments (SURVEY.md section 8 e)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= dist.get_world_size()
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def train_dynamics(n_history, max_ts, batch_size, dynamics, optimizer, replay, dyn_iters=5, device="cuda:0",
                   keep_on_device=True):
    """mbrl.py:94-142: every member of the ensemble regresses the next state on its own bootstrap of windows."""
    dynamics.train()
    dynamics.to(device)
    loss = torch.zeros((), device=device)
    for _ in range(dyn_iters):
        optimizer.zero_grad()
        loss = 0
        for member in dynamics.models:
            sample = replay.sample_contiguous(n_history, max_ts, batch_size)
            state, action, past_obs, past_act, target = _windows(sample, batch_size, n_history, device)
            pred = member(state, action, torch.cat([past_obs, past_act], dim=1))
            assert pred.shape == target.shape
            loss = loss + (target - pred).pow(2).mean()
        loss.backward()
        _average_gradients(list(dynamics.parameters()))
        torch.nn.utils.clip_grad_norm_(dynamics.parameters(), 0.5)
        optimizer.step()
    if not keep_on_device:
        dynamics.to("cpu")
    return loss.item()


def loss_fn(state, action):
    """mbrl.py:145-146."""
    return state.pow(2).mean() + 0.001 * action.pow(2).mean()


def train_policy(opt, policy, dynamics, replay, device, n_history, max_ts, batch_size, T, pol_iters=5,
                 keep_on_device=True):
    """mbrl.py:148-221: back-propagate the T-step model rollout cost through the frozen ensemble."""
    dynamics.train()
    policy.train()
    for p in dynamics.parameters():
        p.requires_grad_(False)
    policy.to(device)
    dynamics.to(device)
    loss = torch.zeros((), device=device)
    for _ in range(pol_iters):
        opt.zero_grad()
        sample = replay.sample_contiguous(n_history, max_ts, batch_size)
        state, action, past_obs, past_act, _ = _windows(sample, batch_size, n_history, device)
        # NB (:188, 204): `losses` is a [batch] vector to which a scalar (batch-mean) stage cost is broadcast
        cost = torch.zeros((), device=device)
        for _t in range(T):
            if n_history > 1:
                history = torch.cat([past_obs, past_act], dim=1)
                action = policy(state, history)
                next_state = dynamics(state, action, history)
            else:
                action = policy(state)
                next_state = dynamics(state, action)
            cost = cost + loss_fn(next_state[:, 0], action)
            past_act = torch.cat([past_act[:, 1:], action], dim=1)       # :207-208
            past_obs = torch.cat([past_obs[:, 1:], state], dim=1)
            state = next_state.mean(dim=1, keepdim=True)                 # :211-212 ensemble mean
        loss = cost
        loss.backward()
        _average_gradients(list(policy.parameters()))
        opt.step()
    for p in dynamics.parameters():
        p.requires_grad_(True)
    if not keep_on_device:
        policy.to("cpu")
        dynamics.to("cpu")
    return loss.item()

"""PO4AO experiment loop — the episode/training schedule of MAIN_CODE/mbrl_main.py:73-160 on top of the batched GPU
environment (no tensorboard / plotting; statistics are returned and printed).

Run on one GPU:           python -m rlao_b200.PO4AO.main --n-envs 1024
Run on N GPUs (one node): python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 -m rlao_b200.PO4AO.main
Each rank owns a shard of the environments (its own atmosphere seeds and replay); gradients are averaged over ranks
inside train_dynamics / train_policy and the episode Strehl is averaged over ranks by env.calculate_strehl_AVG().
"""
import argparse
import os
import time
import types

import torch
import torch.distributed as dist
from torch import optim

from . import mbrl
from .conv_models_simple import ConvPolicy, EnsembleDynamics
from .util_simple import EfficientExperienceReplay


def default_args(**over):
    """The hyper-parameters the reference keeps in Conf/razor_config_po4ao.yaml / papyrus_config.yaml."""
    a = dict(param_file="rlao_b200.Conf.parameter_file_synthetic_SHWFS", oopao_path="", delay=1, n_history=20, max_ts=500,
             warmup_ts=5, iters=20, batch_size=32, T=4, initial_sigma=0.3, nSubaperture=20, r0=0.13, L0=25,
             fractionalR0=[1.0], windSpeed=[10], windDirection=[0], altitude=[0], nLoop=None, gainCL=0.5, n_envs=1024,
             replay_episodes=4, seed=5)
    a.update(over)
    return types.SimpleNamespace(**a)


def main(args=None, verbose=True):
    args = args or default_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)                       # same initial weights on every rank
    B = args.n_envs
    env = mbrl.get_env(args, gainCL=args.gainCL, n_envs=B, device=dev, host_io=False, seed=args.seed, env_offset=rank * B)
    nA = env.nActuator
    replay = EfficientExperienceReplay((nA, nA), (nA, nA), max_size=args.replay_episodes * args.max_ts * B, device=dev, n_envs=B)
    dynamics = EnsembleDynamics(env.xvalid, env.yvalid, args.n_history).to(dev)
    policy = ConvPolicy(env.xvalid, env.yvalid, args.initial_sigma, env.F.float(), args.n_history).to(dev)
    dyn_opt, pol_opt = optim.Adam(dynamics.parameters()), optim.Adam(policy.parameters())
    torch.manual_seed(args.seed + 1000 * rank)         # different replay windows on every rank
    sigma = args.initial_sigma
    past_obs = past_act = obs = None
    history = []
    for i in range(args.iters):
        t0 = time.time()
        if len(replay) + args.max_ts * B > replay.max_size:          # keep the most recent episodes
            keep = replay.max_size - args.max_ts * B
            for buf in (replay.states, replay.next_states, replay.actions, replay.rewards):
                buf[:keep] = buf[len(replay) - keep:len(replay)].clone()
            replay.len = keep
        strehl, reward_sum, past_obs, past_act, obs, rewards, _ = mbrl.run(
            env, past_obs, past_act, obs, replay, policy, dynamics, args.n_history, args.max_ts, args.warmup_ts, sigma,
            episode=i, iteration=i)
        dyn_loss = pol_loss = 0.0
        if i == args.warmup_ts - 1:                                   # mbrl_main.py:118-123
            dyn_loss = mbrl.train_dynamics(args.n_history, args.max_ts, args.batch_size, dynamics, dyn_opt, replay, dyn_iters=100, device=dev)
            pol_loss = mbrl.train_policy(pol_opt, policy, dynamics, replay, dev, args.n_history, args.max_ts, args.batch_size, args.T, pol_iters=60)
        elif i > args.warmup_ts - 1:                                  # :124-133
            dyn_loss = mbrl.train_dynamics(args.n_history, args.max_ts, args.batch_size, dynamics, dyn_opt, replay, dyn_iters=10, device=dev)
            pol_loss = mbrl.train_policy(pol_opt, policy, dynamics, replay, dev, args.n_history, args.max_ts, args.batch_size, args.T, pol_iters=7)
        sigma = max(0.0, sigma - args.initial_sigma / args.warmup_ts)  # :146-147
        rs = float(torch.as_tensor(reward_sum, dtype=torch.float32).mean())
        history.append(dict(episode=i, strehl=strehl, reward_sum=rs, dyn_loss=dyn_loss, pol_loss=pol_loss, seconds=time.time() - t0))
        if verbose and rank == 0:
            print(f"episode {i}: {time.time() - t0:.2f}s  strehl {strehl:.4f}  reward {rs:.2f}  dyn {dyn_loss:.4f}  pol {pol_loss:.4f}", flush=True)
    return history, policy, dynamics


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    for k, v in vars(default_args()).items():
        if isinstance(v, (int, float, str)) and not isinstance(v, bool):
            ap.add_argument("--" + k.replace("_", "-"), type=type(v), default=v)
    ns = ap.parse_args()
    main(default_args(**{k: v for k, v in vars(ns).items()}))
    if dist.is_initialized():
        dist.destroy_process_group()

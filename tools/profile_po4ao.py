"""Operator-level profile of a PO4AO rollout (torch.profiler): where the time of mbrl.run goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from rlao_b200.PO4AO import mbrl
from rlao_b200.PO4AO.conv_models_simple import ConvPolicy, EnsembleDynamics
from rlao_b200.PO4AO.util_simple import EfficientExperienceReplay, TorchWrapper
from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO

dev = torch.device("cuda:0")
B, nH, steps = int(os.environ.get("ENVS", 1024)), 20, 20
base = OOPAO()
base.set_params_file("rlao_b200.Conf.parameter_file_synthetic_SHWFS", "")
base.set_params(bench.make_args(20, 1), "shackhartmann", gainCL=0.5, n_envs=B, device=dev)
env = TorchWrapper(base, host_io=False)
nA = env.nActuator
policy = ConvPolicy(env.xvalid, env.yvalid, 0.0, env.F.float(), nH).to(dev)
dynamics = EnsembleDynamics(env.xvalid, env.yvalid, nH).to(dev)
replay = EfficientExperienceReplay((nA, nA), (nA, nA), max_size=8 * steps * B, device=dev, n_envs=B)
run = lambda it: mbrl.run(env, None, None, None, replay, policy, dynamics, nH, steps, 0, 0.0, episode=1, iteration=it)
run(1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(2)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

# ---- wall-clock split of one step (host time of each piece, GPU idle or not) ----------------------------------
import time
obs = env.reset_soft().reshape(B, nA, nA)
h_obs = mbrl._History(torch.zeros((B, nH - 1, nA, nA), device=dev))
h_act = mbrl._History(torch.zeros((B, nH - 1, nA, nA), device=dev))
acc = {}
def tick(name, t0):
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
torch.cuda.synchronize()
T0 = time.perf_counter()
with torch.no_grad():
    for t in range(steps):
        t0 = time.perf_counter(); history = torch.cat([h_obs.window(), h_act.window()], dim=1); tick("cat", t0)
        t0 = time.perf_counter(); action = policy(obs.unsqueeze(1), history)[:, 0]; tick("policy", t0)
        t0 = time.perf_counter(); next_obs, reward, strehl, done, _ = env.step(t, action); tick("env.step", t0)
        t0 = time.perf_counter(); h_obs.push(obs); h_act.push(action); tick("push", t0)
        t0 = time.perf_counter(); replay.append(obs, action, reward, next_obs, done); tick("replay", t0)
        obs = next_obs
t_host = time.perf_counter() - T0
torch.cuda.synchronize()
t_all = time.perf_counter() - T0
print("host ms/step", 1e3 * t_host / steps, "wall ms/step", 1e3 * t_all / steps, {k: round(1e3 * v / steps, 3) for k, v in acc.items()})

# ---- mbrl.run itself, timed with events -----------------------------------------------------------------------
for n in (20, 50):
    replay.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record()
    mbrl.run(env, None, None, None, replay, policy, dynamics, nH, n, 0, 0.0, episode=1, iteration=3, new_screen=False)
    e1.record()
    torch.cuda.synchronize()
    print("mbrl.run", n, "steps:", e0.elapsed_time(e1) / n, "ms/step (events)", 1e3 * (time.perf_counter() - w0) / n, "ms/step (wall)")

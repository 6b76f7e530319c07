import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rlao_b200.PO4AO import mbrl
from rlao_b200.PO4AO.conv_models_simple import ConvPolicy, EnsembleDynamics
from rlao_b200.PO4AO.util_simple import EfficientExperienceReplay, TorchWrapper
from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO
dev = torch.device("cuda:0")
B, nH = 1024, 20
base = OOPAO(); base.set_params_file("rlao_b200.Conf.parameter_file_synthetic_SHWFS", "")
base.set_params(bench.make_args(20, 1), "shackhartmann", gainCL=0.5, n_envs=B, device=dev)
env = TorchWrapper(base, host_io=False); nA = env.nActuator
policy = ConvPolicy(env.xvalid, env.yvalid, 0.0, env.F.float(), nH).to(dev)
dynamics = EnsembleDynamics(env.xvalid, env.yvalid, nH).to(dev)
replay = EfficientExperienceReplay((nA, nA), (nA, nA), max_size=200 * B, device=dev, n_envs=B)
def timed(label, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mbrl.run(env, None, None, None, replay, policy, dynamics, nH, 20, 0, 0.0, episode=1, **kw)
    torch.cuda.synchronize(); print(label, "%.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
timed("first, new screen", iteration=1)
timed("second, new screen", iteration=2)
pr = cProfile.Profile(); pr.enable()
timed("third, no new screen", iteration=2, new_screen=False)
pr.disable()
timed("fourth, no new screen", iteration=2, new_screen=False)
timed("fifth, new screen", iteration=3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# ---- the bench flow: 5-step warm-up episode, then a 50-step episode without new screens, host time between env.step calls
replay.clear()
mbrl.run(env, None, None, None, replay, policy, dynamics, nH, 5, 0, 0.0, episode=1, iteration=1)
replay.clear()
torch.cuda.synchronize()
stamps = []
orig = env.step
def step(i, a):
    stamps.append(time.perf_counter())
    return orig(i, a)
env.step = step
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
mbrl.run(env, None, None, None, replay, policy, dynamics, nH, 50, 0, 0.0, episode=1, iteration=2, new_screen=False)
e1.record(); torch.cuda.synchronize()
print("50 steps: events %.1f ms, wall %.1f ms" % (e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
d = [1e3 * (b - a) for a, b in zip(stamps[:-1], stamps[1:])]
print("first call after %.1f ms; host intervals between env.step calls (ms):" % (1e3 * (stamps[0] - t0)), [round(x, 2) for x in d])

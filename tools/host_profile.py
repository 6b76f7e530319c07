"""cProfile of the host side of env.step (single environment: the launch-bound case)."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO
dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
nS, nL, B, desc, opts = bench.WORKLOADS[wl]
env = OOPAO(); env.set_params_file("rlao_b200.Conf.parameter_file_synthetic_SHWFS", "")
env.set_params(bench.make_args(nS, nL, opts), "shackhartmann", gainCL=0.5, n_envs=B, device=dev)
env.atm.generateNewPhaseScreen(17); env.dm.coefs = 0; env.tel * env.dm * env.wfs
obs = env.reset_soft()
for i in range(50):
    obs, *_ = env.step(None, 0.5 * obs)
torch.cuda.synchronize()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
for i in range(500):
    obs, *_ = env.step(None, 0.5 * obs)
pr.disable(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("host ms/step (with profiler overhead):", 1e3 * (t1 - t0) / 500)
pstats.Stats(pr).sort_stats("tottime").print_stats(28)

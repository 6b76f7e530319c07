"""Per-step wall time of the environment over a long run without new screens: finds periodic stalls."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO

dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
nS, nL, B, desc, opts = bench.WORKLOADS[wl]
env = OOPAO()
env.set_params_file("rlao_b200.Conf.parameter_file_synthetic_SHWFS", "")
env.set_params(bench.make_args(nS, nL, opts), "shackhartmann", gainCL=0.5, n_envs=B, device=dev)
env.atm.generateNewPhaseScreen(17); env.dm.coefs = 0; env.tel * env.dm * env.wfs
obs = env.reset_soft()
times = []
for t in range(300):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    obs, *_ = env.step(t, 0.5 * obs)
    torch.cuda.synchronize(); times.append(1e3 * (time.perf_counter() - t0))
import numpy as np
a = np.array(times)
print(wl, "median %.3f ms  mean %.3f  p99 %.3f  max %.3f at step %d" % (np.median(a), a.mean(), np.percentile(a, 99), a.max(), a.argmax()))
print("steps > 2x median:", [(i, round(x, 2)) for i, x in enumerate(a) if x > 2 * np.median(a)][:40])

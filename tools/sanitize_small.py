"""A few steps of a tiny environment (4 x 4 lenslets, two environments, noisy camera), the staged camera chain, the PSF
image and one Pyramid frame — the run `compute-sanitizer --tool racecheck|memcheck python tools/sanitize_small.py` checks.
AOENV_GEMM=simt keeps the tcgen05 / TMA GEMM (which the sanitizer's shared-memory tracking does not model) out of the run;
the default exercises it as well."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO


def main():
    dev = torch.device("cuda:0")
    env = OOPAO()
    env.set_params_file("rlao_b200.Conf.parameter_file_synthetic_SHWFS", "")
    args = bench.make_args(4, 2, {"noise": True, "magnitude": 8})
    env.set_params(args, "shackhartmann", gainCL=0.5, n_envs=2, device=dev)
    env.atm.pipelined = "force"                      # the side-stream prefetch as well
    env.atm.generateNewPhaseScreen(3)
    obs = env.reset_soft()
    for i in range(4):
        obs, reward, strehl, done, info = env.step(None, 0.5 * obs)
    img = env.sample_noise(1e-8)
    psf = env.tel.computePSF(2)
    torch.cuda.synchronize()
    print("ok", float(obs.abs().max()), float(img.abs().max()))


if __name__ == "__main__":
    main()

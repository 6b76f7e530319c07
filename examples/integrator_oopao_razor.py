"""The integrator loop of MAIN_CODE/integrator_oopao_razor.py:36-90 on the batched GPU environment: same calls, the
imports are the only change.  Sweeps r0 and wind speed like the reference script and prints the episode
statistics every 500 frames.

    python examples/integrator_oopao_razor.py [--n-envs 1024] [--n-loop 2000] [--n-subap 20]
"""
import argparse
import os
import sys
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rlao_b200.PO4AO.mbrl import get_env                      # reference: from PO4AO.mbrl_funcsRAZOR import get_env
from rlao_b200.Conf.parameter_file_synthetic_SHWFS import layer_profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=1024)
    ap.add_argument("--n-loop", type=int, default=2000)
    ap.add_argument("--n-subap", type=int, default=20)
    ap.add_argument("--layers", type=int, default=5)
    cli = ap.parse_args()
    prof = layer_profile(cli.layers)                          # drl4ao's Conf/papyrus_config.yaml:13-15 profile
    args = SimpleNamespace(param_file="rlao_b200.Conf.parameter_file_synthetic_SHWFS", oopao_path="", delay=0, L0=25,
                           nSubaperture=cli.n_subap, nLoop=cli.n_loop, **prof)
    for r0 in [0.13, 0.0866666667]:
        args.r0 = r0
        env = get_env(args, gainCL=0.9, n_envs=cli.n_envs)    # TorchWrapper(host_io=False): tensors stay on the GPU
        for ws in [prof["windSpeed"], [2 * v for v in prof["windSpeed"]]]:
            env.atm.windSpeed = ws
            env.atm.generateNewPhaseScreen(17)
            env.dm.coefs = 0
            env.tel * env.dm * env.wfs
            obs = env.reset_soft()
            accu_reward, t0 = 0.0, time.time()
            for i in range(args.nLoop):
                action = env.gainCL * obs
                obs, reward, strehl, done, info = env.step(i, action)
                accu_reward = accu_reward + reward
                if (i + 1) % 500 == 0:
                    sr = env.calculate_strehl_AVG()
                    dt = time.time() - t0
                    print(f"r0 {r0:.3f} wind {ws}: frame {i + 1}/{args.nLoop}  mean SR {sr:.4f}  "
                          f"turbulence {float(env.total[i].mean()):.1f} nm  residual {float(env.residual[i].mean()):.1f} nm  "
                          f"mean reward {float(accu_reward.mean()) / 500:.3f}  "
                          f"{cli.n_envs * 500 / dt:,.0f} env-steps/s", flush=True)
                    accu_reward, t0 = 0.0, time.time()


if __name__ == "__main__":
    main()

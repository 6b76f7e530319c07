"""A/B of the SM partition (rlao_b200/sm_partition.py) on the default workload: whole step, device-resident loop.
Usage: python tools/bench_partition.py [workload] [steps] [side_sms ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rlao_b200 import sm_partition  # noqa: E402
from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO  # noqa: E402


def build(workload, dev):
    nS, nL, B, desc, opts = bench.WORKLOADS[workload]
    env = OOPAO()
    env.set_params_file("rlao_b200.Conf.parameter_file_synthetic_SHWFS", "")
    env.set_params(bench.make_args(nS, nL, opts), "shackhartmann", gainCL=0.5, n_envs=B, device=dev, rng="philox", seed=1)
    env.atm.generateNewPhaseScreen(17)
    env.dm.coefs = 0
    env.tel * env.dm * env.wfs
    torch.cuda.synchronize()
    return env, B


def run(env, B, steps, stream):
    dev = env.device
    with torch.cuda.stream(stream):
        obs = env.reset_soft()
        for _ in range(10):
            obs, *_ = env.step(None, env.gainCL * obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            obs, reward, strehl, _, _ = env.step(None, env.gainCL * obs)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, B / ms * 1e3, float(strehl.float().mean()), obs.clone()


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    sides = [int(v) for v in sys.argv[3:]] or [32, 40, 48, 56, 64]
    dev = torch.device("cuda:0")
    out = []

    def fresh():
        env, B = build(workload, dev)
        return env, B

    env, B = fresh()
    ms, v, sr, ref = run(env, B, steps, torch.cuda.current_stream(dev))
    out.append({"mode": "two plain streams, atmosphere behind the slopes (default)", "ms_per_step": ms, "value": v, "strehl": sr})
    print(out[-1], flush=True)
    env, B = fresh()
    env.prefetch_early = True
    ms, v, sr, o = run(env, B, steps, torch.cuda.current_stream(dev))
    out.append({"mode": "two plain streams, atmosphere issued before the spots", "ms_per_step": ms, "value": v, "strehl": sr,
                "bit_identical_to_default": bool(torch.equal(o, ref))})
    print(out[-1], flush=True)
    for side in sides:
        try:
            part = sm_partition.get(dev, side)
        except Exception as ex:
            out.append({"mode": f"green contexts, {side} SMs", "error": repr(ex)})
            print(out[-1], flush=True)
            continue
        env, B = fresh()
        env.atm.sm_partition = part
        ms, v, sr, o = run(env, B, steps, part.main)
        out.append({"mode": part.describe(), "ms_per_step": ms, "value": v, "strehl": sr,
                    "bit_identical_to_default": bool(torch.equal(o, ref))})
        print(out[-1], flush=True)
        # the sensor chain on its share, the atmosphere on an ordinary stream (control: is it the partition or the order?)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

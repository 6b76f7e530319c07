"""Times the calls around the step that round 2 moved onto the device (CUDA events, one B200): episode reset
(Atmosphere.generateNewPhaseScreen -> aoenv_vk_screens), the exploration-noise draw (env.sample_noise), the full PSF image
(Telescope.computePSF -> aoenv_psf_image) and the checkpoint round trip.  Usage: python tools/bench_aux.py [workload] [envs]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    nS, nL, B, desc, opts = bench.WORKLOADS[wl]
    if len(sys.argv) > 2:
        B = int(sys.argv[2])
    dev = torch.device("cuda:0")
    env = OOPAO()
    env.set_params_file("rlao_b200.Conf.parameter_file_synthetic_SHWFS", "")
    env.set_params(bench.make_args(nS, nL, opts), "shackhartmann", gainCL=0.5, n_envs=B, device=dev)
    out = {"workload": wl, "envs": B, "layers": nL, "resolution": env.tel.resolution}
    seed = [100]

    def reset():
        seed[0] += 1
        env.atm.generateNewPhaseScreen(seed[0])
    out["reset_ms"] = timed(reset)
    out["reset_screens_per_s"] = B * nL / out["reset_ms"] * 1e3
    t0 = time.perf_counter()
    env.reset_soft()
    torch.cuda.synchronize()
    out["reset_soft_wall_ms"] = (time.perf_counter() - t0) * 1e3
    out["sample_noise_ms"] = timed(lambda: env.sample_noise(1e-8), reps=20)
    obs = env.reset_soft()
    for _ in range(3):
        obs, *_ = env.step(None, 0.5 * obs)
    nb = min(B, 64)
    from rlao_b200 import psf
    for zp in (2, 4):
        try:
            a = env.atm.OPD_no_pupil[:nb].contiguous()
            out[f"psf_image_zp{zp}_ms_per_{nb}"] = timed(lambda: psf.psf_image(env.tel, a, None, zp))
            out[f"psf_peak_zp{zp}_ms_per_{nb}"] = timed(lambda: psf.psf_peak(env.tel, a, None, zp))
        except Exception as e:                              # noqa: BLE001 (report, keep timing the rest)
            out[f"psf_zp{zp}_error"] = repr(e)
    t0 = time.perf_counter()
    sd = env.state_dict()
    torch.cuda.synchronize()
    out["state_dict_wall_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    env.load_state_dict(sd)
    torch.cuda.synchronize()
    out["load_state_dict_wall_ms"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()

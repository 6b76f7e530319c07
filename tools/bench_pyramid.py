"""Times the Pyramid frame computation alone: the library's FFT kernels against the torch.fft (cuFFT) path.
Usage: python tools/bench_pyramid.py [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlao_b200.Pyramid import Pyramid
from rlao_b200.Source import Source
from rlao_b200.Telescope import Telescope

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
tel = Telescope(120, 8.0, 1 / 500, n_envs=B, device=dev)
Source("I", 8) * tel
wfs = Pyramid(20, tel, 3, 0.1, n_pix_separation=4, n_pix_edge=2)
a = (torch.randn((B, 120, 120), device=dev) * 1e-7).contiguous()
ph = a * tel._pupil_f * (2 * 3.141592653589793 / tel.src.wavelength)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


print(f"pyramid 20x20, N={wfs.nRes}, nTheta={wfs.nTheta}, envs={B}")
print(f"  kernels : {timed(lambda: wfs._frames_kernels(a, None)):8.3f} ms")
print(f"  cuFFT   : {timed(lambda: wfs._frames(ph)):8.3f} ms")

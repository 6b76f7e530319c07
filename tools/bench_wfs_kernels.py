"""Times the kernels of the production WFS chain alone (CUDA events): DM surface, frame (every n = 6 variant), slopes.
Usage: python tools/bench_wfs_kernels.py [nS] [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlao_b200 import _lib
from rlao_b200.DeformableMirror import DeformableMirror
from rlao_b200.ShackHartmann import ShackHartmann
from rlao_b200.Source import Source
from rlao_b200.Telescope import Telescope


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps * 1e3


nS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
R = nS * 6
tel = Telescope(R, 8.0, 1 / 500, n_envs=B, device=dev)
Source("I", 8) * tel
wfs = ShackHartmann(nS, tel, 0.5)
dm = DeformableMirror(tel, nS, 0.35)
g = torch.Generator(device=dev).manual_seed(1)
opd = (torch.randn((B, R, R), device=dev, generator=g) * 1e-7).contiguous()
coefs = torch.zeros((B, dm._Kp), device=dev)
coefs[:, :dm.nValidAct] = torch.randn((B, dm.nValidAct), device=dev, generator=g) * 1e-7
lib = _lib.load()
print(f"nS={nS} envs={B}")
print(f"  dm surface            {timed(lambda: dm._surface(coefs, dm._opd[0])):8.1f} us")
for variant, name in ((3, "term by term, 3 lanes"), (2, "factorised, 3 lanes"), (1, "factorised, 1 thread"), (0, "term by term, 1 thread")):
    lib.aoenv_set_wfs6_variant(variant)
    t = timed(lambda: wfs._measure_terms(opd, dm._opd[0], 0))
    print(f"  frame + slopes [{name:22s}] {t:8.1f} us")
lib.aoenv_set_wfs6_variant(-1)

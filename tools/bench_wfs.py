"""Times the wavefront-sensing part of the step alone (CUDA events): the unfused chain (DM surface kernel + frame kernel +
slopes kernel) against aoenv_shwfs_fused for several cluster / warp-group shapes.  Usage:
    python tools/bench_wfs.py [nS] [envs]            (default 40 1024: the benchmark shape)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


def main():
    nS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    n, dev = 6, torch.device("cuda:0")
    R = nS * n
    tel = Telescope(R, 8.0, 1 / 500, n_envs=B, device=dev)
    Source("I", 8) * tel
    wfs = ShackHartmann(nS, tel, 0.5)
    dm = DeformableMirror(tel, nS, 0.35)
    dm.lazy_surface = True
    g = torch.Generator(device=dev).manual_seed(1)
    opd = (torch.randn((B, R, R), device=dev, generator=g) * 1e-7).contiguous()
    coefs = torch.zeros((B, dm._Kp), device=dev)
    coefs[:, :dm.nValidAct] = torch.randn((B, dm.nValidAct), device=dev, generator=g) * 1e-7
    dm._set_coefs_batch(coefs)
    ref = dm.surface_ref()

    def unfused():
        dm._surface(coefs, dm._opd[0])
        wfs.use_fused = False
        wfs._measure_terms(opd, dm._opd[0], 0)
    print(f"nS={nS} envs={B}: unfused (dm + frame + slopes) {timed(unfused):8.1f} us")
    divs = [2, 4, 5, 6, 8, 10, 12, 16]
    for keep in (False, True):
        for cluster in divs:
            for groups in (2, 4, 6):
                os.environ["AOENV_WFS_CLUSTER"], os.environ["AOENV_WFS_GROUPS"] = str(cluster), str(groups)
                wfs._fused_plans = {}
                wfs.use_fused, wfs.keep_frame = True, keep
                try:
                    t = timed(lambda: wfs._measure_terms(opd, ref, 0))
                except Exception as e:                                     # does not fit / not compiled
                    print(f"  fused cluster={cluster:2d} groups={groups} frame={int(keep)}: {str(e)[:90]}")
                    continue
                print(f"  fused cluster={cluster:2d} groups={groups} frame={int(keep)}: {t:8.1f} us")


if __name__ == "__main__":
    main()

"""A few launches of aoenv_shwfs_fused at the benchmark shape (for `ncu -k regex:shwfs_fused`).
Usage: python tools/prof_wfs.py [nS] [envs] [cluster] [groups] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
if len(sys.argv) > 3:
    os.environ["AOENV_WFS_CLUSTER"] = sys.argv[3]
if len(sys.argv) > 4:
    os.environ["AOENV_WFS_GROUPS"] = sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 6
from rlao_b200.DeformableMirror import DeformableMirror
from rlao_b200.ShackHartmann import ShackHartmann
from rlao_b200.Source import Source
from rlao_b200.Telescope import Telescope

dev = torch.device("cuda:0")
R = nS * 6
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    wfs._measure_terms(opd, ref, 0)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    wfs._measure_terms(opd, ref, 0)
e1.record()
torch.cuda.synchronize()
print(f"fused nS={nS} envs={B} plan={ {k: v for k, v in next(iter(wfs._fused_plans.values())).items() if k in ('cluster', 'groups', 't_rows')} }: "
      f"{e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch")

"""InteractionMatrix — mirror of OOPAO/calibration/InteractionMatrix.py:13-135.

Same protocol as the reference (`dm.coefs = M2C[:, chunk]*stroke; tel*dm; tel*wfs`), `nMeasurements` commands
at a time through the multi-frame WFS branch, which on the GPU is one DM-surface GEMM and one batched WFS
launch per chunk (the chunk size is kept because the centroiding threshold is global over a chunk,
ShackHartmann.py:659 -> :314-316)."""
import math

import numpy as np
import torch

from .CalibrationVault import CalibrationVault


def InteractionMatrix(ngs, atm, tel, dm, wfs, M2C, stroke, phaseOffset=0, nMeasurements=50, noise="off", invert=True,
                      print_time=False, display=False, single_pass=True):
    if np.ndim(phaseOffset) != 0 or phaseOffset != 0:
        raise NotImplementedError("phaseOffset is not supported")
    saved = (wfs.cam.photonNoise, wfs.cam.readoutNoise, wfs.cam.backgroundNoise)
    if noise == "off":
        wfs.cam.photonNoise = 0
        wfs.cam.readoutNoise = 0
        wfs.cam.backgroundNoise = 0
    else:
        print("Warning: Keeping the noise configuration for the WFS")
    tel.isPaired = False
    ngs * tel
    M2C = torch.as_tensor(M2C, dtype=torch.float64, device=tel.device)
    nModes = M2C.shape[1] if M2C.ndim == 2 else 1
    M2C = M2C.reshape(M2C.shape[0], nModes)
    intMat = torch.zeros((wfs.nSignal, nModes), dtype=torch.float64, device=tel.device)
    nMeasurements = min(nMeasurements, nModes)
    nCycle = int(math.ceil(nModes / nMeasurements))
    nExtra = nModes % nMeasurements
    for i in range(nCycle):
        if i == nCycle - 1 and nExtra != 0:
            cols = slice(nModes - nExtra, nModes)
        else:
            cols = slice(i * nMeasurements, (i + 1) * nMeasurements)
        cmd = M2C[:, cols] * stroke

        def push(c):
            if c.shape[1] == 1:
                # a single command goes through the single-frame branch, detector included (ndim(OPD) == 2)
                dm.coefs = c[:, 0]
                tel * dm
                tel * wfs
                return wfs._signal[0, :wfs.nSignal].double()[:, None]
            dm.coefs = c
            tel * dm
            return wfs.measure_frames(tel._materialise()).T.double()

        sp = push(cmd)
        if single_pass:
            sm, factor = 0 * sp, 2
        else:
            sm, factor = push(-cmd), 1
        intMat[:, cols] = 0.5 * (sp - sm) / stroke
    wfs.cam.photonNoise, wfs.cam.readoutNoise, wfs.cam.backgroundNoise = saved if noise != "off" else (0, 0, 0)
    return CalibrationVault(factor * intMat, invert=invert)

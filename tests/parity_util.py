"""Shared helpers for the parity tests."""
import types

import numpy as np


def param_from_config(cfg):
    """oracle.ao_oracle.AOConfig -> parameter dict in the format rlao_b200's env mirror reads
    (same keys as the reference's parameter files, see rlao_b200/Conf/parameter_file_synthetic_SHWFS.py)."""
    d = cfg.detector
    p = dict(
        r0=cfg.r0, L0=cfg.L0, fractionalR0=list(cfg.fractionalR0), windSpeed=list(cfg.windSpeed),
        windDirection=list(cfg.windDirection), altitude=list(cfg.altitude), diameter=cfg.diameter,
        nSubaperture=cfg.nSubap, nPixelPerSubap=cfg.nPixPerSubap, resolution=cfg.resolution,
        samplingTime=cfg.samplingTime, centralObstruction=cfg.centralObstruction, magnitude=cfg.magnitude,
        opticalBand=cfg.opticalBand, nActuator=cfg.nSubap + 1, mechanicalCoupling=cfg.mechCoupling, isM4=False,
        dm_geometry="cartesian", shiftX=0, shiftY=0, rotationAngle=0, anamorphosisAngle=0, radialScaling=0,
        tangentialScaling=0, lightRatio=cfg.lightRatio, threshold_cog=cfg.threshold_cog, shannon_sampling=False,
        cam_photonNoise=d.photonNoise, cam_readoutNoise=d.readoutNoise, cam_sensor=d.sensor, cam_FWC=d.FWC,
        cam_bits=d.bits, cam_QE=d.QE, cam_darkCurrent=d.darkCurrent, nZernike=cfg.nZernike,
        nMeasurements=cfg.nMeasurements, nLoop=cfg.nLoop, gainCL=cfg.gainCL)
    return p


def build_env(cfg, n_envs=1, rng="reference", seed=0, device=None, env_offset=0, canvas_slack=32):
    from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO
    env = OOPAO()
    env.set_params_file(param_from_config(cfg), "")
    env.set_params(types.SimpleNamespace(), "shackhartmann", gainCL=cfg.gainCL, n_envs=n_envs, rng=rng, seed=seed,
                   device=device, env_offset=env_offset, canvas_slack=canvas_slack)
    env.leak = cfg.leak
    return env


def new_episode(env, seed):
    """Caller pattern of MAIN_CODE/integrator_oopao_razor.py:46-60 / PO4AO/mbrl.py:49-55."""
    env.atm.generateNewPhaseScreen(seed)
    env.dm.coefs = 0
    env.dm_prev = 0 * env.dm_prev
    env.tel * env.dm * env.wfs
    return env.reset_soft()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)

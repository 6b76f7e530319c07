"""Runs the UNMODIFIED reference (OOPAO + drl4ao env) inside the build container — TEST INFRASTRUCTURE ONLY.

/root/reference exists only in the build container, so nothing under tests/ -m gpu, smoke() or bench.py may
import this module; it is used by oracle/make_golden.py (fixture generation) and by the optional
`tests/test_oracle_vs_reference.py`, which skips itself when /root/reference is absent.
"""
import contextlib
import io
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "ref_shims")
# The reference lives in /root/reference in the build container.  __graft_entry__.build() stages an UNMODIFIED copy of
# the files of the path (OOPAO package + the Razor environment) under the git-ignored oracle/_ref/, which travels to
# the GPU box with the built library, so that bench.py's CPU arm can time the reference itself there.
STAGED_ROOT = os.path.join(_HERE, "_ref", "drl4ao")
REF_ROOT = ("/root/reference/drl4ao" if os.path.isdir("/root/reference/drl4ao/AO_OOPAO/OOPAO")
            and os.environ.get("AOENV_REF_ROOT") != "staged" else STAGED_ROOT)
STAGED_FILES = ("AO_OOPAO/OOPAO", "MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py", "MAIN_CODE/OOPAOEnv/__load__oopao.py")


def stage_reference(src_root="/root/reference/drl4ao"):
    """Copies the reference files of the path, byte for byte, into oracle/_ref/ (build-time; outputs only there)."""
    import shutil
    if not os.path.isdir(os.path.join(src_root, "AO_OOPAO", "OOPAO")):
        return False
    for rel in STAGED_FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(STAGED_ROOT, rel)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.npy", "*.fits", "*.png", "*.jpg"))
        elif os.path.exists(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return True


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "AO_OOPAO", "OOPAO"))


def _prepare_imports():
    if not hasattr(np, "math"):
        np.math = math          # numpy >= 2 dropped np.math; OOPAO/phaseStats.py:21,76,78 still uses it
    for p in (_SHIMS, os.path.join(REF_ROOT, "AO_OOPAO"), os.path.join(REF_ROOT, "MAIN_CODE")):
        if p not in sys.path:
            sys.path.insert(0, p)


@contextlib.contextmanager
def quiet():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


def build_reference_env(cfg, verbose=False):
    """Builds the reference optical train for the synthetic SH configuration `cfg` (oracle.ao_oracle.AOConfig)
    in the order of MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:set_params, then hosts it in the reference's own
    `OOPAO` gym class so that its real `step()` / `reset_soft()` run on top."""
    _prepare_imports()
    ctx = contextlib.nullcontext() if verbose else quiet()
    with ctx:
        from OOPAO.Telescope import Telescope
        from OOPAO.Source import Source
        from OOPAO.Atmosphere import Atmosphere
        from OOPAO.DeformableMirror import DeformableMirror
        from OOPAO.ShackHartmann import ShackHartmann
        from OOPAO.Zernike import Zernike
        from OOPAO.calibration.InteractionMatrix import InteractionMatrix
        from OOPAO.calibration.CalibrationVault import CalibrationVault
        from OOPAOEnv.OOPAOEnvRazor import OOPAO

        tel = Telescope(resolution=cfg.resolution, diameter=cfg.diameter, samplingTime=cfg.samplingTime,
                        centralObstruction=cfg.centralObstruction)
        src = Source(optBand=cfg.opticalBand, magnitude=cfg.magnitude)
        src * tel
        atm = Atmosphere(telescope=tel, r0=cfg.r0, L0=cfg.L0, windSpeed=list(cfg.windSpeed),
                         fractionalR0=list(cfg.fractionalR0), windDirection=list(cfg.windDirection),
                         altitude=list(cfg.altitude))
        atm.initializeAtmosphere(tel)
        atm.update()
        tel + atm
        assert cfg.dm_geometry == "cartesian"
        dm = DeformableMirror(telescope=tel, nSubap=cfg.nSubap, mechCoupling=cfg.mechCoupling)
        tel - atm
        wfs = ShackHartmann(telescope=tel, nSubap=cfg.nSubap, lightRatio=cfg.lightRatio,
                            threshold_cog=cfg.threshold_cog, is_geometric=False, shannon_sampling=False)
        d = cfg.detector
        wfs.cam.sensor = d.sensor
        wfs.cam.FWC = d.FWC
        wfs.cam.bits = d.bits
        wfs.cam.QE = d.QE
        wfs.cam.gain = d.gain
        wfs.cam.darkCurrent = d.darkCurrent
        wfs.cam.integrationTime = cfg.samplingTime
        tel * wfs
        if cfg.nZernike > 0:
            Z = Zernike(tel, cfg.nZernike)
            Z.computeZernike(tel)
            M2C = np.linalg.pinv(np.squeeze(dm.modes[tel.pupilLogical, :])) @ Z.modes
        else:
            M2C = np.eye(dm.nValidAct)
        calib_zonal = InteractionMatrix(ngs=src, atm=atm, tel=tel, dm=dm, wfs=wfs, M2C=np.eye(dm.nValidAct),
                                        stroke=cfg.stroke, nMeasurements=cfg.nMeasurements, noise="off")
        calib = CalibrationVault(calib_zonal.D @ M2C)
        tel.resetOPD()
        dm.coefs = 0
        env = OOPAO()
        env.tel, env.source, env.atm, env.dm, env.wfs = tel, src, atm, dm, wfs
        env.dm_prev = dm.coefs.copy()
        src * tel * dm * wfs
        tel + atm
        env.nActuator = cfg.nSubap + 1
        env.dm_mask = np.reshape(dm.validAct, (env.nActuator, env.nActuator)).astype(int)
        env.xvalid, env.yvalid = np.nonzero(env.dm_mask)
        env.M2C_CL = M2C
        env.calib_zonal = calib_zonal
        env.calib_CL = calib
        env.reconstructor = M2C @ calib.M
        env.F = M2C @ np.linalg.pinv(M2C)
        env.SR = []
        env.total = np.zeros(cfg.nLoop)
        env.residual = np.zeros(cfg.nLoop)
        env.leak = cfg.leak
        env.gainCL = cfg.gainCL
        wfs.cam.photonNoise = d.photonNoise
        wfs.cam.readoutNoise = d.readoutNoise
    return env


class ReferenceStepper:
    """The reference environment behind the two calls bench.py's CPU arm uses (new_episode / step), following the caller
    pattern of MAIN_CODE/integrator_oopao_razor.py:46-70."""

    def __init__(self, cfg):
        self.env = build_reference_env(cfg)
        self.gainCL = cfg.gainCL

    def new_episode(self, seed):
        env = self.env
        with quiet():
            env.atm.generateNewPhaseScreen(seed)
            env.dm.coefs = 0
            env.dm_prev = env.dm.coefs.copy()
            env.tel * env.dm * env.wfs
            return env.reset_soft()

    def step(self, i, action):
        with quiet():
            return self.env.step(i, action)


def record_xi(env):
    """Wraps each layer's RandomState.normal so that every innovation vector drawn by
    OOPAO/Atmosphere.py:308,582 is appended to the returned list."""
    log = []
    for i in range(env.atm.nLayer):
        layer = getattr(env.atm, "layer_" + str(i + 1))
        _wrap_layer_rng(layer, log)
    return log


def _wrap_layer_rng(layer, log):
    rs = layer.randomState

    class _Recorder:
        def normal(self, *a, **k):
            out = rs.normal(*a, **k)
            log.append(np.array(out, copy=True))
            return out

        def __getattr__(self, name):
            return getattr(rs, name)

    layer.randomState = _Recorder()

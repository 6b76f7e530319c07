"""Named configurations shared by oracle/make_golden.py and the tests (TEST INFRASTRUCTURE ONLY)."""
from .ao_oracle import AOConfig, DetectorConfig

RAZOR_DETECTOR = dict(photonNoise=True, readoutNoise=14.0, darkCurrent=5.0, QE=0.56, FWC=10000, bits=10,
                      sensor="CMOS")        # MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:243-250,332-333


def tiny():
    return AOConfig(nSubap=8, windSpeed=[10.0, 12.0], windDirection=[0.0, 72.0], fractionalR0=[0.6, 0.4],
                    altitude=[0.0, 0.0], nZernike=20, nLoop=64)


def tiny_noise():
    c = tiny()
    c.magnitude = 6.0
    c.detector = DetectorConfig(**RAZOR_DETECTOR)
    return c


def cfg1():
    """BASELINE.json configs[0]: 8 m, 20x20 SH, 21x21 DM, one layer (SURVEY.md section 8 d)."""
    return AOConfig(nSubap=20, windSpeed=[10.0], windDirection=[0.0], fractionalR0=[1.0], altitude=[0.0],
                    nZernike=50, nLoop=64)


def cfg3():
    """BASELINE.json configs[2], the benchmark configuration: 8 m, 40x40 SH, 41x41 DM, three layers (SURVEY.md section
    8 d; one environment of the 8192)."""
    return AOConfig(nSubap=40, windSpeed=[10.0, 12.0, 11.0], windDirection=[0.0, 72.0, 144.0],
                    fractionalR0=[0.45 / 0.65, 0.1 / 0.65, 0.1 / 0.65], altitude=[0.0, 0.0, 0.0], nZernike=50, nLoop=64)


def cfg5():
    """BASELINE.json configs[4]: ELT-scale 80x80 SH, 81x81 DM (5209 actuators), five layers (SURVEY.md section 8 d; one
    environment of the 2048)."""
    return AOConfig(nSubap=80, windSpeed=[10.0, 12.0, 11.0, 15.0, 20.0], windDirection=[0.0, 72.0, 144.0, 216.0, 288.0],
                    fractionalR0=[0.45, 0.1, 0.1, 0.25, 0.1], altitude=[0.0] * 5, nZernike=50, nLoop=64)


def psf_formula_opd(R):
    """A smooth ~lambda/12 rms wavefront (metres) defined by a formula, the input of the large-size PSF fixture."""
    import numpy as np
    yy, xx = np.mgrid[:R, :R] / R
    return 60e-9 * (np.sin(5 * xx + 2 * yy) + 0.5 * np.cos(9 * yy * xx) + xx * yy)


CONFIGS = {"tiny": tiny, "tiny_noise": tiny_noise, "cfg1": cfg1, "cfg3": cfg3, "cfg5": cfg5}
STEPS = {"tiny": 30, "tiny_noise": 12, "cfg1": 24, "cfg3": 8, "cfg5": 5}
# fixtures of the large configurations keep float32 snapshots of the last step only (size)
COMPACT = {"cfg3", "cfg5"}
# the largest one additionally records what a test needs to run WITHOUT the CPU oracle beside it (whose set-up takes
# tens of minutes at this size): per-step knife-edge lenslet masks from the reference's own spots, every third row of
# the snapshots, no reconstructor
LITE = {"cfg5"}
EPISODE_SEED = 17          # MAIN_CODE/integrator_oopao_razor.py:46

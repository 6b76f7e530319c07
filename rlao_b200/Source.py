"""Source — mirror of OOPAO/Source.py (NGS only): photometry, `src*tel` coupling."""
import math

import numpy as np

# [wavelength m, bandwidth m, zero point ph/m2/s] — the photometric system of OOPAO/Source.py:164-242
_PHOTOMETRY = {
    "U": (0.360e-6, 0.070e-6, 1.96e12), "B": (0.440e-6, 0.100e-6, 5.38e12), "V0": (0.500e-6, 0.090e-6, 3.64e12),
    "V": (0.550e-6, 0.090e-6, 3.31e12), "R": (0.640e-6, 0.150e-6, 4.01e12), "R2": (0.650e-6, 0.300e-6, 7.9e12),
    "R3": (0.600e-6, 0.300e-6, 8.56e12), "R4": (0.670e-6, 0.300e-6, 7.66e12), "I": (0.790e-6, 0.150e-6, 2.69e12),
    "I1": (0.700e-6, 0.033e-6, 0.67e12), "I2": (0.750e-6, 0.033e-6, 0.62e12), "I3": (0.800e-6, 0.033e-6, 0.58e12),
    "I4": (0.700e-6, 0.100e-6, 2.02e12), "I5": (0.850e-6, 0.100e-6, 1.67e12), "I6": (1.000e-6, 0.100e-6, 1.42e12),
    "I7": (0.850e-6, 0.300e-6, 5.00e12), "I8": (0.750e-6, 0.100e-6, 1.89e12), "I9": (0.850e-6, 0.300e-6, 5.00e12),
    "I10": (0.900e-6, 0.300e-6, 4.72e12), "J": (1.215e-6, 0.260e-6, 1.90e12), "J2": (1.550e-6, 0.260e-6, 1.49e12),
    "H": (1.654e-6, 0.290e-6, 1.05e12), "Kp": (2.1245e-6, 0.351e-6, 0.62e12), "Ks": (2.157e-6, 0.320e-6, 0.55e12),
    "K": (2.179e-6, 0.410e-6, 0.70e12), "K0": (2.000e-6, 0.410e-6, 0.76e12), "K1": (2.400e-6, 0.410e-6, 0.64e12),
    "L": (3.547e-6, 0.570e-6, 2.5e11), "M": (4.769e-6, 0.450e-6, 8.4e10), "Na": (0.589e-6, 0.0, 3.3e12),
    "EOS": (1.064e-6, 0.0, 3.3e12), "IR1310": (1.310e-6, 0.0, 2e12),
}


class Source:
    """Natural guide star (OOPAO/Source.py:15-129).  LGS / asterisms are out of scope."""

    def __init__(self, optBand, magnitude, coordinates=[0, 0], altitude=np.inf, display_properties=False,
                 chromatic_shift=None):
        if optBand not in _PHOTOMETRY:
            raise ValueError("Error: Wrong name for the photometry object")     # Source.py:233-238 prints + returns -1
        # Off-axis sources (coordinates = [zenith arcsec, azimuth deg], Source.py:15-24) are accepted; the Atmosphere
        # propagates them when every layer is at the ground (altitude 0: the footprint does not move,
        # OOPAO/Atmosphere.py:222-230) and refuses layers in altitude (anisoplanatism is out of scope).
        self.optBand = optBand
        self.wavelength, self.bandwidth, zp = _PHOTOMETRY[optBand]
        self.zeroPoint = zp / 368
        self.magnitude = magnitude
        self.nPhoton = self.zeroPoint * 10 ** (-0.4 * magnitude)        # photons / m2 / s  (Source.py:108)
        self.coordinates = coordinates
        self.altitude = altitude
        self.chromatic_shift = chromatic_shift
        self.tag = "source"
        self.type = "NGS"
        self.fluxMap = []
        self.telescope = None
        self.is_initialized = True

    # phase views follow the coupled telescope's OPD (Telescope.py:404-412 keeps them in sync by assignment)
    @property
    def phase(self):
        return self.telescope.OPD * (2 * math.pi / self.wavelength)

    @property
    def phase_no_pupil(self):
        return self.telescope.OPD_no_pupil * (2 * math.pi / self.wavelength)

    def __mul__(self, telescope):
        """src*tel (Source.py:133-159)."""
        telescope.src = self
        self.telescope = telescope
        telescope._on_new_source()
        return telescope

"""Generates tests/golden/pyramid.npz by running the UNMODIFIED reference Pyramid WFS (OOPAO/Pyramid.py) on the CPU of
the build container — TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden_pyramid

8 m telescope at 48 px, 12 subapertures, modulation 3 lambda/D, 4 px between the pupil images, slopesMaps
post-processing (the options of MAIN_CODE/OOPAOEnv/OOPAOEnv.py:239-246).  The fixture holds the valid-pixel mask, the
reference slopes map, and camera frames + slopes for a flat wavefront, two smooth aberrations and an unmodulated run.
"""
import os

import numpy as np

from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pyramid.npz")
SETUP = dict(resolution=48, diameter=8.0, samplingTime=1 / 500, band="I", magnitude=8.0, nSubap=12, modulation=3,
             lightRatio=0.1, n_pix_separation=4, n_pix_edge=2)


def test_phases(resolution, pupil, wavelength):
    """Deterministic wavefronts (radians), shared with tests/test_pyramid_oracle.py through the fixture."""
    yy, xx = np.mgrid[:resolution, :resolution] / resolution - 0.5
    opd = [40e-9 * (3 * xx - 2 * yy) + 25e-9 * np.sin(7 * xx) * np.cos(5 * yy),
           60e-9 * (xx ** 2 - yy ** 2) + 15e-9 * np.cos(11 * xx * yy + 1.0)]
    return [o * pupil * 2 * np.pi / wavelength for o in opd]


def main():
    rh._prepare_imports()
    with rh.quiet():
        from OOPAO.Pyramid import Pyramid
        from OOPAO.Source import Source
        from OOPAO.Telescope import Telescope
        s = SETUP
        tel = Telescope(resolution=s["resolution"], diameter=s["diameter"], samplingTime=s["samplingTime"])
        src = Source(optBand=s["band"], magnitude=s["magnitude"])
        src * tel
        wfs = Pyramid(nSubap=s["nSubap"], telescope=tel, modulation=s["modulation"], lightRatio=s["lightRatio"],
                      n_pix_separation=s["n_pix_separation"], n_pix_edge=s["n_pix_edge"], postProcessing="slopesMaps",
                      binning=1)
    g = dict(pupil=tel.pupil.astype(np.uint8), fluxMap=np.asarray(src.fluxMap, dtype=np.float64),
             wavelength=np.float64(src.wavelength), nRes=np.int64(wfs.nRes), nTheta=np.int64(wfs.nTheta),
             cam_resolution=np.int64(wfs.cam.resolution), validI4Q=wfs.validI4Q.astype(np.uint8),
             nSignal=np.int64(wfs.nSignal), referenceSignal_2D=np.asarray(wfs.referenceSignal_2D, dtype=np.float64),
             mask_phase=np.asarray(wfs.m, dtype=np.float64), setup=np.array([s["resolution"], s["nSubap"], s["modulation"],
                                                                             s["n_pix_separation"], s["n_pix_edge"]], dtype=np.float64),
             lightRatio=np.float64(s["lightRatio"]))
    with rh.quiet():
        tel.resetOPD()
        tel * wfs
    g["frame_flat"] = np.asarray(wfs.cam.frame, dtype=np.float64)
    g["signal_flat"] = np.asarray(wfs.signal, dtype=np.float64)
    phases = test_phases(s["resolution"], tel.pupil, src.wavelength)
    for k, ph in enumerate(phases):
        with rh.quiet():
            wfs.wfs_measure(phase_in=ph)
        g[f"phase_{k}"] = ph
        g[f"frame_{k}"] = np.asarray(wfs.cam.frame, dtype=np.float64)
        g[f"signal_{k}"] = np.asarray(wfs.signal, dtype=np.float64)
        g[f"signal_2D_{k}"] = np.asarray(wfs.signal_2D, dtype=np.float64)
    with rh.quiet():
        wfs.modulation = 0                      # re-calibrates the reference slopes (Pyramid.py:977-984)
        wfs.wfs_measure(phase_in=phases[0])
    g["referenceSignal_2D_unmodulated"] = np.asarray(wfs.referenceSignal_2D, dtype=np.float64)
    g["signal_unmodulated_0"] = np.asarray(wfs.signal, dtype=np.float64)
    with rh.quiet():
        tel.resetOPD()
        wfs2 = Pyramid(nSubap=s["nSubap"], telescope=tel, modulation=s["modulation"], lightRatio=s["lightRatio"],
                       n_pix_separation=s["n_pix_separation"], n_pix_edge=s["n_pix_edge"],
                       postProcessing="slopesMaps_incidence_flux", binning=1)
        wfs2.wfs_measure(phase_in=phases[1])
    g["incidence_referenceSignal_2D"] = np.asarray(wfs2.referenceSignal_2D, dtype=np.float64)
    g["incidence_signal_1"] = np.asarray(wfs2.signal, dtype=np.float64)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB; nRes", wfs.nRes, "nTheta", int(g["nTheta"]), "nSignal", wfs.nSignal)


if __name__ == "__main__":
    main()

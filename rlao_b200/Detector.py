"""Detector — mirror of OOPAO/Detector.py for the WFS camera: a parameter holder whose noise chain
(photon -> QE -> dark -> full well -> read noise -> gain -> ADC, Detector.py:232-301) runs inside the WFS
camera pass (rlao_b200/csrc/wfs.cu: shwfs_detector_kernel), one Philox stream per pixel, environment and frame."""
import ctypes

import torch

from . import _lib


class Detector:
    def __init__(self, nRes=None, integrationTime=None, bits=None, output_precision=None, FWC=None, gain=1,
                 sensor="CCD", QE=1, binning=1, psf_sampling=2, darkCurrent=0, readoutNoise=0, photonNoise=False,
                 backgroundNoise=False, backgroundNoiseMap=None, seed=0):
        self.resolution = nRes
        self.integrationTime = integrationTime
        self.bits = bits
        self.output_precision = output_precision
        self.FWC = FWC
        self.gain = gain
        if sensor not in ("EMCCD", "CCD", "CMOS"):
            raise ValueError("Sensor must be 'EMCCD', 'CCD', or 'CMOS'")
        self.sensor = sensor
        self.psf_sampling = psf_sampling
        self.QE = QE
        self.binning = binning
        self.darkCurrent = darkCurrent
        self.readoutNoise = readoutNoise
        self.photonNoise = photonNoise
        self.backgroundNoise = backgroundNoise
        self.backgroundNoiseMap = backgroundNoiseMap
        self.tag = "detector"
        self._integrated_time = 0
        self.seed = int(seed)
        self.frame_counter = 0
        self._frame, self._frame_src = None, None

    @property
    def frame(self):
        """The last camera frame.  A WFS that did not write the frame of its last measurement registers a producer in
        `_frame_src`; it runs the first time the frame is read."""
        if self._frame_src is not None:
            src, self._frame_src = self._frame_src, None
            self._frame = src()
        return self._frame

    @frame.setter
    def frame(self, val):
        self._frame, self._frame_src = val, None

    def is_ideal(self):
        """True when the chain is the identity (the default WFS camera, ShackHartmann.py:166-168)."""
        return (not self.photonNoise and not self.readoutNoise and not self.darkCurrent and self.QE == 1
                and self.FWC is None and self.bits is None and self.gain == 1)

    def as_struct(self, env_offset=0):
        """aoenv_detector_t for the next frame, or None for the ideal detector."""
        if self.is_ideal():
            return None
        if self.backgroundNoise:
            raise NotImplementedError("background noise maps are out of scope")
        if self.binning != 1:
            raise NotImplementedError("detector binning is out of scope")
        if self.bits is not None and self.FWC is None:
            raise NotImplementedError("ADC without FWC needs the frame maximum (Detector.py:192-193); not supported")
        d = _lib.DetectorStruct()
        d.photon_noise = int(bool(self.photonNoise))
        d.sensor_emccd = int(self.sensor == "EMCCD")
        d.has_fwc = int(self.FWC is not None)
        d.bits = int(self.bits or 0)
        d.qe = float(self.QE)
        d.dark_electrons = float(self.darkCurrent * (self.integrationTime or 0))
        d.fwc = float(self.FWC or 0)
        d.gain = float(self.gain)
        d.readout_noise = float(self.readoutNoise or 0)
        d.seed = (self.seed * 2654435761 + env_offset * 97 + 1) & 0xFFFFFFFFFFFFFFFF
        d.frame_counter = self.frame_counter
        self.frame_counter += 1
        return d

    def integrate(self, frame, env_offset=0):
        """Detector.py:279-301: applies the camera to a frame of photons ([rows, cols] or [B, rows, cols], CUDA float32) and
        stores / returns `self.frame`.  Every call advances the frame counter of the random streams."""
        f = torch.as_tensor(frame, dtype=torch.float32)
        if f.device.type != "cuda":
            raise _lib.AOEnvLibraryError("Detector.integrate needs a CUDA tensor (there is no CPU path)")
        out = (f.unsqueeze(0) if f.ndim == 2 else f).contiguous().clone()
        det = self.as_struct(env_offset)
        if det is not None:
            _lib.check(_lib.load().aoenv_detector_integrate(_lib.ptr(out), out.shape[0], out.shape[1], out.shape[2],
                                                            ctypes.byref(det), _lib.stream_ptr(out.device)), "detector_integrate")
        self.frame = out[0] if f.ndim == 2 else out
        return self.frame


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
        self._buffer_sum, self.n_buffered, self._env_offset, self._squeeze_out = None, 0, 0, False
        self.fov_arcsec = self.fov_rad = self.pixel_size_rad = self.pixel_size_arcsec = None
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
        if self.binning != 1 and not getattr(self, "_staged", False):
            raise NotImplementedError("detector binning inside the WFS camera pass is out of scope")
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

    def _chain(self, frame, stages, env_offset):
        """The camera kernel on `frame` [B, rows, cols] in place, restricted to `stages` (aoenv_detector_t.reserved)."""
        self._staged = True
        try:
            det = self.as_struct(env_offset)
        finally:
            self._staged = False
        if det is None:
            return
        det.reserved = stages
        _lib.check(_lib.load().aoenv_detector_integrate(_lib.ptr(frame), frame.shape[0], frame.shape[1], frame.shape[2],
                                                        ctypes.byref(det), _lib.stream_ptr(frame.device)), "detector_integrate")

    def integrate(self, frame, env_offset=0):
        """Detector.py:279-301: one sub-frame of photons ([rows, cols] or [B, rows, cols], CUDA float32) — photon noise, QE —
        is added to the exposure buffer; when the integrated time reaches `integrationTime` (the caller advances
        `_integrated_time`, Telescope.py:495) the buffer is read out (`readout`, Detector.py:232-276) into `self.frame`.
        Every kernel call advances the frame counter of the random streams."""
        f = torch.as_tensor(frame, dtype=torch.float32)
        _lib.require_cuda(f.device)                                       # raises: there is no CPU path
        self._squeeze_out = f.ndim == 2
        sub = (f.unsqueeze(0) if f.ndim == 2 else f).contiguous().clone()
        self.perfect_frame = sub[0] if self._squeeze_out else sub
        self.flux_max_px = sub.amax(dim=(-2, -1))
        self.signal = self.QE * self.flux_max_px
        self._chain(sub, 1, env_offset)                                   # photon noise, QE
        self._buffer_sum = sub if self._buffer_sum is None else self._buffer_sum + sub
        self.n_buffered += 1
        self._env_offset = env_offset
        if self.integrationTime is None or self._integrated_time >= self.integrationTime:
            self.readout()
        return self.frame

    def readout(self):
        """Detector.py:232-276: sum of the buffered sub-frames -> dark current, full well, [EM gain] -> hardware binning ->
        read noise, gain, ADC; resets the buffer and the integrated time."""
        out = self._buffer_sum
        self._chain(out, 2, self._env_offset)                             # dark shot noise, saturation, EM gain
        if self.binning != 1:                                             # tools.py:409-416 (sum)
            b = int(self.binning)
            B, r, c = out.shape
            if r % b != 0:
                raise ValueError("the frame size must be a multiple of the detector binning")
            out = out.reshape(B, r // b, b, c // b, b).sum(dim=(2, 4)).contiguous()
        self._chain(out, 4, self._env_offset)                             # read-out noise, gain, quantisation
        self.frame = out[0] if self._squeeze_out else out
        if self.resolution is None:
            self.resolution = out.shape[-1]
        if self.fov_arcsec is not None:
            self.pixel_size_rad = self.fov_rad / self.resolution
            self.pixel_size_arcsec = self.fov_arcsec / self.resolution
        self.n_frames_last_exposure = self.n_buffered
        self._buffer_sum, self.n_buffered, self._integrated_time = None, 0, 0


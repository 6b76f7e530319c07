"""Spatial partition of one B200 between the two pipelines of the environment step.

The step is two independent chains: the wavefront sensor of frame t (bound by the FP32 pipe: `shwfs_frame` fills the
register file of every SM it runs on) and the atmosphere of frame t+1 (bound by HBM: sub-pixel shift, ring writes,
gathers).  Issued on two ordinary streams they barely overlap — the block scheduler hands an SM to the second kernel only
when the first has no CTA left to place — so the GPU alternates between an FP32-bound and an HBM-bound phase and each
leaves the other resource idle.  CUDA green contexts split the SMs instead: the atmosphere chain gets `side_sms` SMs (an
HBM-bound kernel saturates the memory system from a fraction of them), the sensor chain the rest, and both run all the
time.  Streams of the two green contexts are handed to PyTorch as external streams; memory, events and modules are those
of the primary context, so nothing else changes.

Plumbing only (driver API through cuda-python, which ships in this image); every kernel is still launched by
libaoenv_b200.so on the stream it is given.
"""
import torch

_partitions = {}


class SMPartition:
    def __init__(self, device, side_sms):
        from cuda.bindings import driver as drv
        self._drv = drv
        dev_index = torch.device(device).index or 0
        torch.cuda.init()
        with torch.cuda.device(dev_index):
            torch.zeros(1, device=f"cuda:{dev_index}")          # the primary context exists and is current
            (err,) = drv.cuInit(0)
            self._check(err, "cuInit")
            err, dev = drv.cuDeviceGet(dev_index)
            self._check(err, "cuDeviceGet")
            err, res = drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM)
            self._check(err, "cuDeviceGetDevResource")
            self.total_sms = int(res.sm.smCount)
            err, groups, n, remaining = drv.cuDevSmResourceSplitByCount(1, res, 0, int(side_sms))
            self._check(err, "cuDevSmResourceSplitByCount")
            if n < 1:
                raise RuntimeError(f"cannot split {side_sms} SMs off a device with {self.total_sms}")
            self.side_sms, self.main_sms = int(groups[0].sm.smCount), int(remaining.sm.smCount)
            self._ctx, self._streams = [], []
            for r in (remaining, groups[0]):
                err, desc = drv.cuDevResourceGenerateDesc([r], 1)
                self._check(err, "cuDevResourceGenerateDesc")
                err, g = drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM)
                self._check(err, "cuGreenCtxCreate")
                err, s = drv.cuGreenCtxStreamCreate(g, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0)
                self._check(err, "cuGreenCtxStreamCreate")
                self._ctx.append(g)
                self._streams.append(s)
            self.main = torch.cuda.ExternalStream(int(self._streams[0]), device=f"cuda:{dev_index}")
            self.side = torch.cuda.ExternalStream(int(self._streams[1]), device=f"cuda:{dev_index}")

    def _check(self, err, what):
        if int(err) != 0:
            raise RuntimeError(f"{what} failed: {self._drv.cuGetErrorName(err)[1]}")

    def describe(self):
        return f"{self.main_sms} SMs sensor / control chain + {self.side_sms} SMs atmosphere chain (green contexts)"


def get(device, side_sms):
    """The partition of `device` with `side_sms` SMs for the atmosphere chain (created once per process and device)."""
    key = (torch.device(device).index or 0, int(side_sms))
    if key not in _partitions:
        _partitions[key] = SMPartition(device, side_sms)
    return _partitions[key]

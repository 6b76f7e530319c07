"""rlao_b200 — the closed-loop adaptive-optics environment step of drl4ao/OOPAO, batched over environments and
run by hand-written sm_100a CUDA kernels (libaoenv_b200.so, C ABI in include/aoenv.h).

The modules keep the reference's names so existing scripts switch by changing one import root:

    reference                                  this package
    OOPAO.Telescope.Telescope                  rlao_b200.Telescope.Telescope
    OOPAO.Source.Source                        rlao_b200.Source.Source
    OOPAO.Atmosphere.Atmosphere                rlao_b200.Atmosphere.Atmosphere
    OOPAO.DeformableMirror.DeformableMirror    rlao_b200.DeformableMirror.DeformableMirror
    OOPAO.ShackHartmann.ShackHartmann          rlao_b200.ShackHartmann.ShackHartmann
    OOPAO.Detector.Detector                    rlao_b200.Detector.Detector
    OOPAO.Zernike.Zernike                      rlao_b200.Zernike.Zernike
    OOPAO.calibration.InteractionMatrix        rlao_b200.calibration.InteractionMatrix
    OOPAO.calibration.CalibrationVault         rlao_b200.calibration.CalibrationVault
    OOPAOEnv.OOPAOEnvRazor.OOPAO               rlao_b200.OOPAOEnv.OOPAOEnvRazor.OOPAO
    PO4AO.util_simple.{TorchWrapper,...}       rlao_b200.PO4AO.util_simple

Every array gains a leading `n_envs` dimension and lives on the GPU (torch tensors); with n_envs == 1 the
public attributes are squeezed to the reference's shapes.  There is no CPU fallback.
"""
__version__ = "0.1.0"

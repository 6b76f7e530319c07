"""Pyramid environment — mirror of MAIN_CODE/OOPAOEnv/OOPAOEnv.py `class OOPAO` (the environment drl4ao's test_int.sh /
test_po4ao.sh scripts build): the same optical train and step as the Shack-Hartmann environment with a Pyramid WFS, and
a step that also returns the WFS camera frame (OOPAOEnv.py:485-536: `obs, wfsf, reward, strehl, done, info`)."""
from .OOPAOEnvRazor import OOPAO as _RazorOOPAO


class OOPAO(_RazorOOPAO):
    returns_frame = True          # step() is a 6-tuple: wrappers must not take the step apart and drop the frame

    def set_params(self, args=None, wfs_type="pyramid", modal_basis="zernike", gainCL=0.5, **kw):
        """OOPAOEnv.py:93-404."""
        return super().set_params(args, wfs_type, modal_basis, gainCL, **kw)

    def step(self, i, action):
        """OOPAOEnv.py:485-536."""
        obs, reward, strehl, done, info = super().step(i, action)
        frame = self.wfs.cam.frame
        return obs, frame.clone(), reward, strehl, done, info

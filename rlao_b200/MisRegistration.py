"""MisRegistration — the six-field struct of OOPAO/MisRegistration.py:14-66 consumed by DeformableMirror."""

_FIELDS = ("rotationAngle", "shiftX", "shiftY", "anamorphosisAngle", "tangentialScaling", "radialScaling")


class MisRegistration:
    def __init__(self, param=None):
        self.tag = "misRegistration"
        for f in _FIELDS:
            if param is None:
                v = 0
            elif isinstance(param, dict):
                v = param[f]
            elif getattr(param, "tag", None) == "misRegistration":
                v = getattr(param, f)
            else:
                raise TypeError("wrong type of object passed to a MisRegistration object")
            setattr(self, f, v)
        self.isInitialized = True

    @property
    def misRegName(self):
        p = "%.2f" if (self.radialScaling == 0 and self.tangentialScaling == 0) else "%.4f"
        return ("rot_%.2f_sX_%.2f_m_sY_%.2f_m_anam_%.2f_" % (self.rotationAngle, self.shiftX, self.shiftY, self.anamorphosisAngle)
                + "mR_" + p % (self.radialScaling + 1.0) + "_mT_" + p % (self.tangentialScaling + 1.0))

"""Display stub: the reference imports matplotlib at module top (Atmosphere.py:13,16,
Telescope.py:10, tools/displayTools.py:8-12, CalibrationVault.py:10) but the closed-loop
path never needs a figure."""
import sys
from unittest.mock import MagicMock

for _sub in ("pyplot", "gridspec", "offsetbox", "animation", "colors", "cm", "patches"):
    _m = MagicMock()
    sys.modules["matplotlib." + _sub] = _m
    globals()[_sub] = _m

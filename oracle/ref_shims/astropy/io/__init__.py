"""Stub for `from astropy.io import fits` (OOPAO/tools/tools.py:15); FITS I/O is unused."""
from unittest.mock import MagicMock
fits = MagicMock()

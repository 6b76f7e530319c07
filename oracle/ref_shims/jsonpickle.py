"""Stub: only reached when Atmosphere(param=...) is given (Atmosphere.py:510-545)."""
def encode(x):
    raise NotImplementedError("jsonpickle stub")
def decode(x):
    raise NotImplementedError("jsonpickle stub")

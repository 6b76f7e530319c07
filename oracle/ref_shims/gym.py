"""Minimal stand-in for the two gym names the drl4ao env layer uses
(MAIN_CODE/OOPAOEnv/OOPAOEnvRazor.py:9,11; MAIN_CODE/PO4AO/util_simple.py:4,25,201)."""
class Env:
    metadata = {}
class Wrapper:
    def __init__(self, env):
        self.env = env
    def __getattr__(self, name):
        return getattr(self.env, name)

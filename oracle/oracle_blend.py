"""The CPU oracle built with the reference's feature="blend" probability model (BlendCDF16, probability/blend_cdf.rs): the module
oracle_py executed a second time against oracle/_build/libdivans_oracle_blend.so; every attribute of that second instance is
reachable here (oracle_blend.encode_raw, .decode, .Commands, .lib(), ...).  TEST INFRASTRUCTURE, like oracle_py."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location("oracle_py_blend", os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_py.py"))
_mod = importlib.util.module_from_spec(_spec)
_mod._VARIANT = "blend"
_spec.loader.exec_module(_mod)


def __getattr__(name):
    return getattr(_mod, name)

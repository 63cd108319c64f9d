"""Import shim: the package directory `mri-super-resolution_b200` is not a valid Python identifier."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("mri-super-resolution_b200")
inr = importlib.import_module("mri-super-resolution_b200.inr")
SRDWI = importlib.import_module("mri-super-resolution_b200.SRDWI")
INRmodel = importlib.import_module("mri-super-resolution_b200.INRmodel")
nn_mri = importlib.import_module("mri-super-resolution_b200.nn_mri")
phantom = importlib.import_module("mri-super-resolution_b200.phantom")
parallel = importlib.import_module("mri-super-resolution_b200.parallel")
_lib = importlib.import_module("mri-super-resolution_b200._lib")
from_pkg = _pkg.__all__
globals().update({k: getattr(_pkg, k) for k in from_pkg})
__all__ = list(from_pkg) + ["inr", "SRDWI", "INRmodel", "nn_mri", "phantom", "parallel"]

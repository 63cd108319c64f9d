"""`nn_mri` as a TOP-LEVEL module name (see SRDWI.py in this directory): INR/inr_toy.py:3, INR/INR_ERD.py:1 and
INR/automate_INR.py:9 import their model classes from it."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
import b200inr as _b200inr  # noqa: E402  (a module, not a package: the package directory name has hyphens)

_impl = _b200inr.nn_mri
globals().update({_k: getattr(_impl, _k) for _k in _impl.__all__})
__all__ = list(_impl.__all__)

"""`INRmodel` as a TOP-LEVEL module name (see SRDWI.py in this directory): INR/inrDWI.py:9 imports it."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
import b200inr as _b200inr  # noqa: E402  (a module, not a package: the package directory name has hyphens)

_impl = _b200inr.INRmodel
globals().update({_k: getattr(_impl, _k) for _k in _impl.__all__})
__all__ = list(_impl.__all__)

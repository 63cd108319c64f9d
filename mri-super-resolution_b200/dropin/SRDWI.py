"""`SRDWI` as a TOP-LEVEL module name: put this directory first on sys.path / PYTHONPATH and the reference's own import
lines (INR/superresDWI.py:13, INR/superresHybrid.py:13, INR/forbagci.py:10, INR/automate_INR.py:10) resolve to the B200
implementation without editing the scripts."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
import b200inr as _b200inr  # noqa: E402  (a module, not a package: the package directory name has hyphens)

_impl = _b200inr.SRDWI
globals().update({_k: getattr(_impl, _k) for _k in _impl.__all__})
__all__ = list(_impl.__all__)

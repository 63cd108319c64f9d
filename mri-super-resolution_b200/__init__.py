"""B200-native implementation of the INR fit / query hot path of MRIRC/MRI-super-resolution.

The directory name contains hyphens, so import it with importlib (or through the root-level ``b200inr`` shim):

    import b200inr                      # repo root on sys.path
    from b200inr import Siren, get_mgrid, input_mapping

Modules: ``inr`` (reference-facing nn.Module surface + fused fit/query), ``SRDWI`` / ``INRmodel`` (drop-in modules
with exactly the reference's import names), ``perturb`` (the fused PerturbNet loop), ``phantom`` (synthetic DWI volumes), ``_lib`` (ctypes binding of the C
ABI declared in include/b200inr.h), ``csrc`` (the sm_100a kernels).
"""
from .inr import (ComplexGaborLayer2D, FitSession, FourierMLP, Wire, ImageFitting_set, PN, SineLayer, Siren,  # noqa: F401
                  SirenERD, all_combinations, calculate_ADC, calculate_combinations, get_mgrid, input_mapping,
                  resize_array, soft_erd)
from .perturb import PerturbSession, perturb_fit  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["ComplexGaborLayer2D", "FitSession", "FourierMLP", "Wire", "ImageFitting_set", "PN", "SineLayer", "Siren",
           "SirenERD", "all_combinations", "calculate_ADC", "calculate_combinations", "get_mgrid", "input_mapping",
           "resize_array", "soft_erd", "PerturbSession", "perturb_fit"]

"""Drop-in for the reference's INR/SRDWI.py: every name its importers ask for (INR/superresDWI.py:13,
INR/superresHybrid.py:13, INR/forbagci.py:10, INR/automate_INR.py:10) with the reference's signatures."""
from .inr import (PN, ImageFitting_set, SineLayer, Siren, calculate_ADC, calculate_combinations, get_mgrid,  # noqa: F401
                  input_mapping, resize_array)

__all__ = ["ImageFitting_set", "PN", "SineLayer", "Siren", "calculate_ADC", "calculate_combinations", "get_mgrid",
           "input_mapping", "resize_array"]

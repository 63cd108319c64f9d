"""Drop-in for the reference's INR/SRDWI.py: `from SRDWI import *` keeps working (hot-path symbols only;
calculate_ADC / resize_array / calculate_combinations are CPU post-processing outside this path)."""
from .inr import ImageFitting_set, SineLayer, Siren, calculate_ADC, calculate_combinations, get_mgrid, input_mapping  # noqa: F401

__all__ = ["ImageFitting_set", "SineLayer", "Siren", "calculate_ADC", "calculate_combinations", "get_mgrid",
           "input_mapping"]

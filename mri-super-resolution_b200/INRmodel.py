"""Drop-in for the reference's INR/INRmodel.py (Siren :122-151: sine layers constructed before the final linear, no
first_omega_0 argument, coordinates not detached by the module)."""
from .inr import PN, ImageFitting_set, SineLayer, calculate_ADC, calculate_combinations, get_mgrid, input_mapping  # noqa: F401
from .inr import Siren as _Siren


class Siren(_Siren):
    def __init__(self, in_features, hidden_features, hidden_layers, out_features, hidden_omega_0=30.):
        super().__init__(in_features, hidden_features, hidden_layers, out_features, hidden_omega_0=hidden_omega_0,
                         variant="INRmodel")


__all__ = ["ImageFitting_set", "PN", "SineLayer", "Siren", "calculate_ADC", "calculate_combinations", "get_mgrid",
           "input_mapping"]

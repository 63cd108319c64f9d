"""Drop-in for the reference's INR/INRmodel.py (imported at INR/inrDWI.py:9): Siren :122-151 builds its sine layers
before the final linear, has no first_omega_0 argument and does not detach its input; ComplexGaborLayer2D :66-120."""
from .inr import (PN, ComplexGaborLayer2D, ImageFitting_set, SineLayer, calculate_ADC, calculate_combinations,  # noqa: F401
                  get_mgrid, input_mapping, resize_array)
from .inr import Siren as _Siren


class Siren(_Siren):
    def __init__(self, in_features, hidden_features, hidden_layers, out_features, hidden_omega_0=30.):
        super().__init__(in_features, hidden_features, hidden_layers, out_features, hidden_omega_0=hidden_omega_0,
                         variant="INRmodel")


__all__ = ["ComplexGaborLayer2D", "ImageFitting_set", "PN", "SineLayer", "Siren", "calculate_ADC",
           "calculate_combinations", "get_mgrid", "input_mapping", "resize_array"]

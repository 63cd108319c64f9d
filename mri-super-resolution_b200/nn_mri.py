"""Drop-in for the model half of the reference's INR/nn_mri.py (imported by INR/inr_toy.py:3, INR/INR_ERD.py:1,
INR/automate_INR.py:9): get_mgrid(sidelen, dim) :87-94, SineLayer :96-120, Siren :122-146, PN :148-163 (two-dimensional
perturbation), input_mapping :165-171, ImageFitting_set :181-205.  The plotting / metric helpers of that file are outside
the hot path."""
import numpy as np
import torch
from torch.utils.data import Dataset

from .inr import PN as _PN
from .inr import SineLayer, Siren, input_mapping  # noqa: F401
from .inr import get_mgrid as _get_mgrid


def get_mgrid(sidelen, dim=2):
    """Reference INR/nn_mri.py:87-94: a flattened grid of `dim` axes with `sidelen` points each in [-1, 1]."""
    return _get_mgrid((int(sidelen),) * int(dim))


class PN(_PN):
    """Reference INR/nn_mri.py:148-163: the perturbation network with a fixed two-dimensional output."""

    def __init__(self, in_features, hidden_features):
        super().__init__(in_features, hidden_features, 2)


class ImageFitting_set(Dataset):
    """Reference INR/nn_mri.py:181-205 ("rearranged for MR dataset"): a list of square PIL images -> pixels
    [n, side^2, 1] normalised to (x - 0.5) / 0.5 after ToTensor, coords [n, side^2, 2], plus the raw images (`orig`), their
    mean and the shape.  Host-side data preparation with the same torchvision transforms as the reference."""

    def __init__(self, img_dataset):
        super().__init__()
        from torchvision.transforms import Compose, Normalize, Resize, ToTensor
        side = img_dataset[0].size[0]
        tf = Compose([Resize(side), ToTensor(), Normalize(torch.Tensor([0.5]), torch.Tensor([0.5]))])
        n = len(img_dataset)
        self.orig = np.empty((n, img_dataset[0].size[0], img_dataset[0].size[1]))
        self.pixels = torch.empty((n, side ** 2, 1))
        self.coords = torch.empty((n, side ** 2, 2))
        grid = get_mgrid(side, 2)
        for i, img in enumerate(img_dataset):
            self.orig[i] = np.array(img)
            self.pixels[i] = tf(img).permute(1, 2, 0).reshape(-1, 1)
            self.coords[i] = grid
        self.mean = sum(self.orig) / len(self.orig)
        self.shape = img_dataset[0].size

    def __len__(self):
        return len(self.pixels)

    def __getitem__(self, idx):
        return self.coords, self.pixels


__all__ = ["ImageFitting_set", "PN", "SineLayer", "Siren", "get_mgrid", "input_mapping"]

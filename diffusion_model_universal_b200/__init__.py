"""B200-native denoiser hot path of ChristianLin0420/diffusion-model-universal.

Drop-in for the reference's model classes: same ``MODEL_REGISTRY`` keys
(scripts/train.py:41-46), same ``model_class(config['model_config'])``
construction, same ``state_dict`` layout; the arithmetic runs in
``lib/libdmu_b200.so`` (hand-written sm_100a CUDA behind include/dmu_b200.h).
"""

from .base_model import BaseDiffusion
from .ddpm import DDPM
from .ddim import DDIM
from .score_based import ScoreBasedDiffusion
from .energy_based import EnergyBasedDiffusion
from .unet import UNet
from .losses import DiffusionLoss

MODEL_REGISTRY = {
    "ddpm": DDPM,
    "ddim": DDIM,
    "score_based": ScoreBasedDiffusion,
    "energy_based": EnergyBasedDiffusion,
}

__all__ = ["BaseDiffusion", "DDPM", "DDIM", "ScoreBasedDiffusion", "EnergyBasedDiffusion", "UNet", "DiffusionLoss", "MODEL_REGISTRY"]

"""Energy-based model behind the reference API (models/energy_based.py)."""

import math
from typing import Dict, Optional

import torch

from . import ops
from .base_model import BaseDiffusion


class EnergyBasedDiffusion(BaseDiffusion):
    """Placeholder until the EnergyNet engine lands (see energy_net.py)."""

    def __init__(self, config: Dict):
        super().__init__(config)
        raise NotImplementedError("EnergyBasedDiffusion: CUDA engine for EnergyNet not built yet")

    def forward(self, x, t=None):
        raise NotImplementedError

    def loss_function(self, x):
        raise NotImplementedError

    def generate_samples(self, batch_size, device):
        raise NotImplementedError

"""BaseDiffusion — the class API the hot path sits behind (models/base_model.py:19-149).

Kept method-for-method: abstract ``forward`` / ``loss_function`` /
``generate_samples``, concrete ``save`` / ``load`` writing
``{'model_state_dict', 'config'}``.  ``sample`` is exported as an alias of
``generate_samples`` on every subclass because the README / north-star use
that name while the trainers call ``generate_samples``.
"""

from abc import ABC, abstractmethod
from typing import Dict

import torch
import torch.nn as nn


class BaseDiffusion(nn.Module, ABC):
    def __init__(self, config: Dict):
        super().__init__()
        self.config = config

    @abstractmethod
    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @abstractmethod
    def loss_function(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @abstractmethod
    def generate_samples(self, batch_size: int, device: torch.device) -> torch.Tensor:
        raise NotImplementedError

    def sample(self, batch_size: int, device: torch.device) -> torch.Tensor:
        return self.generate_samples(batch_size, device)

    def save(self, path: str) -> None:
        """models/base_model.py:119-133."""
        torch.save({"model_state_dict": self.state_dict(), "config": self.config}, path)

    def load(self, path: str) -> None:
        """models/base_model.py:135-149."""
        checkpoint = torch.load(path)
        self.load_state_dict(checkpoint["model_state_dict"])
        self.config = checkpoint["config"]

"""Device-side input ingest and sample formatting around the denoiser path.

Mirrors what sits either side of the hot loop in the reference:
  * datasets/dataset_utils.py:58-61 — ``T.ToTensor()`` + ``T.Normalize(mean, std)`` in the DataLoader workers, then
    ``batch[0].to(device)`` (trainers/ddpm_trainer.py:539): here the loader hands over the decoded uint8 pixels, the copy
    carries one byte per value and the conversion is one launch (``DeviceIngest``; ``TrainStep(input_norm=...)`` fuses
    it into q_sample).
  * trainers/ddpm_trainer.py:821-834 — ``make_grid`` + ``save_image`` of the denoising-process rows: here one launch
    writes the 8-bit HWC grid on the device and only those bytes cross to the host (``image_grid``, ``denoising_grid``).
Everything runs in libdmu_b200.so; CPU tensors are staged to the device, never processed on the host.
"""

from typing import List, Optional, Sequence

import torch

from . import ops


class DeviceIngest:
    """uint8 batch ([B,H,W,C] "NHWC" as decoded, or [B,C,H,W]) -> normalised fp32 [B,C,H,W] on ``device``."""

    def __init__(self, mean: Optional[Sequence[float]], std: Optional[Sequence[float]], device, layout: str = "NHWC"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceIngest runs on CUDA (sm_100a) only; there is no CPU fallback.")
        if layout not in ("NCHW", "NHWC"):
            raise ValueError(f"layout must be 'NCHW' or 'NHWC', got {layout!r}")
        self.layout = layout
        self.mean = None if mean is None else torch.tensor(list(mean), dtype=torch.float32, device=self.device)
        self.std = None if std is None else torch.tensor(list(std), dtype=torch.float32, device=self.device)
        self._stage = None

    def __call__(self, batch_u8: torch.Tensor) -> torch.Tensor:
        if batch_u8.dtype != torch.uint8:
            raise TypeError(f"expected the decoded uint8 pixels, got {batch_u8.dtype}")
        if not batch_u8.is_cuda:
            if self._stage is None or self._stage.shape != batch_u8.shape:
                self._stage = torch.empty(batch_u8.shape, device=self.device, dtype=torch.uint8)
            self._stage.copy_(batch_u8, non_blocking=True)
            batch_u8 = self._stage
        x0, _ = ops.ingest_u8(batch_u8.contiguous(), self.mean, self.std, self.layout)
        return x0


def image_grid(samples: torch.Tensor, nrow: int = 8, padding: int = 2, pad_value: float = 0.0, value_range=None) -> torch.Tensor:
    """``save_image(samples, nrow=, padding=, pad_value=[, normalize=True, value_range=])``'s pixel array: uint8
    [Hg, Wg, 3] on the device (scripts/generate.py:119-133 passes value_range=(-1, 1))."""
    return ops.image_grid_u8(samples.contiguous(), nrow, padding, pad_value, value_range=value_range)


def denoising_grid(intermediates: List[torch.Tensor], padding: int = 2, path: Optional[str] = None) -> torch.Tensor:
    """trainers/ddpm_trainer.py:815-834: one row per sample, one column per saved denoising step (``nrow`` = number of
    saved steps, 11 at the trainer's save_interval).  Returns the uint8 [Hg, Wg, 3] grid on the host; writes ``path``
    with PIL when given."""
    stacked = torch.stack(list(intermediates), dim=0)                # [steps, B, C, H, W]; the trainer's cat/cat is a view of it
    grid = ops.image_grid_u8(stacked, nrow=stacked.shape[0], padding=padding, transpose=True).cpu()
    if path is not None:
        from PIL import Image
        Image.fromarray(grid.numpy()).save(path)
    return grid

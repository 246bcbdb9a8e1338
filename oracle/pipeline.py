"""Input transform and sample-grid formatting of the reference, restated.  TEST INFRASTRUCTURE.

The arithmetic lives in torchvision (requirements.txt: `torchvision`, unpinned; this image has 0.26), called from
datasets/dataset_utils.py:58-61 (`T.ToTensor`, `T.Normalize`) and trainers/ddpm_trainer.py:821-834 (`make_grid`,
`save_image`).  Restated here with plain torch ops on CPU; tests/test_oracle_golden.py pins every function against the
installed torchvision on the same inputs (bit-exact).
"""

import math

import torch


def to_tensor(img_hwc_u8: torch.Tensor) -> torch.Tensor:
    """torchvision.transforms.functional.to_tensor for a uint8 HWC image: CHW float32 in [0, 1]."""
    return img_hwc_u8.permute(2, 0, 1).contiguous().to(torch.float32).div(255)


def normalize(x_chw: torch.Tensor, mean, std) -> torch.Tensor:
    """torchvision.transforms.functional.normalize: (x - mean[c]) / std[c]."""
    mean = torch.as_tensor(mean, dtype=torch.float32)[:, None, None]
    std = torch.as_tensor(std, dtype=torch.float32)[:, None, None]
    return x_chw.clone().sub_(mean).div_(std)


def ingest(batch_u8: torch.Tensor, mean=None, std=None, layout="NHWC") -> torch.Tensor:
    """What the DataLoader hands the trainer (datasets/dataset_utils.py:58-61 + default collate): fp32 [B,C,H,W]."""
    out = []
    for img in batch_u8:
        x = to_tensor(img if layout == "NHWC" else img.permute(1, 2, 0))
        if mean is not None and std is not None:
            x = normalize(x, mean, std)
        elif mean is not None:
            x = x - torch.as_tensor(mean, dtype=torch.float32)[:, None, None]
        elif std is not None:
            x = x / torch.as_tensor(std, dtype=torch.float32)[:, None, None]
        out.append(x)
    return torch.stack(out) if out else torch.empty((0,) + tuple(batch_u8.shape[1:]), dtype=torch.float32)


def make_grid(x: torch.Tensor, nrow: int = 8, padding: int = 2, pad_value: float = 0.0, value_range=None) -> torch.Tensor:
    """torchvision.utils.make_grid for a [N,C,H,W] batch -> [C', Hg, Wg]; value_range=(lo, hi) is normalize=True with
    that range (norm_ip: clamp, subtract lo, divide by max(hi - lo, 1e-5)), as scripts/generate.py:119-133 calls it."""
    if x.size(1) == 1:
        x = torch.cat((x, x, x), 1)
    if value_range is not None:
        lo, hi = value_range
        x = x.clone().clamp_(min=lo, max=hi).sub_(lo).div_(max(hi - lo, 1e-5))
    if x.size(0) == 1:
        return x.squeeze(0)
    n = x.size(0)
    xmaps = min(nrow, n)
    ymaps = int(math.ceil(float(n) / xmaps))
    height, width = int(x.size(2) + padding), int(x.size(3) + padding)
    grid = x.new_full((x.size(1), height * ymaps + padding, width * xmaps + padding), pad_value)
    k = 0
    for yy in range(ymaps):
        for xx in range(xmaps):
            if k >= n:
                break
            grid[:, yy * height + padding:(yy + 1) * height, xx * width + padding:(xx + 1) * width] = x[k]
            k += 1
    return grid


def to_u8_hwc(grid_chw: torch.Tensor) -> torch.Tensor:
    """torchvision.utils.save_image's array: mul(255).add_(0.5).clamp_(0, 255) -> HWC uint8."""
    return grid_chw.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)


def denoising_rows(intermediates):
    """trainers/ddpm_trainer.py:821-829: one row per sample, one column per saved step -> [B*steps, C, H, W]."""
    b = intermediates[0].shape[0]
    rows = []
    for i in range(b):
        rows.append(torch.cat([s[i:i + 1] for s in intermediates], dim=0))
    return torch.cat(rows, dim=0)

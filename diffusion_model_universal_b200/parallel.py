"""Data-parallel gradient exchange (SURVEY.md §8 a13 / e).

The reference wraps the model in DDP but never triggers the reducer
(trainers/ddpm_trainer.py:130-136 vs :543-547), so its ranks silently diverge.
Here the flat gradient arena is averaged across ranks in buckets, each an
asynchronous ``all_reduce`` (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU
tests).  The arena is laid out [time MLPs | attention qkv | everything else in
forward order]; the backward finishes gradients in roughly reverse forward
order, so buckets are issued from the arena's tail towards its head and the
first bucket (time MLP gradients, completed last) goes out at the very end.
"""

import os
from typing import List, Tuple

import torch
import torch.distributed as dist

# The all-reduces run BESIDE the backward's two lanes, whose kernels are sized to share every SM: with NCCL's default of 32 CTAs per
# collective the ring kernels take SM slots from them.  Measured on B200s (bench.py, 128 img / GPU): 16 CTAs carry the 16 MB buckets
# just as well - 2 GPUs 92.8k -> 94.2k img/s (0.950 -> 0.964 of linear), 8 GPUs 362.0k -> 371.5k img/s (0.926 -> 0.951); 8 and 4
# CTAs lose (91.4k / 90.8k at 2 GPUs: the last, exposed all-reduce gets slower).  Read by NCCL when the communicator is created, so
# this module must be imported before torch.distributed.init_process_group(..., device_id=...); a value set by the user wins.
os.environ.setdefault("NCCL_MAX_CTAS", "16")


def bucket_ranges(total: int, bucket_elems: int) -> List[Tuple[int, int]]:
    """[lo, hi) element ranges covering [0, total), issued tail-first."""
    if total <= 0:
        return []
    bucket_elems = max(1, int(bucket_elems))
    out, hi = [], total
    while hi > 0:
        lo = max(0, hi - bucket_elems)
        out.append((lo, hi))
        hi = lo
    return out


class GradAllReducer:
    """Averages ``engine.gflat`` across the process group."""

    def __init__(self, unet, bucket_mb: float = 16.0, group=None):
        self.unet, self.group = unet, group
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def launch(self, lo: int = 0, hi: int = None):
        """Issue the buckets of gflat[lo:hi] asynchronously (tail first); returns the work handles."""
        if self.world == 1:
            return []
        g = self.unet.engine.gflat
        hi = g.numel() if hi is None else hi
        works = []
        for a, b in bucket_ranges(hi - lo, self.bucket_elems):
            works.append(dist.all_reduce(g[lo + a:lo + b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return works

    def finish(self, works) -> float:
        """Wait for the buckets; returns the scale (1/world) the optimizer must apply to the summed gradients."""
        for w in works:
            w.wait()
        return 1.0 / self.world

    def allreduce(self) -> float:
        return self.finish(self.launch())

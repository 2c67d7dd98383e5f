"""Output side of the path: progressive accumulation in batches, checkpoint / resume, and a per-pixel
noise estimate so that a relRMSE against a reference image can be read next to its own noise floor.

The reference renders a scene in one go and keeps nothing (camera.rs:79-126).  Here the sample indices of every pixel are
split into two interleaved halves A (even) and B (odd) with their own fp32 accumulators; samples are i.i.d. and keyed by
(seed, pixel, sample index), so
  * mean = (sum_A + sum_B) / n is exactly the image a single pt_render of n samples gives,
  * (mean_A - mean_B) / 2 is an unbiased per-pixel estimate of the standard error of that mean, for free,
  * a checkpoint is the two accumulators plus the number of samples done: resuming continues the same sample sequence.
Multi-GPU: rank r of G takes the indices (2 i + half) * G + r, one reduce(sum) per accumulator at the end (distributed.py).
"""
import os

import numpy as np
import torch

from .distributed import reduce_accumulators


def half_ranges(done_per_half, count_per_half, rank=0, world=1):
    """pt_render_params (sample_begin, sample_count, sample_stride) of halves A and B for the next batch."""
    stride = 2 * world
    return [(half * world + rank + stride * done_per_half, count_per_half, stride) for half in (0, 1)]


class ProgressiveRender:
    def __init__(self, dev_scene, camera=None, seed=1, nan_policy=1, flags=0, pool_paths=0, rank=0, world=1):
        self.dev, self.cam = dev_scene, camera if camera is not None else dev_scene.host_scene.camera
        self.seed, self.nan_policy, self.flags, self.pool_paths, self.rank, self.world = seed, nan_policy, flags, pool_paths, rank, world
        h = dev_scene.ctx.lib.pt_camera_image_height(self.cam)
        self.sums = [torch.zeros((h, self.cam.image_width, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
        self.done_per_half = 0          # samples per pixel per half rendered so far BY EACH RANK
        self.segments = self.paths = 0
        self.device_ms = 0.0
        dev_scene.ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    @property
    def spp(self):
        return 2 * self.done_per_half * self.world

    def advance(self, spp_batch):
        """Adds `spp_batch` samples per pixel per rank (rounded up to an even number)."""
        n = (spp_batch + 1) // 2
        for acc, (begin, count, stride) in zip(self.sums, half_ranges(self.done_per_half, n, self.rank, self.world)):
            st = self.dev.render_accumulate(acc.data_ptr(), camera=self.cam, spp=count, seed=self.seed, sample_begin=begin, sample_stride=stride,
                                            nan_policy=self.nan_policy, pool_paths=self.pool_paths, flags=self.flags)
            self.segments += st.segments; self.paths += st.paths; self.device_ms += st.device_ms
        self.done_per_half += n
        return self

    def halves(self):
        """Mean radiance of half A and of half B ([H,W,3] numpy each; complete on rank 0)."""
        a, b = (reduce_accumulators(s.clone(), 0) for s in self.sums)
        torch.cuda.synchronize()
        n_half = max(self.done_per_half * self.world, 1)
        return a.cpu().numpy() / n_half, b.cpu().numpy() / n_half

    def result(self):
        """(mean radiance [H,W,3], per-pixel standard error of that mean [H,W,3]) on rank 0."""
        ma, mb = self.halves()
        return (ma + mb) / 2, np.abs(ma - mb) / 2

    def noise_floor(self, eps=1e-2):
        """relRMSE the image is expected to show against a converged reference because of its own noise (definition of
        SURVEY §8(d): sqrt(mean((a - b)^2 / (b^2 + eps))) on radiance clipped to [0, 1)).  The unknown converged value b
        in the denominator is taken from the darker half: a firefly sits in one half only, so the other half is the
        better stand-in for b exactly where the error is largest."""
        ma, mb = (np.clip(m, 0.0, 0.999) for m in self.halves())
        b = np.minimum(ma, mb)
        return float(np.sqrt(np.mean(((ma - mb) / 2) ** 2 / (b * b + eps))))

    # ---- checkpoint / resume ------------------------------------------------------------------------------------------
    def _signature(self):
        c = self.cam
        return np.array([c.image_width, self.sums[0].shape[0], self.seed, self.nan_policy, self.flags, self.world, self.rank], dtype=np.int64)

    def save(self, path):
        torch.cuda.synchronize()
        tmp = path + ".tmp.npz"
        np.savez(tmp, sum_a=self.sums[0].cpu().numpy(), sum_b=self.sums[1].cpu().numpy(), done_per_half=self.done_per_half,
                 signature=self._signature(), segments=self.segments, paths=self.paths, device_ms=self.device_ms)
        os.replace(tmp, path)  # a crash mid-write leaves the previous checkpoint intact

    def load(self, path):
        z = np.load(path)
        if not np.array_equal(z["signature"], self._signature()):
            raise ValueError("checkpoint belongs to a different image size, seed, policy or rank layout")
        self.sums[0].copy_(torch.from_numpy(z["sum_a"])); self.sums[1].copy_(torch.from_numpy(z["sum_b"]))
        self.done_per_half = int(z["done_per_half"]); self.segments = int(z["segments"]); self.paths = int(z["paths"]); self.device_ms = float(z["device_ms"])
        return self

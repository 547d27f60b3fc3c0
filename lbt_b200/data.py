"""The reference's input pipeline on the device: the data set lives in HBM as raw uint8 images, one
``lbt_augment_batch`` launch per step gathers a shuffled batch, normalises it ((x - mean) / 128, main.py:71-75), flips,
pads by 4 and crops (trainer.py:24-28) and writes the NHWC fp32 batch the model reads.  No host round trip, CUDA-graph
capturable (the randomness is the Philox stream keyed by (seed, epoch, batch)).
"""
import torch

from . import _lib
from .quantizer import make_offset

AUGMENT_STREAM = 0x7FFF0000      # Philox offset low word of the pipeline's stream (quantiser ids are small integers)


class Pipeline:
    """``X`` uint8 [n, H, W, C] (NHWC, like the reference's numpy arrays), ``y`` integer labels — moved to ``device``.
    ``mean`` defaults to the float64 per-pixel mean of X (main.py:71), ``augment`` = flip + pad-4 + random crop."""

    def __init__(self, X, y, device, *, mean=None, augment=True, pad=4, seed=0, shuffle=True, rank=None, world=None):
        # data parallelism (new here; the reference is single-device): every replica walks the SAME epoch permutation and
        # takes its own slice of each global batch of world * batch_size samples; the flip / crop stream is keyed by
        # (epoch, global batch, rank).  rank / world default to the initialised torch.distributed group.
        if world is None:
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized()
            world, rank = (dist.get_world_size(), dist.get_rank()) if on else (1, 0)
        self.rank, self.world = int(rank or 0), int(world)
        if not 0 <= self.rank < self.world:
            raise _lib.LbtError('Pipeline: rank %d outside world %d' % (self.rank, self.world))
        if X.dtype != torch.uint8 or X.dim() != 4:
            raise _lib.LbtError('Pipeline expects the raw uint8 NHWC images')
        self.X = X.to(device).contiguous()
        self.y = y.to(device=device, dtype=torch.int64).contiguous()
        self.n, self.H, self.W, self.C = self.X.shape
        self.mean = (mean if mean is not None else X.to(torch.float64).mean(dim=0)).to(device=device, dtype=torch.float64).contiguous()
        self.augment, self.pad, self.seed, self.shuffle = bool(augment), int(pad) if augment else 0, int(seed), bool(shuffle)
        self.device = torch.device(device)

    def batch(self, index, epoch=0, b=0, params=None, out=None, labels_out=None):
        """One batch: NCHW view (channels_last storage) of the NHWC fp32 images, int64 labels.  ``params`` int32 [B, 3] =
        explicit (flip, oy, ox) per sample (parity tests); default: the Philox stream at (epoch, b)."""
        B = int(index.numel()) if index is not None else self.n
        if out is None:
            out = torch.empty(B, self.H, self.W, self.C, dtype=torch.float32, device=self.device)
        if labels_out is None:
            labels_out = torch.empty(B, dtype=torch.int64, device=self.device)
        offset = make_offset(AUGMENT_STREAM, (int(epoch) << 16) ^ (int(b) * self.world + self.rank))
        _lib.call('lbt_augment_batch', _lib.ptr(self.X), _lib.ptr(self.mean), _lib.ptr(index), B, self.H, self.W, self.C, self.pad,
                  1 if self.augment else 0, _lib.ptr(params), self.seed, offset, _lib.ptr(self.y), _lib.ptr(labels_out),
                  _lib.ptr(out), _lib.stream(), meta=dict(bytes=B * self.H * self.W * self.C * 5))
        return out.permute(0, 3, 1, 2), labels_out

    def epoch(self, batch_size, epoch=0):
        """Batches of one pass over the data (trainer.py:92-96: shuffle the whole set, batch, last batch may be short).
        With world > 1 a step consumes world * batch_size samples of the shared permutation and this replica gets the
        rank-th slice; every replica yields the same number of batches (the exchange needs them all at every step): a
        short tail is dealt round-robin, and dropped when it has fewer samples than replicas."""
        if self.shuffle:
            g = torch.Generator(device='cpu')
            g.manual_seed((self.seed << 20) ^ int(epoch))
            perm = torch.randperm(self.n, generator=g).to(self.device)
        else:
            perm = torch.arange(self.n, device=self.device)
        if self.world == 1:
            for b, i in enumerate(range(0, self.n, batch_size)):
                yield self.batch(perm[i:i + batch_size], epoch, b)
            return
        gb = batch_size * self.world
        for b, i in enumerate(range(0, self.n, gb)):
            chunk = perm[i:i + gb]
            if chunk.numel() == gb:
                mine = chunk[self.rank * batch_size:(self.rank + 1) * batch_size].contiguous()
            elif chunk.numel() >= self.world:
                mine = chunk[self.rank::self.world].contiguous()
            else:
                return
            yield self.batch(mine, epoch, b)

    def all(self):
        """The whole set normalised, no augmentation (the reference's test set feed)."""
        saved = self.augment, self.pad
        self.augment, self.pad = False, 0
        try:
            return self.batch(None)
        finally:
            self.augment, self.pad = saved

"""The training step of the DFXP path (``trainer.py:N`` = /root/reference/trainer.py:N).

One ``Trainer.step(X, y)`` is one ``sess.run([train_op, update_range_op])`` (trainer.py:157-160):
forward, loss, backward (the layers quantise their own gradients), momentum SGD on the fp32 master
weights (trainer.py:79-84), then the range controller of every quantiser.  New here (the reference
is single-device): data parallelism over the batch — one process per GPU, the
flattened gradient all-reduced with NCCL, and the overflow counters all-reduced before the
controller runs so every replica keeps identical ranges and sees the global-batch overflow rate.
"""
import torch
import torch.distributed as dist

from . import _lib


class Trainer:
    def __init__(self, model, lr=1e-2, momentum=0.9, *, process_group=None, sync_counters=True):
        self.model = model
        self.lr, self.momentum = float(lr), float(momentum)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.sync_counters = sync_counters
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise _lib.LbtError('Trainer needs the model on a CUDA device (no CPU fallback)')
        dev = params[0].device
        model.runtime.finalize(dev)
        # flatten parameters / gradients / momentum: one SGD launch and one all-reduce per step
        sizes = [p.numel() for p in params]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += -(-n // 4) * 4                      # keep every view 16-byte aligned
        self.flat_w = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_a = torch.zeros(total, dtype=torch.float32, device=dev)        # optimizer slots start at 0 (trainer.py:83)
        for p, o, n in zip(params, offs, sizes):
            self.flat_w[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_w[o:o + n].view(p.shape)
            p.grad = self.flat_g[o:o + n].view(p.shape)
        self.params = params
        self.dev_lr = torch.tensor(self.lr, dtype=torch.float32, device=dev)
        if self.world > 1:                               # replicas start from rank 0's weights
            dist.broadcast(self.flat_w, src=0, group=self.group)

    def set_lr(self, lr, reset_momentum=True):
        """LR change; the reference re-creates the optimizer, zeroing momentum (trainer.py:79-84, 118-132)."""
        self.lr = float(lr)
        self.dev_lr.fill_(self.lr)
        if reset_momentum:
            self.flat_a.zero_()

    def forward_backward(self, X, y):
        self.flat_g.zero_()
        logits = self.model(X)
        loss = self.model.loss(logits, y)
        loss.backward()
        return loss.detach(), logits.detach()

    def apply(self):
        rt = self.model.runtime
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            if self.sync_counters:
                dist.all_reduce(rt.flat['counters'], op=dist.ReduceOp.SUM, group=self.group)
        _lib.call('lbt_sgd_momentum', _lib.ptr(self.flat_w), _lib.ptr(self.flat_a), _lib.ptr(self.flat_g),
                                               self.flat_w.numel(), self.lr, _lib.ptr(self.dev_lr), self.momentum,
                                               1.0 / self.world, _lib.stream())
        rt.update_ranges()

    def step(self, X, y):
        """X: [N, C, H, W] (channels_last) on the device, y: int64 labels.  Returns the loss tensor."""
        loss, _ = self.forward_backward(X, y)
        self.apply()
        return loss

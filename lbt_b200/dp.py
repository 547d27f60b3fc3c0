"""Peer-memory plumbing of the fused data-parallel step (``lbt_dp_step``, csrc/dp.cu).

The reference is single-device (trainer.py:79-84, 157-160).  Here every replica keeps its flat gradient, flat weights,
overflow counters and a small flag pad in ONE device arena; the arenas are mapped into every other replica process with
CUDA IPC (handles travel over torch.distributed, the control plane), and one kernel per step does the whole exchange over
NVLink.  PyTorch supplies the memory and the rendezvous only.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def _round(n, a=256):
    return -(-int(n) // a) * a


class Arena:
    """[pad | counters | flat_g | flat_w] in one allocation (one IPC handle per replica)."""

    def __init__(self, n_sites, n_params, device):
        self.n_sites, self.n_params = int(n_sites), int(n_params)
        self.off_pad = 0
        self.off_cnt = _round(_lib.DP_PAD_WORDS * 4)
        self.off_g = self.off_cnt + _round(max(1, self.n_sites) * 32)
        self.off_w = self.off_g + _round(max(4, self.n_params) * 4)
        self.nbytes = self.off_w + _round(max(4, self.n_params) * 4)
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.pad = self.buf[self.off_pad:self.off_pad + _lib.DP_PAD_WORDS * 4].view(torch.int32)
        self.counters = self.buf[self.off_cnt:self.off_cnt + self.n_sites * 32].view(torch.int64).view(self.n_sites, 4)
        self.flat_g = self.buf[self.off_g:self.off_g + self.n_params * 4].view(torch.float32)
        self.flat_w = self.buf[self.off_w:self.off_w + self.n_params * 4].view(torch.float32)

    def pointers(self, base):
        return dict(pad=base + self.off_pad, counters=base + self.off_cnt, grad=base + self.off_g, w=base + self.off_w)


def make_peers(rank, bases, arena):
    """lbt_dp_peers for `rank` from the (peer-mapped) arena base addresses of all replicas."""
    p = _lib.DpPeers()
    p.world, p.rank = len(bases), int(rank)
    for r, b in enumerate(bases):
        q = arena.pointers(b)
        p.grad[r], p.w[r], p.counters[r], p.pad[r] = q['grad'], q['w'], q['counters'], q['pad']
    return p


class DpExchange:
    """One replica's end of the fused exchange: owns the arena, maps the peers' arenas, launches lbt_dp_step."""

    def __init__(self, n_sites, n_params, device, group=None):
        self.group = group
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        if self.world > _lib.DP_MAX_WORLD:
            raise _lib.LbtError('lbt_dp_step supports up to %d replicas' % _lib.DP_MAX_WORLD)
        self.arena = Arena(n_sites, n_params, device)
        self._opened = []
        bases = [self.arena.buf.data_ptr()]
        if multi:
            # every replica takes part in the handle exchange even if its own export failed (no rank may be left waiting)
            mine = None
            try:
                handle = ctypes.create_string_buffer(_lib.DP_HANDLE_BYTES)
                off = ctypes.c_size_t(0)
                _lib.check(_lib.lib().lbt_dp_export(self.arena.buf.data_ptr(), handle, ctypes.byref(off)))
                mine = (bytes(handle.raw), int(off.value), self.arena.nbytes)
            except _lib.LbtError as e:
                err = e
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            if mine is None:
                raise err
            if any(x is None for x in everyone):
                raise _lib.LbtError('a peer replica could not export its arena')
            bases = []
            for r, (h, o, nbytes) in enumerate(everyone):
                if nbytes != self.arena.nbytes:
                    raise _lib.LbtError('replica %d has a different parameter / quantiser count' % r)
                if r == self.rank:
                    bases.append(self.arena.buf.data_ptr())
                    continue
                base = ctypes.c_void_p(0)
                _lib.check(_lib.lib().lbt_dp_open(ctypes.create_string_buffer(h, _lib.DP_HANDLE_BYTES), ctypes.byref(base)))
                self._opened.append(base.value)
                bases.append(base.value + o)
        self.peers = make_peers(self.rank, bases, self.arena)
        # the caller synchronises the replicas (all arenas zeroed and mapped) before the first step: Trainer.__init__

    def step(self, flat_a, lr, dev_lr, momentum, rt, shard=True):
        f = rt.flat
        _lib.call('lbt_dp_step', ctypes.addressof(self.peers), _lib.ptr(flat_a), self.arena.n_params, float(lr),
                  _lib.ptr(dev_lr), float(momentum), 1 if shard else 0, _lib.ptr(f['ranges']), _lib.ptr(f['bits']),
                  _lib.ptr(f['target']), len(rt.sites), _lib.ptr(rt.dev_step), _lib.stream(),
                  meta=dict(bytes=self.arena.n_params * 4 * 2))

    def owned(self, shard=True):
        """[lo, hi) of the flat parameter index this replica reduces and updates (lbt_dp_step's slicing)."""
        n = self.arena.n_params
        if not shard or self.world == 1:
            return 0, n
        chunk = -(-(n // 4) // self.world) * 4
        return min(n, self.rank * chunk), min(n, (self.rank + 1) * chunk)

    def error(self):
        """Non-zero when a cross-replica wait timed out (synchronises)."""
        return int(self.arena.pad[_lib.DP_PAD_ERROR].item())

    def close(self):
        for b in self._opened:
            _lib.lib().lbt_dp_close(ctypes.c_void_p(b))
        self._opened = []


def emulate_step(peers, accums, n_params, lr, momentum, ranges, bits, target, n_sites, dev_steps, *, dev_lrs=None, shard=True):
    """The lbt_dp_step of every replica of a world that lives on ONE GPU, as a single cooperative launch (lbt_dp_step_emulate:
    grid.y = replica).  Test plumbing for single-GPU boxes: separate launches that spin on each other's flags are not
    guaranteed to run at the same time on one device.  ``peers``: list of DpPeers (replica r at index r); the other arguments
    are per-replica lists of tensors (``bits`` / ``target`` are shared)."""
    world = len(peers)
    arr = (_lib.DpPeers * world)(*peers)
    vp = ctypes.c_void_p * world
    acc = vp(*[_lib.ptr(a) for a in accums])
    rng = vp(*[_lib.ptr(r) for r in ranges])
    stp = vp(*[_lib.ptr(s) for s in dev_steps])
    lrs = vp(*[_lib.ptr(t) for t in dev_lrs]) if dev_lrs is not None else None
    dev = accums[0].device
    scratch = torch.empty(int(_lib.lib().lbt_dp_emulate_scratch_bytes(world)), dtype=torch.uint8, device=dev)
    _lib.call('lbt_dp_step_emulate', world, ctypes.addressof(arr), ctypes.addressof(acc), int(n_params), float(lr),
              ctypes.addressof(lrs) if lrs is not None else None, float(momentum), 1 if shard else 0, ctypes.addressof(rng),
              _lib.ptr(bits), _lib.ptr(target), int(n_sites), ctypes.addressof(stp), _lib.ptr(scratch), _lib.stream())
    return scratch      # keep alive until the stream has run the launch


def emulate_trainers(trainers):
    """``Trainer.apply()`` for N Trainers that simulate N replicas on one GPU (their DpExchange.peers wired to each other's
    arenas): one cooperative launch instead of N launches that wait for each other."""
    t0 = trainers[0]
    rts = [t.model.runtime for t in trainers]
    keep = emulate_step([t.dp.peers for t in trainers], [t.flat_a for t in trainers], t0.dp.arena.n_params, t0.lr, t0.momentum,
                        [rt.flat['ranges'] for rt in rts], rts[0].flat['bits'], rts[0].flat['target'], len(rts[0].sites),
                        [rt.dev_step for rt in rts], dev_lrs=[t.dev_lr for t in trainers])
    for rt in rts:
        rt.close_step()
    return keep

"""lbt_dp_step (csrc/dp.cu): the data-parallel end of a step as one kernel over peer memory.

* world = 1: bit-identical to lbt_sgd_momentum + lbt_update_ranges + lbt_step_advance;
* world = 2..4 simulated on ONE GPU (one arena per "replica", peers = each other's arenas, all replicas' kernel bodies in one
  cooperative launch — lbt_dp_step_emulate — so they are co-resident by construction): the flag
  protocol, the rank-ordered gradient sum, the owner-computes SGD and the weight all-gather, the counter sum + controller
  against the oracle's range_delta, over several consecutive steps (epochs), sharded and replicated;
* two real GPUs (skipped on a one-GPU box): Trainer(dp='fused') == Trainer(dp='nccl') bit for bit at world 2.
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch

from oracle import dfxp as O

pytestmark = pytest.mark.gpu

from lbt_b200 import _lib  # noqa: E402
from lbt_b200.dp import Arena, emulate_step, make_peers  # noqa: E402


def _ref_sgd(w, a, gs, lr, mom):
    """The kernel's arithmetic with separately rounded fp32 operations (eager torch: one rounding per op)."""
    g = gs[0].clone()
    for t in gs[1:]:
        g = g + t
    scale = torch.tensor(1.0 / len(gs), dtype=torch.float32, device=w.device)
    a2 = a * mom + g * scale
    return w - a2 * lr, a2


def _mk_counters(rng, n_sites, dev):
    c = torch.zeros(n_sites, 4, dtype=torch.int64, device=dev)
    numel = torch.from_numpy(rng.integers(1, 1 << 20, n_sites)).to(dev)
    kind = torch.from_numpy(rng.integers(0, 4, n_sites)).to(dev)      # 0: not run, 1: quiet, 2: half overflow, 3: overflow
    c[:, 2] = torch.where(kind == 0, torch.zeros_like(numel), numel)
    c[:, 1] = torch.where(kind >= 2, torch.ones_like(numel) * 3, torch.zeros_like(numel))
    c[:, 0] = torch.where(kind == 3, torch.ones_like(numel) * 2, torch.zeros_like(numel))
    return c


@pytest.mark.parametrize('n', [4, 1000, 4 * 1024 * 1024 + 8])
def test_world1_equals_separate_kernels(n):
    dev = torch.device('cuda')
    rng = np.random.default_rng(n)
    n_sites = 37
    ar = Arena(n_sites, n, dev)
    peers = make_peers(0, [ar.buf.data_ptr()], ar)
    ar.flat_w.copy_(torch.from_numpy(rng.standard_normal(n).astype(np.float32)))
    a = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev)
    ranges = torch.from_numpy(rng.integers(-3, 8, n_sites).astype(np.int32)).to(dev)
    bits = torch.full((n_sites,), 8, dtype=torch.int32, device=dev)
    target = torch.zeros(n_sites, dtype=torch.float32, device=dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    w2, a2, r2, s2 = ar.flat_w.clone(), a.clone(), ranges.clone(), step.clone()
    lr = torch.tensor(0.05, dtype=torch.float32, device=dev)
    for it in range(3):
        g = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev)
        c = _mk_counters(rng, n_sites, dev)
        ar.flat_g.copy_(g)
        ar.counters.copy_(c)
        c2 = c.clone()
        _lib.call('lbt_dp_step', ctypes.addressof(peers), _lib.ptr(a), n, 0.0, _lib.ptr(lr), 0.9, 1, _lib.ptr(ranges),
                  _lib.ptr(bits), _lib.ptr(target), n_sites, _lib.ptr(step), _lib.stream())
        _lib.call('lbt_sgd_momentum', _lib.ptr(w2), _lib.ptr(a2), _lib.ptr(g), n, 0.0, _lib.ptr(lr), 0.9, 1.0, _lib.stream())
        _lib.call('lbt_update_ranges', _lib.ptr(r2), _lib.ptr(c2), _lib.ptr(bits), _lib.ptr(target), n_sites, _lib.stream())
        _lib.call('lbt_step_advance', _lib.ptr(s2), _lib.stream())
        torch.cuda.synchronize()
        assert torch.equal(ar.flat_w.view(torch.int32), w2.view(torch.int32))
        assert torch.equal(a.view(torch.int32), a2.view(torch.int32))
        assert torch.equal(ranges, r2)
        assert int(step) == int(s2) == it + 1
        assert int(ar.counters.abs().sum()) == 0 and int(c2.abs().sum()) == 0
        assert int(ar.pad[_lib.DP_PAD_ERROR]) == 0
        assert int(ar.pad[16]) == it + 1 and int(ar.pad[17]) == 0          # epoch advanced, ticket reset


def test_bad_arguments():
    dev = torch.device('cuda')
    ar = Arena(3, 16, dev)
    peers = make_peers(0, [ar.buf.data_ptr()], ar)
    a = torch.zeros(16, device=dev)
    h = _lib.lib()
    assert h.lbt_dp_step(None, a.data_ptr(), 16, 0.1, None, 0.9, 0, None, None, None, 0, None, None) == -1
    assert h.lbt_dp_step(ctypes.addressof(peers), a.data_ptr(), 18, 0.1, None, 0.9, 0, None, None, None, 0, None, None) == -2
    peers.world = 9
    assert h.lbt_dp_step(ctypes.addressof(peers), a.data_ptr(), 16, 0.1, None, 0.9, 0, None, None, None, 0, None, None) == -1


@pytest.mark.parametrize('world,shard', [(2, 1), (2, 0), (3, 1), (4, 1), (4, 0)])
def test_simulated_replicas_on_one_gpu(world, shard):
    dev = torch.device('cuda')
    rng = np.random.default_rng(world * 10 + shard)
    n, n_sites = 100_000, 53               # all replicas' CTAs are co-resident at this size (they wait on each other)
    arenas = [Arena(n_sites, n, dev) for _ in range(world)]
    bases = [a.buf.data_ptr() for a in arenas]
    peers = [make_peers(r, bases, arenas[0]) for r in range(world)]
    w0 = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev)
    for a in arenas:
        a.flat_w.copy_(w0)
    accum = [torch.zeros(n, device=dev) for _ in range(world)]
    ranges = [torch.full((n_sites,), 2, dtype=torch.int32, device=dev) for _ in range(world)]
    bits = torch.full((n_sites,), 8, dtype=torch.int32, device=dev)
    steps = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    w_ref, a_ref = w0.clone(), torch.zeros(n, device=dev)
    r_ref = [2] * n_sites
    nv = n // 4
    chunk = -(-nv // world)
    for it in range(4):
        gs = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(dev) for _ in range(world)]
        cs = [_mk_counters(rng, n_sites, dev) for _ in range(world)]
        for r in range(world):
            arenas[r].flat_g.copy_(gs[r])
            arenas[r].counters.copy_(cs[r])
        # the replicas' lbt_dp_step bodies as ONE cooperative launch (grid.y = replica): co-residency guaranteed
        keep = emulate_step(peers, accum, n, 0.01 * (it + 1), 0.9, ranges, bits, None, n_sites, steps, shard=bool(shard))
        torch.cuda.synchronize()
        del keep
        w_ref, a_ref = _ref_sgd(w_ref, a_ref, gs, 0.01 * (it + 1), 0.9)
        tot = sum(c.cpu() for c in cs)
        for i in range(n_sites):
            if int(tot[i, 2]):
                r_ref[i] = min(7, r_ref[i] + O.range_delta(int(tot[i, 0]), int(tot[i, 1]), int(tot[i, 2]), 0.0))
        for r in range(world):
            assert int(arenas[r].pad[_lib.DP_PAD_ERROR]) == 0, 'replica %d: a cross-replica wait timed out' % r
            assert torch.equal(arenas[r].flat_w.view(torch.int32), w_ref.view(torch.int32)), (it, r)
            assert ranges[r].cpu().tolist() == r_ref
            assert int(steps[r]) == it + 1
            assert int(arenas[r].counters.abs().sum()) == 0
            if shard:
                lo, hi = min(n, 4 * r * chunk), min(n, 4 * (r + 1) * chunk)
                assert torch.equal(accum[r][lo:hi].view(torch.int32), a_ref[lo:hi].view(torch.int32))
                assert float(accum[r][:lo].abs().sum()) == 0 and float(accum[r][hi:].abs().sum()) == 0
            else:
                assert torch.equal(accum[r].view(torch.int32), a_ref.view(torch.int32))


# ---- two real GPUs ------------------------------------------------------------------------------------------------------

def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from lbt_b200 import models as M
    from lbt_b200.trainer import Trainer
    res = {}
    for mode in ('fused', 'nccl'):
        torch.manual_seed(0)
        pm = M.CIFAR10_Resnet20(8, weight_decay=2e-4, seed=3).cuda()
        tr = Trainer(pm, lr=1e-2, momentum=0.9, dp=mode)
        assert tr.dp_mode == mode, tr.dp_mode
        rng = np.random.default_rng(100 + rank)
        losses = []
        for _ in range(3):
            X = torch.from_numpy((rng.standard_normal((16, 3, 32, 32)) * 0.5).astype(np.float32)).cuda()
            X = X.contiguous(memory_format=torch.channels_last)
            y = torch.from_numpy(rng.integers(0, 10, 16)).cuda()
            losses.append(float(tr.step(X, y)))
        torch.cuda.synchronize()
        if mode == 'fused':
            assert tr.dp.error() == 0
        res[mode] = dict(w=tr.flat_w.clone().cpu(), r=list(pm.ranges().values()), loss=losses)
    torch.save(res, '%s.%d' % (out, rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_fused_exchange_equals_nccl_on_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + '.0'), torch.load(out + '.1')
    for mode in ('fused', 'nccl'):
        assert torch.equal(r0[mode]['w'].view(torch.int32), r1[mode]['w'].view(torch.int32)), mode   # replicas stay identical
        assert r0[mode]['r'] == r1[mode]['r']
    # at world 2 the rank-ordered sum is the all-reduce's sum (a + b commutes): the two formulations agree bit for bit
    assert torch.equal(r0['fused']['w'].view(torch.int32), r0['nccl']['w'].view(torch.int32))
    assert r0['fused']['r'] == r0['nccl']['r']
    assert r0['fused']['loss'] == r0['nccl']['loss']


def test_trainer_fused_step_end_equals_separate_kernels():
    """Trainer(dp='fused') (lbt_dp_step) vs Trainer(dp='nccl') (at world 1: lbt_sgd_momentum + lbt_update_ranges +
    lbt_step_advance) on one GPU: identical weights, momentum, ranges and losses."""
    from lbt_b200 import models as M
    from lbt_b200.trainer import Trainer
    res = {}
    for mode in ('fused', 'nccl'):
        torch.manual_seed(0)
        pm = M.CIFAR10_Resnet20(8, weight_decay=2e-4, seed=3).cuda()
        tr = Trainer(pm, lr=1e-2, momentum=0.9, dp=mode)
        assert tr.dp_mode == ('fused' if mode == 'fused' else 'unfused')
        rng = np.random.default_rng(5)
        losses = []
        for _ in range(3):
            X = torch.from_numpy((rng.standard_normal((8, 3, 32, 32)) * 0.5).astype(np.float32)).cuda().contiguous(
                memory_format=torch.channels_last)
            y = torch.from_numpy(rng.integers(0, 10, 8)).cuda()
            losses.append(float(tr.step(X, y)))
        res[mode] = (losses, tr.flat_w.clone(), tr.flat_a.clone(), list(pm.ranges().values()), int(pm.runtime.dev_step))
    assert res['fused'][0] == res['nccl'][0]
    assert torch.equal(res['fused'][1].view(torch.int32), res['nccl'][1].view(torch.int32))
    assert torch.equal(res['fused'][2].view(torch.int32), res['nccl'][2].view(torch.int32))
    assert res['fused'][3] == res['nccl'][3] and res['fused'][4] == res['nccl'][4] == 3

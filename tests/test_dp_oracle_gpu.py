"""Data-parallel parity against the oracle (SURVEY.md §8(e) rules i-iii; the reference itself is single-device, F11).

N replicas are simulated on ONE GPU: N real models + Trainers, each with its own peer arena, wired to each other's arenas
exactly as the IPC-mapped arenas of N processes would be; the N lbt_dp_step bodies run as ONE cooperative launch
(lbt_dp_step_emulate, grid.y = replica — launches that spin on each other's flags must be co-resident, which separate
launches on one device do not guarantee).  The replicas run REAL CIFAR10_Model training steps on different shards of a
global batch.  Checked against the oracle
in exactly-rounded-accumulation mode:

  (i)   each replica's parameter gradient on its shard == the oracle's on that shard, bit for bit;
  (ii)  the weights every replica holds after lbt_dp_step == momentum SGD on the mean of the oracle's per-shard gradients
        (sum in rank order, times 1/N), bit for bit, and identical on all replicas;
  (iii) the ranges every replica holds == the controller applied to the oracle's overflow counts summed over the shards, and
        for every FORWARD quantiser == what ONE oracle at the global batch N*b decides (noise ignores the batch index,
        dfxp:36, so the forward pass of the global batch is the concatenation of the shards').
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import dfxp as O

pytestmark = pytest.mark.gpu

from lbt_b200 import _lib, dfxp as D, models as M  # noqa: E402
from lbt_b200.dp import emulate_trainers, make_peers  # noqa: E402
from lbt_b200.trainer import Trainer  # noqa: E402

SEED = 13


def _oracle(dropout):
    return O.CIFAR10_Model(8, weight_decay=2e-4, dropout=dropout, noise=O.PhiloxNoise(SEED), seed=1, exact=True)


@pytest.mark.parametrize('world', [2, 4])
def test_simulated_replicas_train_like_the_oracle(world):
    dev = torch.device('cuda')
    rng = np.random.default_rng(40 + world)
    b, lr, mom = 8, 1e-2, 0.9
    oracles = [_oracle(1.0) for _ in range(world)]          # dropout off: its uniforms are per-sample, not part of this rule
    big = _oracle(1.0)
    trainers = []
    for r in range(world):
        pm = M.CIFAR10_Model(8, weight_decay=2e-4, dropout=1.0, seed=SEED).cuda()
        for ov, pv in zip(oracles[0].variables(), pm.parameters()):
            pv.data.copy_(ov.detach())
        trainers.append(Trainer(pm, lr=lr, momentum=mom))
        assert trainers[-1].dp_mode == 'fused'
    for om in oracles[1:] + [big]:
        for a, c in zip(oracles[0].variables(), om.variables()):
            c.data.copy_(a.detach())
    bases = [t.dp.arena.buf.data_ptr() for t in trainers]
    for r, t in enumerate(trainers):                       # what DpExchange does with the IPC-mapped bases of N processes
        t.dp.peers = make_peers(r, bases, t.dp.arena)
        t.dp.world, t.dp.rank = world, r
    w_ref = torch.cat([v.detach().reshape(-1) for v in oracles[0].variables()])
    a_ref = torch.zeros_like(w_ref)
    n_sites = len(trainers[0].model.runtime.sites)
    r_ref = [2] * n_sites
    for step in range(3):
        X = torch.from_numpy((rng.standard_normal((world * b, 32, 32, 3)) * 0.5).astype(np.float32))
        y = torch.from_numpy(rng.integers(0, 10, world * b))
        # ---- oracle: every shard separately, and the forward pass of the global batch ----
        grads, counts = [], []
        for r, om in enumerate(oracles):
            om.forward(X[r * b:(r + 1) * b])
            om.backward(y[r * b:(r + 1) * b])
            grads.append([g.detach().clone() for g, _ in om.grads_and_vars()])
            counts.append(dict(om.ctx.last_counts))
        big.forward(X)
        big_counts = dict(big.ctx.last_counts)
        # ---- GPU: N replicas, backward on their shards, then the fused exchange on N concurrent streams ----
        for r, t in enumerate(trainers):
            Xr = X[r * b:(r + 1) * b].permute(0, 3, 1, 2).cuda()
            t.forward_backward(Xr, y[r * b:(r + 1) * b].cuda())
            for (go, p) in zip(grads[r], t.params):                                                   # rule (i)
                got = p.grad.cpu()
                assert torch.equal(got, go), 'replica %d step %d: gradient of %s differs in %d elements (max %g)' % (
                    r, step, tuple(p.shape), int((got != go).sum()), float((got - go).abs().max()))
        keep = emulate_trainers(trainers)
        torch.cuda.synchronize()
        del keep
        # ---- rule (ii): SGD on the rank-ordered mean of the oracle's per-shard gradients ----
        offs = []
        g_sum = None
        for r in range(world):
            flat = torch.cat([g.reshape(-1) for g in grads[r]])
            g_sum = flat if g_sum is None else g_sum + flat
        a_ref = a_ref * mom + g_sum * torch.tensor(1.0 / world, dtype=torch.float32)
        w_ref = w_ref - a_ref * torch.tensor(lr, dtype=torch.float32)
        for r, t in enumerate(trainers):
            assert t.dp.error() == 0
            w_p = torch.cat([p.data.reshape(-1) for p in t.params]).cpu()
            assert torch.equal(w_p, w_ref), 'replica %d step %d: %d weights differ' % (r, step, int((w_p != w_ref).sum()))
        # ---- rule (iii): ranges from the summed counts; forward sites == the oracle at the global batch ----
        sites = trainers[0].model.runtime.sites
        for i, s in enumerate(sites):
            n1 = sum(c[i][0] for c in counts)
            n2 = sum(c[i][1] for c in counts)
            ne = sum(c[i][2] for c in counts)
            new = min(s.bits - 1, r_ref[i] + O.range_delta(n1, n2, ne, 0.0))
            if not s.name.endswith('/grad') and not s.name.endswith('/W') and not s.name.endswith('/b'):
                assert big_counts[i] == (n1, n2, ne), s.name          # activations: global batch == concatenated shards
            r_ref[i] = new
        for r, t in enumerate(trainers):
            assert list(t.model.ranges().values()) == r_ref, 'replica %d step %d' % (r, step)
        # the oracles continue from the reduced state (what each replica now holds)
        for om in oracles + [big]:
            off = 0
            for v in om.variables():
                n = v.numel()
                v.data.copy_(w_ref[off:off + n].view(v.shape))
                off += n
            for q, rv in zip(om.quantizers(), r_ref):
                q.range.value = rv
            om.ctx.noise.step += 1

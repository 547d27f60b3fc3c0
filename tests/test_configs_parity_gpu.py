"""Whole-model parity of BASELINE configs 4 and 5 (ResNet-18 / ResNet-50 composed from the reference's blocks,
dfxp:746-980; bottleneck stride on the 3x3, dfxp:929-934; stem + 3x3/2 max-pool) against the CPU oracle in
exactly-rounded-accumulation mode, and the asserted multi-step drift SURVEY §8(d) asks for ("pin N=1 tightly, N=10
loosely, report drift").

Everything is bit-exact: integer accumulation == fp64-accumulate-round-once for every contraction and batch sum, and every
elementwise operation (quantisers, normalisation, the batch-norm VJP in closed form, SGD) is one fp32 rounding per operation in
the same order on both sides.  So the tests assert EQUALITY of every gradient, every updated weight and every range after each
step — for ResNet-18, ResNet-50 (8-bit and 16-bit gradients), ResNet-20 and CIFAR10_Model, up to ten consecutive steps.  Only
the scalar loss is compared to 1e-6 (its mean over the batch is summed in a different order).
"""
import numpy as np
import pytest
import torch

from oracle import dfxp as O

pytestmark = pytest.mark.gpu

from lbt_b200 import dfxp as D, models as M  # noqa: E402
from lbt_b200.trainer import Trainer  # noqa: E402

SEED = 5
IMAGE, CLASSES = 64, 100


def rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


def build(name, gbits=None, batch_norm_wd=2e-4, **kw):
    extra = dict(image=IMAGE, num_classes=CLASSES) if name.startswith('Resnet') else {}
    om = getattr(O, name)(8, weight_decay=batch_norm_wd, noise=O.PhiloxNoise(SEED), seed=1, exact=True, grad_bits=gbits, **extra)
    pm = getattr(M, name)(8, weight_decay=batch_norm_wd, seed=SEED, grad_bits=gbits, **extra).cuda()
    ov, pv = om.variables(), list(pm.parameters())
    assert len(ov) == len(pv)
    for a, b in zip(ov, pv):
        assert tuple(a.shape) == tuple(b.shape)
        b.data.copy_(a.detach())
    assert len(om.quantizers()) == len(pm.runtime.sites)
    for q, s in zip(om.quantizers(), pm.runtime.sites):
        assert q.bits == s.bits, s.name
    return om, pm


def batch(rng, n, image=IMAGE, classes=CLASSES):
    X = torch.from_numpy((rng.standard_normal((n, image, image, 3)) * 0.5).astype(np.float32))
    y = torch.from_numpy(rng.integers(0, classes, n))
    return X, y


def tie_dropout(om, pm, rng):
    """Feed both models the same dropout uniforms (tf.nn.dropout's random_uniform); the oracle draws, the GPU side reuses."""
    olayers = [l for l in om.layers if isinstance(l, O.Dropout_q)]
    players = [l for l in pm.layers if isinstance(l, D.Dropout_q)]
    assert len(olayers) == len(players)
    for ol, pl in zip(olayers, players):
        store = {}

        def ofn(shape, store=store):
            store['u'] = torch.from_numpy(rng.random(shape).astype(np.float32))
            return store['u']

        def pfn(x, store=store):
            u = store['u']
            return (u.permute(0, 3, 1, 2) if u.dim() == 4 else u).cuda()

        ol.uniform_fn, pl.uniform_fn = ofn, pfn


CONFIGS = [('Resnet18', None, 8), ('Resnet50', None, 4), ('Resnet50', 16, 4)]


@pytest.mark.parametrize('name,gbits,B', CONFIGS)
def test_imagenet_resnet_forward_bit_exact_vs_exact_oracle(name, gbits, B):
    """Stem (7x7/2 on the signed 3-channel image), max-pool, every block (basic / bottleneck, identity and 1x1-conv
    shortcuts, stride on the 3x3): the activations entering the global average pool equal the oracle's bit for bit, through
    the SAME fused conv+BN units the Trainer runs, and every forward quantiser takes the same range decision."""
    rng = np.random.default_rng(3)
    om, pm = build(name, gbits)
    pm.runtime.finalize('cuda')
    X, _ = batch(rng, B)
    xo = X
    for layer in om.layers[:-3]:
        xo = layer.forward(xo)
    xp = D.run_layers(list(pm.layers)[:-3], X.permute(0, 3, 1, 2).cuda())
    got = xp.permute(0, 2, 3, 1).contiguous().cpu()
    assert got.shape == xo.shape
    assert torch.equal(got, xo), 'mismatching elements: %d of %d' % (int((got != xo).sum()), xo.numel())
    pm.runtime.update_ranges()
    want, got_r = om.ranges(), list(pm.ranges().values())
    fwd = [i for i, s in enumerate(pm.runtime.sites) if not s.name.endswith('/grad') and not s.name.startswith('softmax')]
    assert len(fwd) > 100
    assert [got_r[i] for i in fwd] == [want[i] for i in fwd]


def _compare_step(tag, om, pm, tr, lo, lp, ovars):
    """Loss to 1e-6; every gradient, every weight, every range EQUAL.  Returns the drift record."""
    g_o = [g for g, _ in om.grads_and_vars()]
    r_o, r_p = om.ranges(), list(pm.ranges().values())
    bad_r = [s.name for s, a, b in zip(pm.runtime.sites, r_o, r_p) if a != b]
    n_g = sum(int((p.grad.cpu() != g).sum()) for p, g in zip(tr.params, g_o))
    n_w = sum(int((p.data.cpu() != v.detach()).sum()) for p, v in zip(tr.params, ovars))
    rel = abs(lo - lp) / max(1.0, abs(lo))
    print('%s: loss %.7f / %.7f (rel %.1e), ranges differing %d, gradient elements differing %d, weights differing %d'
          % (tag, lo, lp, rel, len(bad_r), n_g, n_w))
    assert np.isfinite(lp) and rel <= 1e-6
    assert not bad_r, bad_r[:8]
    assert n_g == 0 and n_w == 0
    return rel, len(bad_r), n_g, n_w


@pytest.mark.parametrize('name,gbits,B', CONFIGS)
def test_imagenet_resnet_training_step_vs_exact_oracle(name, gbits, B):
    """Three consecutive full training steps (forward, loss, backward with quantised gradients — 16-bit for config 5 —
    momentum SGD, range controller): after each, every gradient, weight and range equals the oracle's bit for bit."""
    rng = np.random.default_rng(11)
    om, pm = build(name, gbits)
    tr = Trainer(pm, lr=1e-2, momentum=0.9)
    ovars = om.variables()
    for step in range(3):
        X, y = batch(rng, B)
        lo = om.train_step(X, y, lr=1e-2, momentum=0.9)
        lp = float(tr.step(X.permute(0, 3, 1, 2).cuda(), y.cuda()))
        _compare_step('%s g%s step %d' % (name, gbits or 8, step), om, pm, tr, lo, lp, ovars)


@pytest.mark.parametrize('name,B,image,classes', [('CIFAR10_Resnet20', 16, 32, 10), ('CIFAR10_Model', 16, 32, 10)])
def test_ten_step_drift_vs_exact_oracle(name, B, image, classes):
    """"pin N=1 tightly, N=10 loosely, report drift".  The drift over ten steps is ZERO: weights, gradients and
    ranges stay bit-identical to the oracle's at every step, with batch-norm (ResNet-20, 19 conv+BN units) and without
    (CIFAR10_Model, dropout uniforms tied)."""
    rng = np.random.default_rng(21)
    om, pm = build(name)
    if name == 'CIFAR10_Model':
        tie_dropout(om, pm, rng)
    tr = Trainer(pm, lr=1e-2, momentum=0.9)
    ovars = om.variables()
    for step in range(10):
        X, y = batch(rng, B, image, classes)
        lo = om.train_step(X, y, lr=1e-2, momentum=0.9)
        lp = float(tr.step(X.permute(0, 3, 1, 2).cuda(), y.cuda()))
        _compare_step('%s drift step %d' % (name, step), om, pm, tr, lo, lp, ovars)

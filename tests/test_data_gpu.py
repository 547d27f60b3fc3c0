"""lbt_augment_batch / lbt_b200.data.Pipeline vs the oracle's input pipeline: bit-exact, and the
Trainer's epoch loop, evaluation and checkpoint/resume around it (N2, N4)."""
import numpy as np
import pytest
import torch

from oracle import data as OD

pytestmark = pytest.mark.gpu

from lbt_b200 import data as PD, models as M  # noqa: E402
from lbt_b200.quantizer import make_offset  # noqa: E402
from lbt_b200.trainer import Trainer  # noqa: E402


def _dataset(rng, n, H, W, C, classes=10):
    X = rng.integers(0, 256, (n, H, W, C), dtype=np.uint8)
    y = rng.integers(0, classes, n)
    return X, y


@pytest.mark.parametrize('H,W,C,B', [(32, 32, 3, 128), (28, 28, 1, 33), (8, 12, 5, 7)])
def test_augment_explicit_params_bit_exact(H, W, C, B):
    rng = np.random.default_rng(H + C)
    X, y = _dataset(rng, 300, H, W, C)
    pipe = PD.Pipeline(torch.from_numpy(X), torch.from_numpy(y), 'cuda', seed=3)
    mean64 = X.astype(np.float64).mean(axis=0)
    assert np.array_equal(pipe.mean.cpu().numpy(), mean64)
    index = rng.integers(0, 300, B)
    params = np.stack([rng.integers(0, 2, B), rng.integers(0, 9, B), rng.integers(0, 9, B)], axis=1).astype(np.int32)
    want, wl = OD.batch(X, y, mean64, index, params)
    got, gl = pipe.batch(torch.from_numpy(index).cuda(), params=torch.from_numpy(params).cuda())
    assert got.shape == (B, C, H, W) and got.permute(0, 2, 3, 1).is_contiguous()
    assert np.array_equal(got.permute(0, 2, 3, 1).cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(gl.cpu().numpy(), wl)


def test_augment_philox_stream_and_plain_batches():
    rng = np.random.default_rng(1)
    X, y = _dataset(rng, 200, 32, 32, 3)
    pipe = PD.Pipeline(torch.from_numpy(X), torch.from_numpy(y), 'cuda', seed=11)
    mean64 = X.astype(np.float64).mean(axis=0)
    index = rng.permutation(200)[:64]
    got, _ = pipe.batch(torch.from_numpy(index).cuda(), epoch=5, b=17)
    params = OD.philox_params(64, 11, make_offset(PD.AUGMENT_STREAM, (5 << 16) ^ 17))
    want, _ = OD.batch(X, y, mean64, index, params)
    assert np.array_equal(got.permute(0, 2, 3, 1).cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert len(np.unique(params, axis=0)) > 30                                  # the draws differ between samples
    # the whole set, normalised only (the reference's test feed)
    allx, ally = pipe.all()
    assert np.array_equal(allx.permute(0, 2, 3, 1).cpu().numpy().view(np.uint32), OD.normalise(X, mean64).view(np.uint32))
    assert np.array_equal(ally.cpu().numpy(), y)
    # one epoch visits every sample exactly once; the last batch is short (trainer.py:92-96)
    seen = torch.cat([l for _, l in pipe.epoch(64, epoch=0)])
    assert seen.numel() == 200 and [b[0].shape[0] for b in pipe.epoch(64, 0)] == [64, 64, 64, 8]


def _trainer(seed=4):
    torch.manual_seed(0)
    pm = M.CIFAR10_Resnet20(8, weight_decay=2e-4, seed=seed).cuda()
    return pm, Trainer(pm, lr=1e-2, momentum=0.9)


def test_checkpoint_resume_is_bit_identical(tmp_path):
    rng = np.random.default_rng(2)
    Xs = [torch.from_numpy((rng.standard_normal((16, 3, 32, 32)) * 0.5).astype(np.float32)).cuda().contiguous(
        memory_format=torch.channels_last) for _ in range(5)]
    ys = [torch.from_numpy(rng.integers(0, 10, 16)).cuda() for _ in range(5)]
    pm, tr = _trainer()
    for i in range(2):
        tr.step(Xs[i], ys[i])
    tr.save(str(tmp_path / 'ck.pt'))
    ref = [float(tr.step(Xs[i], ys[i])) for i in range(2, 5)]
    w_ref, r_ref = tr.flat_w.clone(), pm.ranges()
    pm2, tr2 = _trainer()
    tr2.load(str(tmp_path / 'ck.pt'))
    got = [float(tr2.step(Xs[i], ys[i])) for i in range(2, 5)]
    assert got == ref
    assert torch.equal(tr2.flat_w.view(torch.int32), w_ref.view(torch.int32))
    assert pm2.ranges() == r_ref
    sd = tr2.state_dict()
    assert sd['step'] == 5 and any(k.endswith('X_mean_running') for k in sd['model']) and any(k.endswith('.range') for k in sd['model'])
    pm3, tr3 = _trainer(seed=5)
    with pytest.raises(Exception):
        tr3.load(str(tmp_path / 'ck.pt'))                                        # another noise seed: refuse


def test_evaluate_changes_no_state():
    rng = np.random.default_rng(3)
    pm, tr = _trainer()
    X = torch.from_numpy((rng.standard_normal((48, 3, 32, 32)) * 0.5).astype(np.float32)).cuda().contiguous(memory_format=torch.channels_last)
    y = torch.from_numpy(rng.integers(0, 10, 48)).cuda()
    tr.step(X[:16], y[:16])
    before = {k: v.clone() for k, v in pm.state_dict().items()}
    step0 = int(pm.runtime.dev_step)
    loss, acc = tr.evaluate(X, y, batch_size=16)
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    after = pm.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before), [k for k in before if not torch.equal(before[k], after[k])]
    assert int(pm.runtime.dev_step) == step0
    l2, a2 = tr.evaluate(X, y, batch_size=16)
    assert (l2, a2) == (loss, acc)                                               # same noise stream position: reproducible
    tr.step(X[16:32], y[16:32])                                                  # and training carries on


def test_fit_epoch_loop_lr_schedule():
    rng = np.random.default_rng(4)
    X, y = _dataset(rng, 96, 32, 32, 3)
    pipe = PD.Pipeline(torch.from_numpy(X), torch.from_numpy(y), 'cuda', seed=2)
    pm, tr = _trainer()
    test = PD.Pipeline(torch.from_numpy(X[:32]), torch.from_numpy(y[:32]), 'cuda', mean=pipe.mean, augment=False).all()
    logs = []
    hist = tr.fit(pipe, 3, 32, lr_decay_factor=0.1, decay_epochs=(1, 2), test=test, log=logs.append)
    assert len(hist) == 3 and all(np.isfinite(h[1]) and 0 <= h[3] <= 1 for h in hist)
    assert abs(tr.lr - 1e-4) < 1e-12 and abs(float(tr.dev_lr) - 1e-4) < 1e-10
    assert int(pm.runtime.dev_step) == 9 and len(logs) == 3


def test_host_feeder_matches_plain_loop():
    """HostFeeder (prefetched H2D, lagged D2H loss) runs the same steps as the synchronous feed loop."""
    from lbt_b200.trainer import HostFeeder
    rng = np.random.default_rng(6)
    hX = [torch.from_numpy((rng.standard_normal((16, 32, 32, 3)) * 0.5).astype(np.float32)).pin_memory() for _ in range(5)]
    hy = [torch.from_numpy(rng.integers(0, 10, 16)).pin_memory() for _ in range(5)]
    pm, tr = _trainer()
    want = [float(tr.step(x.cuda().permute(0, 3, 1, 2), y.cuda())) for x, y in zip(hX, hy)]
    pm2, tr2 = _trainer()
    Xs = torch.empty(16, 32, 32, 3, device='cuda')
    ys = torch.empty(16, dtype=torch.int64, device='cuda')
    feeder = HostFeeder(lambda: tr2.step(Xs.permute(0, 3, 1, 2), ys), Xs, ys)
    got = []
    feeder.prefetch(hX[0], hy[0])
    for i in range(5):
        if i + 1 < 5:
            feeder.prefetch(hX[i + 1], hy[i + 1])
        got.append(feeder.step())
    got.append(feeder.drain())
    assert got[0] is None and got[1:] == want
    assert torch.equal(tr2.flat_w.view(torch.int32), tr.flat_w.view(torch.int32))

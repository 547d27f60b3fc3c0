"""Parity of the CUDA quantiser (through the C ABI) with the CPU oracle — bit-exact."""
import zlib

import numpy as np
import pytest
import torch

from oracle import dfxp as O
from oracle import philox as P

pytestmark = pytest.mark.gpu

try:
    from lbt_b200 import quantizer as Q
except Exception:  # pragma: no cover
    Q = None


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda() if dtype is None else t.to(dtype).cuda()


def ibt(v):
    return torch.tensor(v, dtype=torch.int32, device='cuda')


def same_bits(a, b):
    """fp32 equality including the sign of zero."""
    return np.array_equal(np.asarray(a, dtype=np.float32).view(np.uint32), np.asarray(b, dtype=np.float32).view(np.uint32))


def test_kat_nearest(kat):
    for row in kat['nearest']:
        bits = row.get('bits', 8)
        ib = ibt(2)
        q, m = Q.quantize(dev(np.array([row['x']], dtype=np.float32)), bits, ib, mant_kind=Q.MANT_S16)
        assert float(q[0]) == np.float32(row['q']), row
        assert int(m[0]) == row['k'], row
        if row.get('neg_zero'):
            assert np.signbit(q.cpu().numpy()[0])


def test_kat_stochastic(kat):
    for row in kat['stochastic']:
        ib = ibt(2)
        x = dev(np.array([[row['x']]], dtype=np.float32))
        u = dev(np.array([row['u']], dtype=np.float32))
        q, m = Q.quantize(x, 8, ib, mode=Q.ROUND_NOISE, noise=u, mant_kind=Q.MANT_S8)
        assert float(q[0, 0]) == np.float32(row['q']), row
        assert int(m[0, 0]) == row['k'], row


def test_kat_controller(kat):
    for row in kat['controller']:
        ib = ibt(row['ib'])
        Q.update_range(dev(np.array(row['x'], dtype=np.float32)), 0.0, 8, ib)
        assert int(ib) == row['new_ib'], row


SHAPES = [(1, 1), (3, 7), (256, 1024), (128, 3072), (5, 4), (7, 12), (1, 4096), (300, 260), (64, 10), (17,),
          (3, 3, 16, 16), (8, 32, 32, 16), (2, 1000003)]


@pytest.mark.parametrize('shape', SHAPES)
@pytest.mark.parametrize('bits', [4, 6, 8, 9, 16])
@pytest.mark.parametrize('mode', ['nearest', 'noise', 'philox'])
def test_parity_with_oracle(shape, bits, mode):
    rng = np.random.default_rng(zlib.crc32(repr((shape, bits, mode)).encode()))
    x = (rng.standard_normal(shape) * 2.0).astype(np.float32)
    x.reshape(-1)[:: max(1, x.size // 7)] = np.float32(2.0 ** 2)          # exact range hits
    inner = shape[1:]
    ib0 = 2
    if mode == 'nearest':
        q_ref, k_ref = O.quantize_nearest(x, bits, ib0)
        kw = dict(mode=Q.ROUND_NEAREST)
    elif mode == 'noise':
        u = rng.random(inner).astype(np.float32)
        q_ref, k_ref = O.quantize_stochastic(x, bits, ib0, u)
        kw = dict(mode=Q.ROUND_NOISE, noise=dev(u))
    else:
        n_inner = int(np.prod(inner)) if len(inner) else 1
        u = P.noise(n_inner, 1234, P.make_offset(5, 9)).reshape(inner)
        q_ref, k_ref = O.quantize_stochastic(x, bits, ib0, u)
        kw = dict(mode=Q.ROUND_PHILOX, seed=1234, offset=Q.make_offset(5, 9))
    n1, n2 = O.overflow_counts(x, bits, ib0)
    r = O.Range(ib0)
    O.update_range(x, 0.0, bits, r)

    ib = ibt(ib0)
    cnt = Q.new_counters('cuda')
    kind = Q.MANT_S8 if bits <= 8 else Q.MANT_S16
    q, m = Q.quantize(dev(x), bits, ib, mant_kind=kind, counters=cnt, update_range=False, **kw)
    assert same_bits(q.cpu().numpy(), q_ref), 'fake-quant output differs'
    np.testing.assert_array_equal(m.cpu().numpy().astype(np.float32), k_ref)
    c = cnt.cpu().numpy()
    assert (int(c[0]), int(c[1]), int(c[2]), int(c[3])) == (n1, n2, x.size, 0)
    assert int(ib) == ib0                                                   # update_range=False leaves the range
    # now the fused controller
    cnt.zero_()
    Q.quantize(dev(x), bits, ib, mant_kind=kind, counters=cnt, update_range=True, **kw)
    assert int(ib) == r.value
    assert cnt.cpu().numpy().tolist() == [0, 0, 0, 0]


@pytest.mark.parametrize('bits', [1, 2, 3, 17, 24, 25, 26, 28, 31])
@pytest.mark.parametrize('mode', ['nearest', 'noise'])
def test_every_width_the_reference_accepts(bits, mode):
    """dfxp:21 asserts 1 <= bits <= 32 and 32 is a pass-through: every width below it runs in lbt_quantize (fp32 output;
    packed mantissas stop at 16 bits).  Includes the widths whose clip bound L-1 is not an fp32 number (bits > 25): it
    rounds to L in the reference's fp32 constant, in the oracle and here alike."""
    rng = np.random.default_rng(bits * 7 + len(mode))
    for ib0 in (2, -3, min(bits - 1, 5)):
        scale = 2.0 ** ib0
        x = (rng.standard_normal((37, 515)) * scale * 0.7).astype(np.float32)
        x.reshape(-1)[::97] = np.float32(scale)
        x.reshape(-1)[1::97] = np.float32(-scale)
        x.reshape(-1)[2::97] = np.float32(scale * (1 - 2.0 ** -20))
        if mode == 'nearest':
            q_ref, _ = O.quantize_nearest(x, bits, ib0)
            kw = dict(mode=Q.ROUND_NEAREST)
        else:
            u = rng.random(x.shape[1:]).astype(np.float32)
            q_ref, _ = O.quantize_stochastic(x, bits, ib0, u)
            kw = dict(mode=Q.ROUND_NOISE, noise=dev(u))
        n1, n2 = O.overflow_counts(x, bits, ib0)
        r = O.Range(ib0)
        O.update_range(x, 0.0, bits, r)
        ib = ibt(ib0)
        cnt = Q.new_counters('cuda')
        q, _ = Q.quantize(dev(x), bits, ib, counters=cnt, update_range=False, **kw)
        assert same_bits(q.cpu().numpy(), q_ref), (bits, ib0)
        c = cnt.cpu().numpy()
        assert (int(c[0]), int(c[1]), int(c[2])) == (n1, n2, x.size)
        cnt.zero_()
        Q.quantize(dev(x), bits, ib, want_fp32=False, counters=cnt, update_range=True, **kw)
        assert int(ib) == r.value


def test_philox_noise_fill_matches_oracle():
    for n in [1, 3, 4, 5, 1000, 4099]:
        u = Q.noise_fill(n, seed=0xDEADBEEFCAFE, offset=Q.make_offset(77, 123456))
        np.testing.assert_array_equal(u.cpu().numpy(), P.noise(n, 0xDEADBEEFCAFE, P.make_offset(77, 123456)))


def test_dev_step_advances_noise():
    from lbt_b200 import _lib
    step = torch.zeros(1, dtype=torch.int64, device='cuda')
    x = dev(np.linspace(-3, 3, 4096, dtype=np.float32).reshape(4, 1024))
    outs = []
    for s in range(3):
        q, _ = Q.quantize(x, 8, ibt(2), mode=Q.ROUND_PHILOX, seed=5, offset=Q.make_offset(11, 0), dev_step=step,
                          update_range=False)
        u = P.noise(1024, 5, P.make_offset(11, s))
        q_ref, _ = O.quantize_stochastic(x.cpu().numpy(), 8, 2, u)
        assert same_bits(q.cpu().numpy(), q_ref)
        outs.append(q)
        _lib.check(_lib.lib().lbt_step_advance(_lib.ptr(step), _lib.stream()))
    assert int(step) == 3 and not torch.equal(outs[0], outs[1])


def test_u8_mantissa_for_nonnegative_9bit():
    rng = np.random.default_rng(3)
    x = np.abs(rng.standard_normal((64, 2048)).astype(np.float32)) * 1.5
    u = rng.random(2048).astype(np.float32)
    _, k_ref = O.quantize_stochastic(x, 9, 2, u)
    assert k_ref.max() == 255 and k_ref.min() >= 0
    _, m = Q.quantize(dev(x), 9, ibt(2), mode=Q.ROUND_NOISE, noise=dev(u), want_fp32=False, mant_kind=Q.MANT_U8)
    assert m.dtype == torch.uint8
    np.testing.assert_array_equal(m.cpu().numpy().astype(np.float32), k_ref)


def test_in_place_and_channels_last():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((8, 16, 12, 12)).astype(np.float32)             # logical NCHW
    xt = dev(x).contiguous(memory_format=torch.channels_last)
    nhwc = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    u = rng.random(nhwc.shape[1:]).astype(np.float32)                        # noise is [H, W, C] (TF layout)
    q_ref, _ = O.quantize_stochastic(nhwc, 8, 2, u)
    q, _ = Q.quantize(xt, 8, ibt(2), mode=Q.ROUND_NOISE, noise=dev(u), out=xt)   # in place
    assert q.data_ptr() == xt.data_ptr()
    assert same_bits(q.permute(0, 2, 3, 1).contiguous().cpu().numpy(), q_ref)


def test_controller_sequence_matches_oracle():
    """Ranges after N controller steps: a drifting tensor walks the range up and down."""
    rng = np.random.default_rng(5)
    r = O.Range(2)
    ib = ibt(2)
    base = rng.standard_normal((32, 512)).astype(np.float32)
    for step in range(24):
        scale = np.float32(2.0 ** (6 * np.sin(step / 3.0)))
        x = base * scale
        q_ref = O.weight_quantization(x, 0.0, 8, r)
        q, _ = Q.quantize(dev(x), 8, ib)
        assert same_bits(q.cpu().numpy(), q_ref), step
        assert int(ib) == r.value, step


def test_update_ranges_multi_and_counter_accumulation():
    """update_range=0 accumulates across launches (data-parallel shards); lbt_update_ranges applies them."""
    from lbt_b200 import _lib
    rng = np.random.default_rng(6)
    n = 5
    ranges = torch.full((n,), 2, dtype=torch.int32, device='cuda')
    counters = torch.zeros(n, 4, dtype=torch.int64, device='cuda')
    bits = torch.tensor([8, 8, 4, 16, 8], dtype=torch.int32, device='cuda')
    target = torch.zeros(n, dtype=torch.float32, device='cuda')
    expect = []
    for i in range(n):
        if i == 4:
            expect.append(2)            # never runs -> untouched
            continue
        shards = [(rng.standard_normal((16, 64)) * (0.3 if i == 1 else 2.5)).astype(np.float32) for _ in range(2)]
        for s in shards:                # two "ranks" accumulate into the same counters
            Q.quantize(dev(s), int(bits[i]), ranges[i:i + 1].view(()), counters=counters[i], update_range=False)
        r = O.Range(2)
        O.update_range(np.concatenate(shards), 0.0, int(bits[i]), r)
        expect.append(r.value)
    assert counters[:, 2].cpu().tolist() == [2048, 2048, 2048, 2048, 0]
    _lib.check(_lib.lib().lbt_update_ranges(_lib.ptr(ranges), _lib.ptr(counters), _lib.ptr(bits), _lib.ptr(target), n,
                                            _lib.stream()))
    assert ranges.cpu().tolist() == expect
    assert int(counters.abs().sum()) == 0


def test_target_overflow_rate_nonzero():
    x = np.zeros((10, 100), dtype=np.float32)
    x[0, :10] = 5.0                     # 1% overflow at ib=2
    for t, exp in [(0.0, 3), (0.02, 1), (0.005, 3)]:
        r = O.Range(2)
        O.update_range(x, t, 8, r)
        ib = ibt(2)
        Q.update_range(dev(x), t, 8, ib)
        assert int(ib) == r.value


def test_api_parity_weight_quantization_and_ste():
    x = dev(np.array([[0.1, -5.0], [4.0, 0.3]], dtype=np.float32)).requires_grad_(True)
    ib = ibt(2)
    assert Q.weight_quantization(x, 0, 32, ib) is x and int(ib) == 2
    with pytest.raises(AssertionError):
        Q.weight_quantization(x, 0, 0, ib)
    q = Q.weight_quantization(x, 0, 8, ib)
    assert int(ib) == 3
    assert q.detach().cpu().tolist() == [[0.09375, -4.0], [3.96875, 0.3125]]
    q.backward(torch.tensor([[1.0, 2.0], [3.0, 4.0]], device='cuda'))
    assert x.grad.cpu().tolist() == [[1.0, 2.0], [3.0, 4.0]]                 # identity, clipped elements included
    r1, r2 = Q.overflow_rate(x.detach(), 8, ibt(2))
    o1, o2 = O.overflow_rate(x.detach().cpu().numpy(), 8, 2)
    assert float(r1) == o1 and float(r2) == o2


def test_large_tensor_properties():
    """BASELINE-size tensor (2^28 elems = the config-3 maximum): idempotence, bounds, counter totals."""
    n_outer, n_inner = 256, 1 << 20
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(n_outer, n_inner, device='cuda', generator=g)
    ib = ibt(2)
    cnt = Q.new_counters('cuda')
    q, m = Q.quantize(x, 8, ib, mode=Q.ROUND_PHILOX, seed=1, offset=3, mant_kind=Q.MANT_S8, counters=cnt,
                      update_range=False)
    assert int(cnt[2]) == n_outer * n_inner
    assert int(cnt[0]) == int((x >= 4).sum() + (x < -4).sum())
    assert int(cnt[1]) == int((x >= 2).sum() + (x < -2).sum())
    assert float(q.max()) <= 127 / 32 and float(q.min()) >= -4.0
    assert torch.equal(q, m.to(torch.float32) / 32)                          # mantissa <-> fake-quant
    unclipped = (x < 127 / 32) & (x >= -4)
    assert float((q - x)[unclipped].abs().max()) < 1 / 32 + 1e-6
    q2, _ = Q.quantize(q, 8, ibt(2))                                         # nearest on a grid point: idempotent
    assert torch.equal(q2, q)
    # stochastic rounding is unbiased where unclipped
    inr = (x.abs() < 3.5)
    assert abs(float((q - x)[inr].mean())) < 2e-4
    # row 0 and row 1 share the noise: same fractional decision for equal inputs
    x[1] = x[0]
    q3, _ = Q.quantize(x[:2], 8, ibt(2), mode=Q.ROUND_PHILOX, seed=1, offset=3)
    assert torch.equal(q3[0], q3[1])


def test_fast_division_is_correctly_rounded():
    """fdiv_by (one reciprocal per denominator + two FMA corrections, used by the batch-norm kernels) must equal the IEEE
    division bit for bit: 2^32 device-generated pairs covering the batch-norm operand ranges, zero mismatches."""
    from lbt_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device='cuda')
    first = torch.zeros(2, dtype=torch.float32, device='cuda')
    for seed in range(4):
        _lib.check(_lib.lib().lbt_test_fdiv(1 << 30, seed, bad.data_ptr(), first.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(bad) == 0, (int(bad), first.tolist())


@pytest.mark.parametrize('shape', [(5, 8, 12, 3), (3, 7, 5, 3), (64, 32, 32, 3)])
@pytest.mark.parametrize('mode', [Q.ROUND_NEAREST, Q.ROUND_PHILOX])
def test_image_mantissas_s9c3_equal_split_of_s16(shape, mode):
    """LBT_MANT_S9C3 (the 16-byte pixels {hi x3, hi x3, lo x3, 0 x7} a first Conv2d_q consumes, k = 2*hi + lo) against the
    plain s16 mantissas of the same call: both the 4-pixels-per-thread kernel (n_inner % 12 == 0) and the per-pixel one
    (n_inner % 12 != 0), with the same counters."""
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(*shape, generator=g) * 1.7).cuda()
    ib = torch.tensor(1, dtype=torch.int32, device='cuda')
    c1, c2 = Q.new_counters('cuda'), Q.new_counters('cuda')
    kw = dict(mode=mode, seed=9, offset=Q.make_offset(3, 1), want_fp32=False, update_range=False)
    _, k16 = Q.quantize(x, 9, ib, mant_kind=Q.MANT_S16, counters=c1, **kw)
    out = torch.empty(*shape[:-1], 16, dtype=torch.int8, device='cuda')
    Q.quantize(x, 9, ib, mant_kind=Q.MANT_S9C3, counters=c2, out_mant=out, **kw)
    k = k16.to(torch.int32)
    hi, lo = k >> 1, k & 1
    want = torch.cat([hi, hi, lo, torch.zeros(*shape[:-1], 7, dtype=torch.int32, device='cuda')], dim=-1).to(torch.int8)
    assert torch.equal(out, want)
    assert torch.equal(c1, c2)

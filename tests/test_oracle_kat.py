"""Pin the CPU oracle to the hand-derived known-answer vectors."""
import numpy as np
import pytest

from oracle import dfxp as O


def test_nearest_kat(kat):
    for row in kat['nearest']:
        bits = row.get('bits', 8)
        q, k = O.quantize_nearest(np.array([row['x']], dtype=np.float32), bits, 2)
        assert k[0] == row['k'], row
        assert q[0] == np.float32(row['q']), row
        if row.get('neg_zero'):
            assert np.signbit(q[0]), 'fake-quant of -0.5 ulp must keep the sign bit (-0.0)'


def test_stochastic_kat(kat):
    for row in kat['stochastic']:
        x = np.array([[row['x']]], dtype=np.float32)          # [dim0=1, inner=1]
        u = np.array([row['u']], dtype=np.float32)
        q, k = O.quantize_stochastic(x, 8, 2, u)
        assert k[0, 0] == row['k'], row
        assert q[0, 0] == np.float32(row['q']), row


def test_controller_kat(kat):
    for row in kat['controller']:
        r = O.Range(row['ib'])
        O.update_range(np.array(row['x'], dtype=np.float32), 0.0, 8, r)
        assert r.value == row['new_ib'], row


def test_multiplier_limit_kat(kat):
    for row in kat['multiplier_limit']:
        m, L = O._multiplier_limit(row['bits'], row['ib'])
        assert (m, L) == (row['m'], row['L'])


def test_weight_quantization_dispatch():
    x = np.array([[0.1, -5.0], [4.0, 0.3]], dtype=np.float32)
    r = O.Range(2)
    assert O.weight_quantization(x, 0, 32, r) is not None and r.value == 2      # bypass, no update
    with pytest.raises(AssertionError):
        O.weight_quantization(x, 0, 0, r)
    with pytest.raises(AssertionError):
        O.weight_quantization(x, 0, 33, r)
    q = O.weight_quantization(x, 0, 8, r)                                        # nearest by default
    assert r.value == 3                                                          # 4.0 and -5.0 overflow
    np.testing.assert_array_equal(q, np.array([[0.09375, -4.0], [3.96875, 0.3125]], dtype=np.float32))
    # the next call uses the updated range (read-then-update)
    q2 = O.weight_quantization(x, 0, 8, r, stochastic=True, noise=np.zeros(2, dtype=np.float32))
    assert q2[1, 0] == 4.0 and r.value == 3                                      # -5.0 still < -8? no: keeps 3


def test_noise_broadcast_over_dim0():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 3, 5)).astype(np.float32)
    u = rng.random((3, 5)).astype(np.float32)
    q, k = O.quantize_stochastic(x, 8, 2, u)
    for i in range(4):
        qi, _ = O.quantize_stochastic(x[i:i + 1], 8, 2, u)
        np.testing.assert_array_equal(q[i], qi[0])


def test_overflow_rate_uses_raw_input():
    x = np.array([3.99, 4.0, -4.0, -4.0001, 2.0, -2.0], dtype=np.float32)
    n1, n2 = O.overflow_counts(x, 8, 2)
    assert (n1, n2) == (2, 5)            # >=L: 4.0 ; <-L: -4.0001 ; half: 3.99,4.0,2.0 ; -4.0,-4.0001
    r1, r2 = O.overflow_rate(x, 8, 2)
    assert r1 == np.float32(2) / np.float32(6) and r2 == np.float32(5) / np.float32(6)


def test_gradient_buffer_known_answer():
    """GradientBuffer_q (dfxp:473-509) by hand: 8 bits, noise 0.5 everywhere; the range starts at 2 and, with no element
    reaching half of it, drops by one after every call (dfxp:84-94), so the step is 2^-5, 2^-6, 2^-7.
    call 1: total 0.01   -> floor(0.01 * 32 + 0.5) = 0      -> q = 0,     buffer = 0.01
    call 2: total 0.02   -> floor(0.02 * 64 + 0.5) = 1      -> q = 1/64,  buffer = 0.02 - 1/64 (carried)
    call 3: a batch of 1 into the 2-row buffer: row 0 = 0.01 + buffer -> floor(1.84 + 0.5) = 2 -> q = 2/128, residual < 0;
            row 1 is zero padding + its buffer -> floor(0.56 + 0.5) = 1; only row 0 is returned (dfxp:506)."""
    import torch
    from oracle import dfxp as O
    f = np.float32
    ctx = O.Context(noise=lambda qid, shape: np.full(shape, 0.5, dtype=np.float32))
    gb = O.GradientBuffer_q(ctx, 8, (2, 1))
    g = torch.full((2, 1), 0.01)
    assert gb.backward(g).tolist() == [[0.0], [0.0]]
    assert gb.buffer.tolist() == [[f(0.01)], [f(0.01)]] and gb.qG.range.value == 1
    out = gb.backward(g)
    assert out.tolist() == [[0.015625], [0.015625]] and gb.qG.range.value == 0
    b2 = f(f(0.01) + f(0.01)) - f(0.015625)
    assert gb.buffer[0, 0].item() == b2 and b2 > 0
    out = gb.backward(torch.full((1, 1), 0.01))
    assert out.tolist() == [[0.015625]] and gb.buffer.shape == (2, 1)
    assert gb.buffer[0, 0].item() == f(f(0.01) + b2) - f(0.015625) < 0
    assert gb.buffer[1, 0].item() == b2 - f(0.0078125)

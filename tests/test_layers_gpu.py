"""Layer-level parity of lbt_b200.dfxp with the CPU oracle on the same inputs and the same Philox noise.

Quantiser outputs are compared bit-exactly (through the exact-GEMM identity); GEMM/conv outputs must
equal RN_fp32(exact integer dot * 2^-f) bit for bit (computed here in fp64 from the oracle's fake-quant
operands) and sit within the fp32-accumulation tolerance of the oracle's own fp32 conv/matmul
(|y - y_ref| <= K * 2^-24 * sum|a||b|).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dfxp as O

pytestmark = pytest.mark.gpu

from lbt_b200 import dfxp as D  # noqa: E402

SEED = 77


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).cuda()


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().cpu()


def conv64(xq, wq, strides, padding):
    return O.tf_conv2d(xq.double(), wq.double(), strides, padding)


CONV_CASES = [
    # Cin, Cout, k, stride, H, batch, bias, signed input
    (16, 16, 3, 1, 8, 4, False, False),
    (16, 32, 3, 2, 8, 4, False, False),
    (16, 32, 1, 2, 8, 4, False, False),
    (32, 64, 1, 1, 6, 3, False, False),
    (3, 16, 3, 1, 10, 4, False, True),
    (3, 64, 5, 1, 9, 2, True, True),
    (64, 128, 5, 1, 7, 2, True, False),
    (3, 64, 7, 2, 17, 2, False, True),
    (24, 40, 3, 2, 7, 3, True, False),       # Cin not a multiple of 16 -> scalar gather path
    (16, 16, 3, 1, 5, 2, False, True),       # signed 9-bit on a 16-channel input (hi|hi|lo split)
    # implicit-GEMM shared-memory layouts: 16B interleaved / 32B / 64B / 128B swizzle, 1 and 2 channel chunks
    (16, 16, 5, 1, 9, 2, False, False),      # 25 taps (odd -> zero pad step)
    (16, 48, 3, 1, 33, 5, True, False),      # M = 5445 (ragged last tile), N = 48
    (32, 32, 3, 1, 8, 4, False, False),
    (32, 64, 3, 2, 9, 3, False, False),
    (64, 64, 3, 1, 6, 2, False, False),
    (64, 128, 3, 2, 9, 2, False, False),
    (128, 128, 3, 1, 5, 2, False, False),
    (256, 128, 3, 1, 4, 2, False, False),
    (128, 256, 3, 2, 7, 3, False, False),
    (256, 512, 3, 1, 4, 3, False, False),    # two N tiles
]


@pytest.mark.parametrize('Cin,Cout,k,s,H,B,bias,signed', CONV_CASES)
def test_conv2d_q_forward_backward(Cin, Cout, k, s, H, B, bias, signed):
    _conv_case(Cin, Cout, k, s, H, B, bias, signed, None)


@pytest.mark.parametrize('Cin,Cout,k,s,H,B,bias,signed', [
    (16, 32, 3, 1, 8, 4, False, False), (32, 64, 3, 2, 9, 3, True, False), (64, 64, 1, 1, 6, 3, False, False),
    (3, 16, 3, 1, 10, 4, False, True), (128, 256, 3, 2, 7, 3, False, False), (24, 40, 3, 2, 7, 3, True, False),
    (16, 32, 1, 2, 8, 4, False, False)])
@pytest.mark.parametrize('gbits', [16, 12])
def test_conv2d_q_wide_gradients(Cin, Cout, k, s, H, B, bias, signed, gbits):
    """BASELINE config 5 (8-bit W/A, 16-bit G): gradient mantissas split k = 256*hi + lo, still exactly rounded."""
    _conv_case(Cin, Cout, k, s, H, B, bias, signed, gbits)


def _conv_case(Cin, Cout, k, s, H, B, bias, signed, gbits):
    rng = np.random.default_rng(Cin * 100 + Cout + k + s)
    ctx = O.Context(O.PhiloxNoise(SEED))
    wd = 2e-4
    ol = O.Conv2d_q(ctx, 'c', 8, [k, k, Cin, Cout], [1, s, s, 1], 'SAME', use_bias=bias, weight_decay=wd, rng=rng,
                    grad_bits=gbits)
    rt = D.Runtime(SEED)
    pl = D.Conv2d_q(8, Cin, Cout, k, s, 'SAME', bias=bias, weight_decay=wd, input_signed=signed, runtime=rt,
                    grad_bits=gbits).cuda()
    rt.finalize('cuda')
    pl.weight.data.copy_(ol.W.detach())
    x = torch.from_numpy(rng.standard_normal((B, H, H, Cin)).astype(np.float32) * 1.5)
    if not signed:
        x = x.abs()
    if bias:
        b0 = torch.from_numpy(rng.standard_normal(Cout).astype(np.float32) * 0.1)
        ol.b.data.copy_(b0)
        pl.bias.data.copy_(b0)

    y_o = ol.forward(x)
    xg = nchw(x).requires_grad_(True)
    y_p = pl(xg)
    # exact identity: y == RN(conv64(Xq, Wq)) (+ bq)
    y64 = conv64(ol.Xq.detach(), ol.Wq.detach(), ol.strides, 'SAME').float()
    if bias:
        y64 = y64 + ol.bq.detach()
    assert torch.equal(nhwc(y_p.detach()), y64), 'fprop is not the exactly-rounded integer result'
    tol = (k * k * Cin) * 2.0 ** -24 * float(O.tf_conv2d(ol.Xq.detach().abs(), ol.Wq.detach().abs(), ol.strides, 'SAME').max())
    assert float((nhwc(y_p.detach()) - y_o).abs().max()) <= tol + 1e-12

    g = torch.from_numpy(rng.standard_normal(tuple(y_o.shape)).astype(np.float32) * 0.7)
    dx_o = ol.backward(g)
    y_p.backward(nchw(g))
    gq = ol.gradq
    # dgrad / wgrad exact references in fp64 through autograd on the fp64 conv
    X64 = ol.Xq.detach().double().requires_grad_(True)
    W64 = ol.Wq.detach().double().requires_grad_(True)
    y2 = O.tf_conv2d(X64, W64, ol.strides, 'SAME')
    dX64, dW64 = torch.autograd.grad(y2, [X64, W64], gq.double())
    assert torch.equal(nhwc(xg.grad), dX64.float()), 'dgrad is not the exactly-rounded integer result'
    dW_ref = dW64.float() + (2 * wd) * ol.W.detach()
    assert torch.equal(pl.weight.grad.cpu(), dW_ref), 'wgrad (+2*wd*W) is not the exactly-rounded result'
    assert torch.allclose(pl.weight.grad.cpu(), ol.dW, rtol=1e-4, atol=1e-4 * float(ol.dW.abs().max()))
    assert torch.allclose(nhwc(xg.grad), dx_o, rtol=1e-4, atol=1e-5 * float(dx_o.abs().max()) + 1e-12)
    if bias:
        db64 = gq.double().sum(dim=(0, 1, 2)).float()
        assert torch.equal(pl.bias.grad.cpu(), db64)
    # controller: apply and compare every range
    rt.update_ranges()
    want = [int(q.range) for q in ol.quantizers()]
    assert list(rt.ranges().values()) == want


@pytest.mark.parametrize('B,In,Out,bias', [(128, 2048, 400, True), (128, 400, 10, True), (256, 64, 10, False), (7, 20, 5, True)])
@pytest.mark.parametrize('gbits', [None, 16])
def test_linear_q_forward_backward(B, In, Out, bias, gbits):
    rng = np.random.default_rng(B + In + Out)
    ctx = O.Context(O.PhiloxNoise(SEED))
    wd = 2e-4
    ol = O.Dense_q(ctx, 'd', 8, In, Out, use_bias=bias, weight_decay=wd, rng=rng, grad_bits=gbits)
    rt = D.Runtime(SEED)
    pl = D.Linear_q(8, In, Out, bias=bias, weight_decay=wd, runtime=rt, grad_bits=gbits).cuda()
    rt.finalize('cuda')
    pl.weight.data.copy_(ol.W.detach())
    x = torch.from_numpy(rng.standard_normal((B, In)).astype(np.float32))
    y_o = ol.forward(x)
    xg = x.cuda().requires_grad_(True)
    y_p = pl(xg)
    y64 = (ol.Xq.detach().double() @ ol.Wq.detach().double()).float()
    if bias:
        y64 = y64 + ol.bq.detach()
    assert torch.equal(y_p.detach().cpu(), y64)
    assert torch.allclose(y_p.detach().cpu(), y_o, rtol=1e-5, atol=1e-5)
    g = torch.from_numpy(rng.standard_normal((B, Out)).astype(np.float32) * 0.3)
    dx_o = ol.backward(g)
    y_p.backward(g.cuda())
    gq = ol.gradq.double()
    assert torch.equal(xg.grad.cpu(), (gq @ ol.Wq.detach().double().T).float())
    assert torch.equal(pl.weight.grad.cpu(), (ol.Xq.detach().double().T @ gq).float() + (2 * wd) * ol.W.detach())
    if bias:
        assert torch.equal(pl.bias.grad.cpu(), gq.sum(0).float())
    assert torch.allclose(xg.grad.cpu(), dx_o, rtol=1e-4, atol=1e-6)
    rt.update_ranges()
    assert list(rt.ranges().values()) == [int(q.range) for q in ol.quantizers()]


@pytest.mark.parametrize('gbits', [None, 16, 12])
@pytest.mark.parametrize('C,H,B', [(16, 8, 8), (64, 4, 16), (10, 3, 5)])
def test_batchnorm_q_forward_backward(C, H, B, gbits):
    rng = np.random.default_rng(C + H + B)
    ctx = O.Context(O.PhiloxNoise(SEED))
    wd = 2e-4
    ol = O.BatchNorm_q(ctx, 'bn', 8, C, True, weight_decay=wd, grad_bits=gbits)
    rt = D.Runtime(SEED)
    pl = D.BatchNorm2d_q(8, C, weight_decay=wd, runtime=rt, grad_bits=gbits).cuda()
    rt.finalize('cuda')
    g0 = torch.from_numpy((1 + 0.3 * rng.standard_normal(C)).astype(np.float32))
    b0 = torch.from_numpy((0.2 * rng.standard_normal(C)).astype(np.float32))
    ol.layers[1].gamma.data.copy_(g0)
    ol.layers[1].beta.data.copy_(b0)
    pl[1].gamma.data.copy_(g0)
    pl[1].beta.data.copy_(b0)
    x = torch.from_numpy((rng.standard_normal((B, H, H, C)) * 1.2 + 0.3).astype(np.float32))
    y_o = ol.forward(x)
    xg = nchw(x).requires_grad_(True)
    y_p = pl(xg)
    # the rescale quantiser sees a normalised value that may differ in the last ulp between reduction orders:
    # allow a handful of one-step mantissa differences
    step = 2.0 ** -(8 - 2 - 1) * float(g0.abs().max())
    diff = (nhwc(y_p.detach()) - y_o).abs()
    assert float(diff.max()) <= step + 1e-6
    assert float((diff > 1e-6).float().mean()) < 0.01
    assert torch.allclose(pl[0].X_mean_running.cpu(), ol.layers[0].X_mean_running, rtol=1e-5, atol=1e-7)
    assert torch.allclose(pl[0].X_var_running.cpu(), ol.layers[0].X_var_running, rtol=1e-5, atol=1e-7)
    g = torch.from_numpy(rng.standard_normal(tuple(y_o.shape)).astype(np.float32) * 0.5)
    dx_o = ol.backward(g)
    y_p.backward(nchw(g))
    dgam_o, dbeta_o = ol.layers[1].dgamma, ol.layers[1].dbeta
    assert torch.allclose(pl[1].beta.grad.cpu(), dbeta_o, rtol=1e-5, atol=1e-5)
    assert torch.allclose(pl[1].gamma.grad.cpu(), dgam_o, rtol=1e-3, atol=step * B * H * H * 0.02 + 1e-4)
    ddiff = (nhwc(xg.grad) - dx_o).abs()
    assert float((ddiff > 1e-4 * float(dx_o.abs().max())).float().mean()) < 0.02
    rt.update_ranges()
    assert list(rt.ranges().values()) == [int(q.range) for q in ol.quantizers()]


@pytest.mark.parametrize('gbits', [None, 16])
@pytest.mark.parametrize('C,H,B,relu', [(16, 8, 8, False), (64, 4, 16, True), (128, 7, 6, True), (32, 5, 33, False)])
def test_batchnorm_q_bit_exact_vs_exact_oracle(C, H, B, relu, gbits):
    """Normalization_q + Rescale_q (+ the ReLU the blocks fold in) against the oracle in exactly-rounded-accumulation mode
    (fp64 batch sums rounded once = the integer sums of the kernels; every elementwise operation a single fp32 rounding in the
    same order): forward output, running statistics, dX, dgamma, dbeta and all six ranges BIT FOR BIT, for 8- and 16-bit
    gradient quantisers."""
    rng = np.random.default_rng(C * 3 + H + B + (gbits or 0))
    ctx = O.Context(O.PhiloxNoise(SEED), exact=True)
    wd = 2e-4
    ol = O.BatchNorm_q(ctx, 'bn', 8, C, True, weight_decay=wd, grad_bits=gbits)
    orelu = O.ReLU_q()
    rt = D.Runtime(SEED)
    pl = D.BatchNorm2d_q(8, C, weight_decay=wd, runtime=rt, grad_bits=gbits, relu=relu).cuda()
    rt.finalize('cuda')
    g0 = torch.from_numpy((1 + 0.3 * rng.standard_normal(C)).astype(np.float32))
    b0 = torch.from_numpy((0.2 * rng.standard_normal(C)).astype(np.float32))
    for t, v in ((ol.layers[1].gamma, g0), (ol.layers[1].beta, b0), (pl[1].gamma, g0), (pl[1].beta, b0)):
        t.data.copy_(v)
    x = torch.from_numpy((rng.standard_normal((B, H, H, C)) * 1.2 + 0.3).astype(np.float32))
    y_o = ol.forward(x)
    if relu:
        y_o = orelu.forward(y_o)
    xg = nchw(x).requires_grad_(True)
    y_p = pl(xg)
    assert torch.equal(nhwc(y_p.detach()), y_o), int((nhwc(y_p.detach()) != y_o).sum())
    assert torch.equal(pl[0].X_mean_running.cpu(), ol.layers[0].X_mean_running)
    assert torch.equal(pl[0].X_var_running.cpu(), ol.layers[0].X_var_running)
    g = torch.from_numpy((rng.standard_normal(tuple(y_o.shape)) * (0.5 if gbits is None else 0.01)).astype(np.float32))
    dx_o = ol.backward(orelu.backward(g) if relu else g)
    y_p.backward(nchw(g))
    assert torch.equal(pl[1].beta.grad.cpu(), ol.layers[1].dbeta)
    assert torch.equal(pl[1].gamma.grad.cpu(), ol.layers[1].dgamma)
    got = nhwc(xg.grad)
    assert torch.equal(got, dx_o), '%d of %d elements differ, max %g' % (int((got != dx_o).sum()), dx_o.numel(), float((got - dx_o).abs().max()))
    rt.update_ranges()
    assert list(rt.ranges().values()) == [int(q.range) for q in ol.quantizers()]


def test_plumbing_layers_match_tf_semantics():
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((2, 7, 7, 5)).astype(np.float32))
    # max-pool 3x3/2 SAME on 7 -> 4 (pad 1 before, 1 after); on 8 -> 4 (pad 0 before, 1 after)
    for H in (7, 8, 32):
        xx = torch.from_numpy(rng.standard_normal((2, H, H, 5)).astype(np.float32))
        want = O.tf_max_pool(xx, [1, 3, 3, 1], [1, 2, 2, 1], 'SAME')
        got = nhwc(D.MaxPool_q(3, 2, 'SAME')(nchw(xx)))
        assert torch.equal(got, want)
    want = O.tf_avg_pool(x, [1, 7, 7, 1], [1, 1, 1, 1], 'VALID')
    got = nhwc(D.AvgPool_q(7, 1)(nchw(x)))
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)
    # flatten follows NHWC order
    assert torch.equal(D.Flatten_q(7 * 7 * 5)(nchw(x)).cpu(), x.reshape(2, -1))
    # dropout: keep-prob semantics, x / keep * floor(keep + u)
    d = D.Dropout_q(0.5)
    u = torch.rand(2, 5, 7, 7, device='cuda')
    d.uniform_fn = lambda t: u
    xg = nchw(x)
    assert torch.equal(d(xg), xg / 0.5 * torch.floor(0.5 + u))
    d.eval()
    assert d(xg) is xg


def test_gradient_buffer_error_feedback_bit_exact():
    """GradientBuffer_q (dfxp:473-509): quantised gradient, residual buffer, counters and range over several
    steps incl. a short last batch, against the oracle with the same explicit noise."""
    rng = np.random.default_rng(9)
    shape = (8, 4, 4, 16)                                                    # reference layout [batch, H, W, C]
    ctx = O.Context(noise=O.NumpyNoise(3))
    ol = O.GradientBuffer_q(ctx, 8, shape)
    rt = D.Runtime(seed=0)
    pl = D.GradientBuffer_q(8, shape, runtime=rt).cuda()
    noises = {}
    rt.noise_fn = lambda site, n_inner, dev: noises['u'].to(dev)
    rt.finalize('cuda')
    for step, batch in enumerate([8, 8, 5, 8]):
        g = (rng.standard_normal((batch,) + shape[1:]) * (0.05 if step else 3.0)).astype(np.float32)
        want = ol.backward(torch.from_numpy(g))
        noises['u'] = torch.from_numpy(np.ascontiguousarray(ol.qG.last_noise).reshape(-1))
        x = torch.zeros(batch, shape[3], shape[1], shape[2], device='cuda').contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = pl(x)
        y.backward(torch.from_numpy(g).cuda().permute(0, 3, 1, 2))
        got = x.grad.permute(0, 2, 3, 1).contiguous().cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.numpy().view(np.uint32)), step
        assert np.array_equal(pl.buffer.cpu().numpy().view(np.uint32), ol.buffer.numpy().view(np.uint32)), step
        n1, n2, numel = ctx.last_counts[ol.qG.qid]
        c = pl.qG.counters.cpu().tolist()
        assert c[:3] == [n1, n2, numel] and numel == int(np.prod(shape))
        rt.update_ranges()
        assert int(pl.qG.range) == int(ol.qG.range), step


# ---- halo-patch loader of the gather kernel (stride-1 convolutions on large images) vs the im2col-row gather -------------

@pytest.mark.parametrize('N,H,W,Cin,Cout,k,pad', [(3, 32, 32, 16, 16, 3, 1), (2, 56, 56, 32, 64, 3, 1), (2, 16, 16, 32, 32, 3, 1),
                                                  (2, 24, 40, 64, 64, 5, 2), (2, 32, 32, 16, 32, 3, 0), (1, 64, 48, 64, 128, 3, 1),
                                                  (5, 32, 32, 16, 16, 1, 0), (2, 30, 30, 16, 16, 3, 2)])
def test_conv_halo_loader_equals_row_gather(N, H, W, Cin, Cout, k, pad):
    from lbt_b200 import _lib, quantizer as Q
    rng = np.random.default_rng(N * H + Cin + k)
    OH, OW = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x = torch.from_numpy(rng.integers(0, 256, (N, H, W, Cin), dtype=np.uint8)).cuda()
    Kf = k * k * Cin
    w = torch.from_numpy(rng.integers(-128, 128, (Cout, Kf), dtype=np.int8))
    wt = torch.zeros(Cout, D._pitch16(Kf), dtype=torch.int8, device='cuda')[:, :Kf]
    wt.copy_(w)
    ib = torch.tensor(1, dtype=torch.int32, device='cuda')
    bias = torch.randn(Cout, device='cuda')
    addend = torch.randn(N * OH * OW, Cout, device='cuda')
    rt = D.Runtime(seed=5)
    site = D.QuantSite(rt, 'q', 8, 2).cuda()
    rt.finalize('cuda')
    res = {}
    try:
        for halo in (1, 0):
            _lib.lib().lbt_conv_set_halo(11 if halo else 8)     # bit 3: not the TMA halo kernel (its own test is below)
            y = torch.full((N * OH * OW, Cout), float('nan'), device='cuda')
            D._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, 1, 1, pad, pad, OH, OW, ib, ib, -12, bias, y, addend=addend)
            k_out = torch.zeros(N * OH * OW, Cout, dtype=torch.int8, device='cuda')
            sums = torch.zeros(2 * Cout, dtype=torch.int64, device='cuda')
            site.counters.zero_()
            qs = site.abi(OH * OW * Cout, 'cuda')
            D._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, 1, 1, pad, pad, OH, OW, ib, ib, -12, None, None, bnq=(qs, k_out, sums))
            torch.cuda.synchronize()
            res[halo] = (y, k_out, sums, site.counters.clone())
    finally:
        _lib.lib().lbt_conv_set_halo(1)
    assert _lib.lib().lbt_conv_ldg_debug_error() == 0
    for a, b in zip(res[1][:3], res[0][:3]):
        assert torch.equal(a, b)
    ca, cb = res[1][3], res[0][3]     # min/max statistics count THREADS that saw an overflow: only zero / non-zero is defined
    assert torch.equal(ca[:2] > 0, cb[:2] > 0) and torch.equal(ca[2:], cb[2:])
    # and against the exact convolution (fp64): RN_fp32(exact * 2^e) + bias + addend
    xe = x.double().permute(0, 3, 1, 2)
    we = w.double().view(Cout, k, k, Cin).permute(0, 3, 1, 2).cuda()
    ref = F.conv2d(xe, we, padding=pad).permute(0, 2, 3, 1).reshape(N * OH * OW, Cout)
    want = ((ref * 2.0 ** (-12 + 2)).float() + bias) + addend
    assert torch.equal(res[1][0], want)


@pytest.mark.parametrize('N,H,W,Cout,k', [(2, 64, 64, 64, 7), (3, 48, 40, 32, 7), (2, 32, 32, 16, 3), (1, 66, 50, 64, 5), (2, 37, 45, 64, 7),
                                          (1, 224, 224, 64, 7), (2, 34, 34, 128, 4), (1, 40, 72, 48, 8)])
def test_conv_halo_stride2_equals_row_gather(N, H, W, Cout, k):
    """The strided halo-patch loader (7x7/2 ImageNet stem on 16-byte pixels: column-parity planes in shared memory, taps
    (r, s) / (r, s + 2) per instruction) against the im2col-row gather of the same kernel and against the exact convolution,
    TF 'SAME' padding, fp32 and fused re-quantising epilogues."""
    from lbt_b200 import _lib, quantizer as Q
    rng = np.random.default_rng(N * H + W + Cout + k)
    Cin = 16
    OH, pt, _ = D.same_pad(H, k, 2)
    OW, pl, _ = D.same_pad(W, k, 2)
    x = torch.from_numpy(rng.integers(-128, 128, (N, H, W, Cin), dtype=np.int8)).cuda()
    Kf = k * k * Cin
    w = torch.from_numpy(rng.integers(-128, 128, (Cout, Kf), dtype=np.int8))
    wt = torch.zeros(Cout, D._pitch16(Kf), dtype=torch.int8, device='cuda')[:, :Kf]
    wt.copy_(w)
    ib = torch.tensor(1, dtype=torch.int32, device='cuda')
    bias = torch.randn(Cout, device='cuda')
    rt = D.Runtime(seed=5)
    site = D.QuantSite(rt, 'q', 8, 2).cuda()
    rt.finalize('cuda')
    res = {}
    try:
        for halo in (1, 0):
            _lib.lib().lbt_conv_set_halo(7 if halo else 0)      # bit 2: ragged images too (partial 8 x 16 patches)
            y = torch.full((N * OH * OW, Cout), float('nan'), device='cuda')
            D._conv_implicit(x, Q.MANT_S8, wt, Cout, k, k, 2, 2, pt, pl, OH, OW, ib, ib, -14, bias, y)
            k_out = torch.zeros(N * OH * OW, Cout, dtype=torch.int8, device='cuda')
            sums = torch.zeros(2 * Cout, dtype=torch.int64, device='cuda')
            site.counters.zero_()
            qs = site.abi(OH * OW * Cout, 'cuda')
            D._conv_implicit(x, Q.MANT_S8, wt, Cout, k, k, 2, 2, pt, pl, OH, OW, ib, ib, -14, None, None, bnq=(qs, k_out, sums))
            torch.cuda.synchronize()
            res[halo] = (y, k_out, sums, site.counters.clone())
    finally:
        _lib.lib().lbt_conv_set_halo(1)
    assert _lib.lib().lbt_conv_ldg_debug_error() == 0
    for a, b in zip(res[1][:3], res[0][:3]):
        assert torch.equal(a, b)
    ca, cb = res[1][3], res[0][3]     # min/max statistics count THREADS that saw an overflow: only zero / non-zero is defined
    assert torch.equal(ca[:2] > 0, cb[:2] > 0) and torch.equal(ca[2:], cb[2:])
    xe = x.double().permute(0, 3, 1, 2)
    pb, pr = max((OH - 1) * 2 + k - H, 0) - pt, max((OW - 1) * 2 + k - W, 0) - pl
    xe = F.pad(xe, (pl, pr, pt, pb))
    we = w.double().view(Cout, k, k, Cin).permute(0, 3, 1, 2).cuda()
    ref = F.conv2d(xe, we, stride=2).permute(0, 2, 3, 1).reshape(N * OH * OW, Cout)
    want = (ref * 2.0 ** (-14 + 2)).float() + bias
    assert torch.equal(res[1][0], want)


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,pad', [(2, 56, 56, 64, 64, 3, 1), (3, 28, 28, 128, 128, 3, 1), (2, 32, 48, 64, 128, 3, 1),
                                                  (1, 16, 8, 128, 64, 3, 1), (2, 19, 13, 64, 48, 3, 1), (1, 40, 24, 64, 48, 5, 2),
                                                  (2, 17, 31, 64, 64, 3, 0), (1, 33, 9, 128, 128, 2, 0), (4, 14, 14, 128, 128, 3, 1)])
def test_conv_tma_halo_equals_im2col(N, H, W, Cin, Cout, k, pad):
    """conv_halo.cu (one tiled TMA load of the input patch per 8 x 16 output patch, MMA descriptors that start at unaligned
    pixels of the swizzled patch, resident filter bank) against the im2col-mode TMA kernel and the exact convolution: fp32
    epilogue with bias + addend, and the fused re-quantising epilogue (mantissas, exact sums, counters).  Ragged images run
    with the fill-ratio rule off (partial patches, out-of-image rows masked)."""
    from lbt_b200 import _lib, quantizer as Q
    rng = np.random.default_rng(N * H + Cin + k + W)
    OH, OW = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    x = torch.from_numpy(rng.integers(0, 256, (N, H, W, Cin), dtype=np.uint8)).cuda()
    Kf = k * k * Cin
    w = torch.from_numpy(rng.integers(-128, 128, (Cout, Kf), dtype=np.int8))
    wt = torch.zeros(Cout, D._pitch16(Kf), dtype=torch.int8, device='cuda')[:, :Kf]
    wt.copy_(w)
    ib = torch.tensor(1, dtype=torch.int32, device='cuda')
    bias = torch.randn(Cout, device='cuda')
    addend = torch.randn(N * OH * OW, Cout, device='cuda')
    rt = D.Runtime(seed=5)
    site = D.QuantSite(rt, 'q', 8, 2).cuda()
    rt.finalize('cuda')
    res = {}
    launches = {}
    try:
        for halo in (1, 0):
            _lib.lib().lbt_conv_set_halo(5 if halo else 8)
            before = _lib.lib().lbt_conv_halo_launches()
            y = torch.full((N * OH * OW, Cout), float('nan'), device='cuda')
            D._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, 1, 1, pad, pad, OH, OW, ib, ib, -12, bias, y, addend=addend)
            k_out = torch.zeros(N * OH * OW, Cout, dtype=torch.int8, device='cuda')
            sums = torch.zeros(2 * Cout, dtype=torch.int64, device='cuda')
            site.counters.zero_()
            qs = site.abi(OH * OW * Cout, 'cuda')
            D._conv_implicit(x, Q.MANT_U8, wt, Cout, k, k, 1, 1, pad, pad, OH, OW, ib, ib, -12, None, None, bnq=(qs, k_out, sums))
            torch.cuda.synchronize()
            res[halo] = (y, k_out, sums, site.counters.clone())
            launches[halo] = _lib.lib().lbt_conv_halo_launches() - before
    finally:
        _lib.lib().lbt_conv_set_halo(1)
    assert launches == {1: 2, 0: 0}, launches      # the two routes really are two kernels
    assert _lib.lib().lbt_conv_debug_error() == 0
    for a, b in zip(res[1][:3], res[0][:3]):
        assert torch.equal(a, b)
    # min/max statistics: the overflow counters hold "threads that saw an overflow" (the controller only tests > 0), so they
    # depend on the kernel's thread count; the element count and the zero / non-zero decision must agree
    ca, cb = res[1][3], res[0][3]
    assert torch.equal(ca[:2] > 0, cb[:2] > 0) and torch.equal(ca[2:], cb[2:])
    xe = x.double().permute(0, 3, 1, 2)
    we = w.double().view(Cout, k, k, Cin).permute(0, 3, 1, 2).cuda()
    ref = F.conv2d(xe, we, padding=pad).permute(0, 2, 3, 1).reshape(N * OH * OW, Cout)
    want = ((ref * 2.0 ** (-12 + 2)).float() + bias) + addend
    assert torch.equal(res[1][0], want)


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,s', [(2, 56, 56, 64, 128, 3, 2), (3, 28, 28, 128, 256, 3, 2), (2, 14, 14, 256, 512, 3, 2),
                                                (2, 56, 56, 64, 128, 1, 2), (2, 28, 28, 128, 256, 1, 2), (2, 27, 31, 64, 128, 3, 2),
                                                (1, 30, 30, 128, 128, 5, 3), (2, 16, 16, 256, 1024, 1, 2)])
@pytest.mark.parametrize('with_addend', [False, True])
def test_strided_dgrad_parity_classes_equal_im2col_and_exact(N, H, W, Cin, Cout, k, s, with_addend):
    """lbt_conv_i8_dgrad_strided (one stride-1 sub-convolution per parity class of input pixels on the class-ordered filter
    operand, output rows written through the class's sub-lattice of dX) against the transposed im2col matrix + GEMM and
    against the exact transposed convolution (fp64), TF 'SAME' padding, with and without the shortcut branch's addend."""
    from lbt_b200 import quantizer as Q
    conv = D.Conv2d_q(8, Cin, Cout, k, s, 'SAME', bias=False).cuda()
    rng = np.random.default_rng(N * H + Cin + k + s)
    OH, pt, _ = D.same_pad(H, k, s)
    OW, pl, _ = D.same_pad(W, k, s)
    geom = (N, H, W, Cin, Cout, k, k, s, s, pt, pl, OH, OW)
    xm = torch.from_numpy(rng.integers(0, 256, (N, H, W, Cin), dtype=np.uint8)).cuda()
    wm = torch.from_numpy(rng.integers(-128, 128, (k, k, Cin, Cout), dtype=np.int8)).cuda()
    gm = torch.from_numpy(rng.integers(-128, 128, (N, OH, OW, Cout), dtype=np.int8)).cuda()
    addend = torch.randn(N, H, W, Cin, device='cuda') if with_addend else None
    from lbt_b200 import _lib
    calls = []
    orig = _lib.call

    def spy(fn, *a, **kw):
        calls.append(fn)
        return orig(fn, *a, **kw)

    res = {}
    _lib.call = spy
    try:
        for cls in (True, False):
            D.DGRAD_CLASSES = cls
            del calls[:]
            dx, _, _ = D._conv_backward(conv, geom, xm, Q.MANT_U8, wm, None, gm, True, False, False, addend=addend)
            torch.cuda.synchronize()
            res[cls] = dx.clone()
            assert ('lbt_conv_i8_dgrad_strided' in calls) == cls, calls
    finally:
        _lib.call = orig
        D.DGRAD_CLASSES = True
    assert _lib.lib().lbt_conv_debug_error() == 0
    assert torch.equal(res[True], res[False])
    # exact: dX = conv_transpose(g, W) in fp64, scaled by 2^(ib_g + ib_w - 7 - 7)
    ge = gm.double().permute(0, 3, 1, 2)
    we = wm.double().permute(3, 2, 0, 1)                       # [Cout, Cin, kh, kw] = conv_transpose2d's weight layout
    full = F.conv_transpose2d(ge, we, stride=s)                 # [(OH-1)*s + k] grid that starts at input row -pt
    ref = torch.zeros(N, Cin, H, W, dtype=torch.float64, device='cuda')
    hh, ww = min(H, full.shape[2] - pt), min(W, full.shape[3] - pl)
    ref[:, :, :hh, :ww] = full[:, :, pt:pt + hh, pl:pl + ww]
    e = int(conv.qG.range.item()) + int(conv.qW.range.item()) - 14
    want = (ref * 2.0 ** e).float().permute(0, 2, 3, 1)
    if addend is not None:
        want = want + addend
    assert torch.equal(res[True], want.contiguous())


@pytest.mark.parametrize('N,H,W,k', [(2, 32, 32, 7), (3, 30, 34, 7), (2, 64, 48, 7), (2, 18, 22, 3), (1, 20, 16, 5), (2, 224, 224, 7)])
def test_stem_wgrad_c3_equals_general_kernel_and_exact_sum(N, H, W, k):
    """lbt_conv_i8_wgrad_c3 (8-byte pixels, one tiled TMA load per filter row, all accumulator tiles resident) against
    lbt_conv_i8_wgrad on the 16-byte pixels and against the exact integer sum (fp64 conv2d_weight on the CPU for the
    small shapes): dW[r,s,c,co] = sum k[n, 2oh+r-pt, 2ow+s-pl, c] * g[n,oh,ow,co], k = 2*hi + lo.  Ragged patches included."""
    from lbt_b200 import _lib, quantizer as Q
    Cout, s = 64, 2
    OH, pt, _ = D.same_pad(H, k, s)
    OW, pl, _ = D.same_pad(W, k, s)
    gen = torch.Generator().manual_seed(N * 1000 + H + k)
    kx = torch.randint(-256, 256, (N, H, W, 3), generator=gen, dtype=torch.int32)
    hi, lo = kx >> 1, kx & 1
    x16 = torch.cat([hi, hi, lo, torch.zeros(N, H, W, 7, dtype=torch.int32)], dim=-1).to(torch.int8).cuda()
    g = torch.randint(-128, 128, (N * OH * OW, Cout), generator=gen, dtype=torch.int32).to(torch.int8).cuda()
    nb = int(_lib.lib().lbt_stem_pack8_bytes(N, H, OW))
    work = torch.empty(nb, dtype=torch.int8, device='cuda')
    acc8 = torch.zeros(512, Cout, dtype=torch.int64, device='cuda')
    ok = _lib.try_call('lbt_conv_i8_wgrad_c3', _lib.ptr(x16), N, H, W, _lib.ptr(g), Q.MANT_S8, Cout, k, k, pt, pl, OH, OW,
                       _lib.ptr(work), 1, _lib.ptr(acc8), 1, _lib.stream())
    assert ok, 'lbt_conv_i8_wgrad_c3 declined a stem shape'
    # a second, unsigned gradient plane with alpha = 3 into its own sums, against the work buffer of the first call
    gu = torch.randint(0, 256, (N * OH * OW, Cout), generator=gen, dtype=torch.int32).to(torch.uint8).cuda()
    acc8u = torch.zeros(512, Cout, dtype=torch.int64, device='cuda')
    assert _lib.try_call('lbt_conv_i8_wgrad_c3', _lib.ptr(x16), N, H, W, _lib.ptr(gu), Q.MANT_U8, Cout, k, k, pt, pl, OH, OW,
                         _lib.ptr(work), 0, _lib.ptr(acc8u), 3, _lib.stream())
    acc16u = torch.zeros(k * k * 16, Cout, dtype=torch.int64, device='cuda')
    _lib.call('lbt_conv_i8_wgrad', _lib.ptr(x16), Q.MANT_S8, N, H, W, 16, _lib.ptr(gu), Q.MANT_U8, Cout, k, k, s, s, pt, pl, OH, OW,
              _lib.ptr(acc16u), 3, 0, _lib.stream())
    au, bu = acc8u.view(8, 8, 8, Cout)[:k, :k], acc16u.view(k * k, 16, Cout)
    assert torch.equal((2 * au[:, :, 0:3] + au[:, :, 4:7]).reshape(k * k * 3, Cout), (bu[:, 0:3] + bu[:, 3:6] + bu[:, 6:9]).reshape(k * k * 3, Cout))
    torch.cuda.synchronize()
    assert _lib.lib().lbt_conv_debug_error() == 0
    a = acc8.view(8, 8, 8, Cout)[:k, :k]
    got = (2 * a[:, :, 0:3] + a[:, :, 4:7]).reshape(k * k * 3, Cout)
    acc16 = torch.zeros(k * k * 16, Cout, dtype=torch.int64, device='cuda')
    _lib.call('lbt_conv_i8_wgrad', _lib.ptr(x16), Q.MANT_S8, N, H, W, 16, _lib.ptr(g), Q.MANT_S8, Cout, k, k, s, s, pt, pl, OH, OW,
              _lib.ptr(acc16), 1, 0, _lib.stream())
    b = acc16.view(k * k, 16, Cout)
    want = (b[:, 0:3] + b[:, 3:6] + b[:, 6:9]).reshape(k * k * 3, Cout)
    assert torch.equal(got, want)
    if N * H * W <= 8192:
        xp = torch.zeros(N, 3, H + k, W + k, dtype=torch.float64)
        xp[:, :, pt:pt + H, pl:pl + W] = kx.permute(0, 3, 1, 2).double()
        gg = g.cpu().double().view(N, OH, OW, Cout).permute(0, 3, 1, 2)
        cols = torch.nn.functional.unfold(xp, (k, k), stride=s)[:, :, :]          # [N, 3*k*k, L] over the padded image
        Lw = (W + k - k) // s + 1
        cols = cols.view(N, 3, k, k, -1, Lw)[:, :, :, :, :OH, :OW]                 # [N, c, r, s, oh, ow]
        ref = torch.einsum('ncrsyx,nkyx->rsck', cols, gg).reshape(k * k * 3, Cout)
        assert torch.equal(got.cpu().double(), ref)


@pytest.mark.parametrize('N,H,W,C,Cout,k,halo', [(2, 16, 16, 64, 64, 3, True), (3, 30, 23, 64, 64, 3, True), (1, 32, 40, 64, 128, 3, True),
                                                   (3, 20, 13, 64, 64, 3, False),      # ragged: the halo kernel declines, im2col-TMA kernel
                                                   (2, 14, 14, 128, 128, 3, False),    # filter bank + two patches do not fit: im2col-TMA
                                                   (2, 14, 14, 256, 256, 3, False), (1, 7, 7, 512, 512, 3, False)])
def test_conv_dual_planes_equals_exact_sum(N, H, W, C, Cout, k, halo):
    """lbt_conv_i8_fprop_dual (16-bit source as hi / lo byte planes, two accumulators in tensor memory, one rounding) against
    the exact integer convolution in fp64: out = RN_fp32((256 * conv(hi, W) + conv(lo, W)) * 2^e) + addend, stride 1 'SAME',
    on the TMA halo kernel and on the im2col-TMA kernel — the arithmetic of lbt_gemm_i8_dual on im2col matrices, which it
    replaces for the 3x3 input gradients of BASELINE config 5."""
    from lbt_b200 import _lib, quantizer as Q
    pad = k // 2
    gen = torch.Generator().manual_seed(N * 100 + H + Cout)
    hi = torch.randint(-128, 128, (N, H, W, C), generator=gen, dtype=torch.int32)
    lo = torch.randint(0, 256, (N, H, W, C), generator=gen, dtype=torch.int32)
    wt = torch.randint(-128, 128, (Cout, k, k, C), generator=gen, dtype=torch.int32)
    addend = torch.randn(N * H * W, Cout, generator=gen)
    ib_s = torch.tensor(2, dtype=torch.int32, device='cuda')
    ib_w = torch.tensor(-1, dtype=torch.int32, device='cuda')
    e = -24
    out = torch.empty(N * H * W, Cout, dtype=torch.float32, device='cuda')
    wp = wt.reshape(Cout, k * k * C).to(torch.int8).cuda().contiguous()
    before = _lib.lib().lbt_conv_halo_launches()
    hi_d, lo_d, ad_d = hi.to(torch.int8).cuda(), lo.to(torch.uint8).cuda(), addend.cuda()     # (kept alive across the call)
    ok = _lib.try_call('lbt_conv_i8_fprop_dual', _lib.ptr(hi_d), _lib.ptr(lo_d), N, H, W, C,
                       _lib.ptr(wp), Q.MANT_S8, wp.stride(0), Cout, k, k, pad, pad, H, W, _lib.ptr(ib_s), _lib.ptr(ib_w), e,
                       _lib.ptr(out), Cout, _lib.ptr(ad_d), _lib.stream())
    assert ok and _lib.lib().lbt_conv_halo_launches() == before + (1 if halo else 0)
    torch.cuda.synchronize()
    assert _lib.lib().lbt_conv_debug_error() == 0
    src = (256 * hi + lo).permute(0, 3, 1, 2).double()
    ref = F.conv2d(src, wt.permute(0, 3, 1, 2).double(), padding=pad).permute(0, 2, 3, 1).reshape(N * H * W, Cout)
    want = (ref * 2.0 ** (e + 2 - 1)).float() + addend
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize('N,H,W,C,Cout,k,s', [(2, 16, 16, 64, 64, 3, 1), (3, 14, 14, 128, 256, 1, 1), (2, 17, 15, 16, 16, 3, 1),
                                                (2, 16, 16, 64, 128, 3, 2), (1, 8, 8, 256, 512, 3, 1)])
def test_wgrad_dual_planes_equals_two_passes(N, H, W, C, Cout, k, s):
    """lbt_conv_i8_wgrad_dual (both byte planes of a 16-bit gradient, two accumulators, one atomic per element) against two
    lbt_conv_i8_wgrad passes with alpha = 256 | 1: the int64 sums are equal."""
    from lbt_b200 import _lib, quantizer as Q
    OH, pt, _ = D.same_pad(H, k, s)
    OW, pl, _ = D.same_pad(W, k, s)
    gen = torch.Generator().manual_seed(N * 100 + C + Cout)
    x = torch.randint(0, 256, (N, H, W, C), generator=gen, dtype=torch.int32).to(torch.uint8).cuda()
    hi = torch.randint(-128, 128, (N * OH * OW, Cout), generator=gen, dtype=torch.int32).to(torch.int8).cuda()
    lo = torch.randint(0, 256, (N * OH * OW, Cout), generator=gen, dtype=torch.int32).to(torch.uint8).cuda()
    Kf = k * k * C
    a = torch.zeros(Kf, Cout, dtype=torch.int64, device='cuda')
    b = torch.zeros(Kf, Cout, dtype=torch.int64, device='cuda')
    _lib.call('lbt_conv_i8_wgrad_dual', _lib.ptr(x), Q.MANT_U8, N, H, W, C, _lib.ptr(hi), _lib.ptr(lo), Cout, k, k, s, s, pt, pl, OH, OW,
              _lib.ptr(a), 1, 0, _lib.stream())
    for g_, kind, alpha in ((hi, Q.MANT_S8, 256), (lo, Q.MANT_U8, 1)):
        _lib.call('lbt_conv_i8_wgrad', _lib.ptr(x), Q.MANT_U8, N, H, W, C, _lib.ptr(g_), kind, Cout, k, k, s, s, pt, pl, OH, OW,
                  _lib.ptr(b), alpha, 0, _lib.stream())
    torch.cuda.synchronize()
    assert _lib.lib().lbt_conv_debug_error() == 0
    assert torch.equal(a, b)

"""tcgen05 int8 GEMM vs an exact int64 reference: bit-for-bit (SURVEY.md §7.4 self-check)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lbt_b200 import gemm as G  # noqa: E402


def _operand(rng, rows, K, signed, ld=None):
    ld = ld or -(-K // 16) * 16
    buf = torch.zeros(rows, ld, dtype=torch.int8 if signed else torch.uint8)
    lo, hi = (-128, 128) if signed else (0, 256)
    vals = torch.from_numpy(rng.integers(lo, hi, size=(rows, K)).astype(np.int8 if signed else np.uint8))
    buf[:, :K] = vals
    buf[:, K:] = 77 if signed else 200          # pitch padding must never be read
    return buf.cuda()[:, :K], vals.to(torch.int64)


def _exact(a64, b64):
    return a64.double().cuda() @ b64.double().cuda().T          # |sum| < 2^53: exact in fp64


SHAPES = [(128, 16, 128), (128, 128, 128), (256, 256, 512), (1, 1, 1), (5, 3, 7), (130, 10, 64), (257, 33, 130),
          (1000, 100, 1000), (4096, 16, 144), (2048, 64, 576), (512, 1000, 512), (300, 400, 2048), (128, 256, 4608),
          (8192, 128, 1600)]


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('kinds', ['ss', 'us', 'su', 'uu'])
def test_gemm_f32_bit_exact(M, N, K, kinds):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A, a64 = _operand(rng, M, K, kinds[0] == 's')
    B, b64 = _operand(rng, N, K, kinds[1] == 's')
    out = G.gemm_i8(A, B, exp_const=0)
    torch.cuda.synchronize()
    assert G.debug_error() == 0, 'GEMM pipeline watchdog fired'
    ref = _exact(a64, b64)
    assert torch.equal(out.double(), ref.float().double()), (out.double() - ref).abs().max()


@pytest.mark.parametrize('M,N,K', [(128, 16, 128), (5, 3, 16), (257, 33, 144), (1000, 128, 1024), (4096, 64, 576), (300, 512, 2048),
                                   (3136, 256, 64), (512, 1000, 4608), (8192, 130, 1152)])
@pytest.mark.parametrize('b_signed', [True, False])
def test_gemm_dual_16bit_operand_bit_exact(M, N, K, b_signed):
    """lbt_gemm_i8_dual: a 16-bit A operand as its byte planes k = 256*hi + lo, two accumulators in tensor memory combined before
    ONE rounding: equal to RN_fp32(exact 16x8-bit integer dot * 2^e) (+ addend), bit for bit — extremes included."""
    rng = np.random.default_rng(M + N * 5 + K * 11 + b_signed)
    a16 = rng.integers(-32768, 32768, size=(M, K)).astype(np.int64)
    a16[0, :] = 32767
    a16[-1, :] = -32768
    ld = -(-K // 16) * 16
    hi = torch.zeros(M, ld, dtype=torch.int8)
    lo = torch.zeros(M, ld, dtype=torch.uint8)
    hi[:, :K] = torch.from_numpy((a16 >> 8).astype(np.int8))
    lo[:, :K] = torch.from_numpy((a16 & 255).astype(np.uint8))
    hi[:, K:] = 55
    lo[:, K:] = 201                                  # pitch padding must never be read
    B, b64 = _operand(rng, N, K, b_signed)
    if b_signed:
        b64[0, :] = -128
        B[0, :] = -128
    ibA = torch.tensor(-3, dtype=torch.int32, device='cuda')
    addend = torch.randn(M, N, device='cuda')
    out = G.gemm_i8_dual(hi.cuda()[:, :K], lo.cuda()[:, :K], B, ibA=ibA, exp_const=-18)
    out2 = G.gemm_i8_dual(hi.cuda()[:, :K], lo.cuda()[:, :K], B, ibA=ibA, exp_const=-18, addend=addend)
    torch.cuda.synchronize()
    assert G.debug_error() == 0, 'GEMM pipeline watchdog fired'
    ref = (_exact(torch.from_numpy(a16), b64) * 2.0 ** -21).float()      # |dot| < 2^15 * 2^8 * 4608 < 2^53: exact in fp64
    assert torch.equal(out, ref), float((out.double() - ref.double()).abs().max())
    assert torch.equal(out2, ref + addend)


def test_gemm_scale_and_bias_from_device_ranges():
    rng = np.random.default_rng(1)
    M, N, K = 384, 48, 300
    A, a64 = _operand(rng, M, K, False)
    B, b64 = _operand(rng, N, K, True)
    ibA = torch.tensor(1, dtype=torch.int32, device='cuda')
    ibB = torch.tensor(-2, dtype=torch.int32, device='cuda')
    bias = torch.randn(N, device='cuda')
    out = G.gemm_i8(A, B, ibA=ibA, ibB=ibB, exp_const=-(9 - 1) - (8 - 1), bias=bias)
    scale = 2.0 ** (1 - 2 - 8 - 7)
    ref = (_exact(a64, b64).float() * scale) + bias            # RN(exact * 2^e) then one rounded add
    assert torch.equal(out, ref)


@pytest.mark.parametrize('M,N,K,splits', [(144, 16, 262144, 0), (27, 16, 65536 * 3 + 5, 0), (576, 64, 16384, 7),
                                          (4608, 512, 12544, 0), (128, 16, 128, 4)])
def test_gemm_split_k_acc64_exact(M, N, K, splits):
    rng = np.random.default_rng(K)
    A, a64 = _operand(rng, M, K, False)
    B, b64 = _operand(rng, N, K, True)
    acc = torch.zeros(M, N, dtype=torch.int64, device='cuda')
    G.gemm_i8_acc64(A, B, acc, alpha=1, k_splits=splits)
    G.gemm_i8_acc64(A, B, acc, alpha=2, k_splits=splits)       # hi/lo recombination uses alpha
    torch.cuda.synchronize()
    assert G.debug_error() == 0
    ref = _exact(a64, b64).to(torch.int64)
    assert torch.equal(acc, 3 * ref)
    W = torch.randn(M, N, device='cuda')
    ib = torch.tensor(3, dtype=torch.int32, device='cuda')
    out = G.acc64_finalize(acc, ibA=ib, exp_const=-20, add=W, add_scale=4e-4)
    ref32 = (3 * ref).float() * 2.0 ** -17 + torch.tensor(4e-4, device='cuda') * W
    assert torch.equal(out, ref32)


def test_gemm_worst_case_magnitudes_do_not_overflow():
    """All operands at the extreme values and K at the exactness bound of a single s32 accumulator."""
    K = 65536
    A = torch.full((128, K), 255, dtype=torch.uint8, device='cuda')
    B = torch.full((16, K), -128, dtype=torch.int8, device='cuda')
    out = G.gemm_i8(A, B)
    assert torch.all(out == float(-255 * 128 * K))
    with pytest.raises(Exception):
        G.gemm_i8(torch.zeros(128, K + 128, dtype=torch.uint8, device='cuda'),
                  torch.zeros(16, K + 128, dtype=torch.int8, device='cuda'))


# ---- CTA-pair kernel (tcgen05 cta_group::2, 256x256 tiles): taken for large fp32-epilogue GEMMs with N > 128 ----

@pytest.mark.parametrize('M,N,K', [(2048, 2560, 384), (2304 + 57, 2048 + 24, 1000), (4096, 4096, 4096), (256, 19000, 130),
                                   (18944 + 129, 256, 64)])
@pytest.mark.parametrize('kinds', ['ss', 'us'])
def test_gemm_pair_kernel_bit_exact(M, N, K, kinds):
    from lbt_b200 import _lib
    rng = np.random.default_rng(M + N + K)
    A, a64 = _operand(rng, M, K, kinds[0] == 's')
    B, b64 = _operand(rng, N, K, kinds[1] == 's')
    bias = torch.randn(N, device='cuda')
    addend = torch.randn(M, N, device='cuda')
    ib = torch.tensor(-3, dtype=torch.int32, device='cuda')
    try:
        _lib.lib().lbt_gemm_set_pair(1)
        out = G.gemm_i8(A, B, exp_const=0)
        out2 = G.gemm_i8(A, B, ibA=ib, exp_const=-9, bias=bias, addend=addend)
        torch.cuda.synchronize()
        assert G.debug_error() == 0, 'GEMM pipeline watchdog fired'
        _lib.lib().lbt_gemm_set_pair(0)
        single = G.gemm_i8(A, B, exp_const=0)
        single2 = G.gemm_i8(A, B, ibA=ib, exp_const=-9, bias=bias, addend=addend)
        torch.cuda.synchronize()
    finally:
        _lib.lib().lbt_gemm_set_pair(1)
    ref = _exact(a64, b64)
    assert torch.equal(out.double(), ref.float().double()), (out.double() - ref).abs().max()
    assert torch.equal(out, single)
    assert torch.equal(out2, single2)
    assert torch.equal(out2, (ref.float() * 2.0 ** -12 + bias) + addend)


@pytest.mark.parametrize('M,N,K', [(128, 16, 128), (300, 200, 1024), (2048, 512, 4096)])
def test_gemm_16bit_operands_exact(M, N, K):
    """16-bit mantissas on BOTH operands (config 3's 16-bit point): four 8-bit tensor-core passes, exact in int64."""
    rng = np.random.default_rng(M + K)
    a = torch.from_numpy(rng.integers(-32768, 32768, (M, K)).astype(np.int16)).cuda()
    b = torch.from_numpy(rng.integers(-32768, 32768, (N, K)).astype(np.int16)).cuda()
    a[0, :] = -32768
    b[0, :] = -32768                                      # the extreme corner: K * 2^30
    acc = G.gemm_i16_acc64(a, b)
    torch.cuda.synchronize()
    assert G.debug_error() == 0
    # int64 reference in chunks (the products reach 2^30: fp64 accumulation is exact up to K = 2^23)
    ref = (a.double() @ b.double().T).to(torch.int64)
    assert torch.equal(acc, ref)
    ib = torch.tensor(-1, dtype=torch.int32, device='cuda')
    out = G.acc64_finalize(acc, ibA=ib, ibB=ib, exp_const=-30)
    assert torch.equal(out, (ref.double() * 2.0 ** -32).float())

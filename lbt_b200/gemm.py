"""Host wrappers of the tcgen05 int8 GEMM (csrc/gemm_i8.cu) — raw pointers across the C ABI."""
import ctypes

import torch

from . import _lib
from .quantizer import MANT_S8, MANT_U8

EPI_F32, EPI_ACC64 = 0, 1


def _kind(t):
    if t.dtype == torch.int8:
        return MANT_S8
    if t.dtype == torch.uint8:
        return MANT_U8
    raise _lib.LbtError('GEMM operands must be int8/uint8 mantissas, got %s' % t.dtype)


def _check_operand(t, K):
    if t.dim() != 2 or t.stride(1) != 1 or t.shape[1] != K:
        raise _lib.LbtError('GEMM operand must be a K-major [rows, K] matrix with unit inner stride')
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), K)


def gemm_i8(A, B, *, ibA=None, ibB=None, exp_const=0, bias=None, out=None, bnq=None, addend=None):
    """fp32 out[M,N] = (A[M,K] @ B[N,K]^T) * 2^(exp_const + ibA + ibB) (+ bias).

    ``bnq = (QSiteStruct, k_out int8 [M,N], sums int64 [2N], rows_per_image)`` switches to the fused re-quantising
    epilogue (no fp32 output): k_out = Q_site(result), sums += (k, k^2) per column.  ``addend`` (fp32, the shape and row
    pitch of ``out``) is added to the fp32 result in the epilogue."""
    M, K = A.shape
    N = B.shape[0]
    lda, ldb = _check_operand(A, K), _check_operand(B, K)
    if bnq is not None:
        qs, k_out, sums, rpi = bnq
        _lib.call('lbt_gemm_i8', _lib.ptr(A), _kind(A), lda, _lib.ptr(B), _kind(B), ldb, M, N, K, EPI_F32,
                  _lib.ptr(ibA), _lib.ptr(ibB), int(exp_const), _lib.ptr(bias), None, None, N, 1, 1,
                  ctypes.addressof(qs), _lib.ptr(k_out), _lib.ptr(sums), int(rpi), None, _lib.stream(),
                  meta=dict(ops=2 * M * N * K, bytes=M * K + N * K + M * N))
        return None
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    _lib.call('lbt_gemm_i8', _lib.ptr(A), _kind(A), lda, _lib.ptr(B), _kind(B), ldb, M, N, K, EPI_F32,
                                      _lib.ptr(ibA), _lib.ptr(ibB), int(exp_const), _lib.ptr(bias), _lib.ptr(out), None,
                                      out.stride(0), 1, 1, None, None, None, 0, _lib.ptr(addend), _lib.stream(),
              meta=dict(ops=2 * M * N * K, bytes=M * K + N * K + 4 * M * N * (2 if addend is not None else 1)))
    return out


def gemm_i8_dual(A_hi, A_lo, B, *, ibA=None, ibB=None, exp_const=0, out=None, addend=None):
    """fp32 out[M,N] = ((256 * A_hi + A_lo)[M,K] @ B[N,K]^T) * 2^(exp_const + ibA + ibB) (+ addend), one rounding: a 9..16-bit
    A operand as its two byte planes (lbt_gemm_i8_dual: both halves per K block, two accumulators in tensor memory)."""
    M, K = A_hi.shape
    N = B.shape[0]
    if A_hi.dtype != torch.int8 or A_lo.dtype != torch.uint8 or A_lo.shape != A_hi.shape or A_lo.stride() != A_hi.stride():
        raise _lib.LbtError('gemm_i8_dual takes an int8 high plane and a uint8 low plane of the same shape and pitch')
    lda, ldb = _check_operand(A_hi, K), _check_operand(B, K)
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A_hi.device)
    _lib.call('lbt_gemm_i8_dual', _lib.ptr(A_hi), _lib.ptr(A_lo), lda, _lib.ptr(B), _kind(B), ldb, M, N, K, _lib.ptr(ibA), _lib.ptr(ibB),
              int(exp_const), _lib.ptr(out), out.stride(0), _lib.ptr(addend), _lib.stream(),
              meta=dict(ops=2 * M * N * K, bytes=2 * M * K + N * K + 4 * M * N * (2 if addend is not None else 1)))
    return out


def gemm_i8_acc64(A, B, acc64, *, alpha=1, k_splits=0):
    """acc64[M,N] += alpha * (A[M,K] @ B[N,K]^T), exact, K split across the SMs."""
    M, K = A.shape
    N = B.shape[0]
    lda, ldb = _check_operand(A, K), _check_operand(B, K)
    if acc64.dtype != torch.int64 or acc64.shape != (M, N) or not acc64.is_contiguous():
        raise _lib.LbtError('acc64 must be a contiguous int64 [M, N] tensor')
    if k_splits <= 0:
        tiles = -(-M // 128) * -(-N // min(256, max(16, 1 << (max(N, 1) - 1).bit_length())))
        sms = torch.cuda.get_device_properties(A.device).multi_processor_count
        k_splits = max(1, sms // max(1, tiles))
    _lib.call('lbt_gemm_i8', _lib.ptr(A), _kind(A), lda, _lib.ptr(B), _kind(B), ldb, M, N, K, EPI_ACC64,
                                      None, None, 0, None, None, _lib.ptr(acc64), N, int(alpha), int(k_splits),
                                      None, None, None, 0, None, _lib.stream(),
              meta=dict(ops=2 * M * N * K, bytes=M * K + N * K + 8 * M * N))
    return acc64


def split_s16(m16):
    """s16 mantissas -> (hi s8, lo u8) with k = 256 * hi + lo (lbt_split_s16)."""
    hi = torch.empty(m16.shape, dtype=torch.int8, device=m16.device)
    lo = torch.empty(m16.shape, dtype=torch.uint8, device=m16.device)
    _lib.call('lbt_split_s16', _lib.ptr(m16), m16.numel(), _lib.ptr(hi), _lib.ptr(lo), _lib.stream(), meta=dict(bytes=m16.numel() * 4))
    return hi, lo


def gemm_i16_acc64(A16, B16, acc64=None):
    """acc64[M,N] += A16[M,K] @ B16[N,K]^T for s16 mantissas (DFXP quantisers of 9..16 bits on BOTH operands: the 16-bit point
    of the bit-width sweep, BASELINE config 3), exact: a = 256*ah + al, b = 256*bh + bl, so
    a*b = 65536*ah*bh + 256*(ah*bl + al*bh) + al*bl — four passes of the 8-bit tensor-core kernel into the int64 accumulator
    (alpha = 65536, 256, 256, 1).  K must be a multiple of 16 (operand row pitch)."""
    M, K = A16.shape
    N = B16.shape[0]
    if A16.dtype != torch.int16 or B16.dtype != torch.int16 or B16.shape[1] != K:
        raise _lib.LbtError('gemm_i16_acc64 takes int16 [M,K] and [N,K] mantissa tensors')
    if acc64 is None:
        acc64 = torch.zeros(M, N, dtype=torch.int64, device=A16.device)
    ah, al = split_s16(A16.contiguous())
    bh, bl = split_s16(B16.contiguous())
    for a, b, alpha in ((ah, bh, 65536), (ah, bl, 256), (al, bh, 256), (al, bl, 1)):
        gemm_i8_acc64(a, b, acc64, alpha=alpha)
    return acc64


def acc64_finalize(acc64, *, ibA=None, ibB=None, exp_const=0, add=None, add_scale=0.0, out=None):
    """fp32 out = acc64 * 2^(exp_const + ibA + ibB) + add_scale * add."""
    if out is None:
        out = torch.empty(acc64.shape, dtype=torch.float32, device=acc64.device)
    _lib.call('lbt_acc64_finalize', _lib.ptr(acc64), acc64.numel(), _lib.ptr(ibA), _lib.ptr(ibB),
                                             int(exp_const), _lib.ptr(add), float(add_scale), _lib.ptr(out),
                                             _lib.stream())
    return out


def debug_error():
    return int(_lib.lib().lbt_gemm_debug_error())

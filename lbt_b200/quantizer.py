"""Host side of the DFXP quantiser: the reference's three functions, same names and argument order.

    weight_quantization(X, target_overflow_rate, bits, integer_bits, stochastic=False)   dfxp:4
    overflow_rate(X, bits, integer_bits)                                                 dfxp:48
    update_range(X, target_overflow_rate, bits, integer_bits)                            dfxp:70

(``dfxp:N`` = /root/reference/dynamic_fixed_point.py:N).  ``integer_bits`` is an int32 CUDA scalar
tensor owned by the caller (the layer's ``*_range`` variable, dfxp:161-171); the kernel reads it and,
like the reference's ``update_range`` op fetched in the same ``sess.run`` (trainer.py:157), updates it
after use (read-then-update).  Everything runs in ``lbt_quantize`` (csrc/quantize.cu) — there is no
PyTorch or CPU implementation behind these functions.
"""
import torch

from . import _lib

ROUND_NEAREST, ROUND_NOISE, ROUND_PHILOX = 0, 1, 2
STATS_MINMAX = 0x100       # OR into the mode: min/max overflow statistics (exact enough for target_overflow_rate == 0)
MANT_NONE, MANT_S8, MANT_U8, MANT_S16, MANT_S9C3 = 0, 1, 2, 3, 4
_MANT_DTYPE = {MANT_S8: torch.int8, MANT_U8: torch.uint8, MANT_S16: torch.int16}
_MANT_BYTES = {MANT_NONE: 0, MANT_S8: 1, MANT_U8: 1, MANT_S16: 2, MANT_S9C3: 6}     # per element, for the roofline
CNT_WORDS = 4


def rows_view(x):
    """(n_outer, n_inner) of ``x`` in MEMORY order = [dim0, prod(rest)] of the TF-layout tensor.

    Accepts default-contiguous tensors of any rank and 4-D channels_last tensors (logical NCHW whose
    memory is NHWC, i.e. the reference's layout)."""
    if x.dim() == 0:
        return 1, 1
    if not (x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))):
        raise _lib.LbtError('lbt_b200.quantize needs a contiguous (or channels_last) tensor')
    n_outer = x.shape[0]
    n_inner = x.numel() // n_outer if n_outer else 0
    return n_outer, n_inner


def new_counters(device):
    return torch.zeros(CNT_WORDS, dtype=torch.int64, device=device)


def quantize(x, bits, integer_bits, *, target_overflow_rate=0.0, mode=ROUND_NEAREST, noise=None, seed=0, offset=0,
             dev_step=None, want_fp32=True, mant_kind=MANT_NONE, counters=None, update_range=True, out=None,
             out_mant=None):
    """One ``lbt_quantize`` launch.  Returns (fake-quant fp32 tensor or None, mantissa tensor or None)."""
    if x.dtype != torch.float32:
        raise _lib.LbtError('DFXP quantiser takes fp32 input (like the reference), got %s' % x.dtype)
    if integer_bits.dtype != torch.int32:
        raise _lib.LbtError('integer_bits must be an int32 CUDA tensor')
    n_outer, n_inner = rows_view(x)
    if want_fp32 and out is None:
        out = torch.empty_like(x)
    if mant_kind != MANT_NONE and out_mant is None:
        out_mant = torch.empty_like(x, dtype=_MANT_DTYPE[mant_kind])
    if counters is None and update_range:
        counters = new_counters(x.device)
    if noise is not None:
        if noise.numel() != n_inner or noise.dtype != torch.float32:
            raise _lib.LbtError('noise must be fp32 with X.shape[1:] elements (%d), got %d' % (n_inner, noise.numel()))
        noise = noise.contiguous()
    _lib.call('lbt_quantize', _lib.ptr(x), n_outer, n_inner, int(bits), _lib.ptr(integer_bits),
                              float(target_overflow_rate), int(mode), _lib.ptr(noise), int(seed), int(offset),
                              _lib.ptr(dev_step), _lib.ptr(out) if want_fp32 else None, _lib.ptr(out_mant),
                              int(mant_kind), _lib.ptr(counters), 1 if update_range else 0, _lib.stream(),
              meta=dict(bytes=n_outer * n_inner * (4 + (4 if want_fp32 else 0) + _MANT_BYTES[int(mant_kind)])))
    return (out if want_fp32 else None), out_mant


class _QuantizeSTE(torch.autograd.Function):
    """Straight-through estimator of dfxp:30,38: ``lambda dy: dy`` for every element (no clip mask)."""

    @staticmethod
    def forward(ctx, x, bits, integer_bits, target, mode, noise, seed, offset):
        q, _ = quantize(x, bits, integer_bits, target_overflow_rate=target, mode=mode, noise=noise, seed=seed,
                        offset=offset)
        return q

    @staticmethod
    def backward(ctx, dy):
        return dy, None, None, None, None, None, None, None


def weight_quantization(X, target_overflow_rate, bits, integer_bits, stochastic=False, *, noise=None, seed=0,
                        offset=0):
    """dfxp:4-45.  Fake-quantise ``X`` to DFXP and update ``integer_bits`` (read-then-update).

    ``stochastic=True`` takes either an explicit ``noise`` tensor of shape ``X.shape[1:]`` (parity
    with the reference when fed its noise) or, when ``noise`` is None, the in-kernel Philox stream
    keyed by (seed, offset)."""
    assert 1 <= bits <= 32, 'invalid value for bits: %d' % bits          # dfxp:21
    if bits == 32:                                                       # dfxp:22-23
        return X
    mode = ROUND_NEAREST if not stochastic else (ROUND_NOISE if noise is not None else ROUND_PHILOX)
    return _QuantizeSTE.apply(X, bits, integer_bits, float(target_overflow_rate), mode, noise, seed, offset)


def _counts(X, bits, integer_bits):
    """(n_over, n_over_half, numel) as a device int64 tensor, without touching integer_bits."""
    cnt = new_counters(X.device)
    quantize(X, int(bits), integer_bits, want_fp32=False, counters=cnt, update_range=False)   # statistics only
    return cnt


def overflow_rate(X, bits, integer_bits):
    """dfxp:48-67: (overflow_rate(X), overflow_rate(2X)) as fp32 CUDA scalars."""
    cnt = _counts(X, bits, integer_bits)
    n = cnt[2].clamp(min=1).to(torch.float32)
    return cnt[0].to(torch.float32) / n, cnt[1].to(torch.float32) / n


def update_range(X, target_overflow_rate, bits, integer_bits):
    """dfxp:70-94: apply the controller to ``integer_bits`` in place; returns it."""
    quantize(X, int(bits), integer_bits, target_overflow_rate=target_overflow_rate, want_fp32=False,
             update_range=True)                                           # statistics + controller only
    return integer_bits


def noise_fill(n_inner, seed, offset, device='cuda', dev_step=None):
    """The Philox noise vector ``lbt_quantize(mode=ROUND_PHILOX)`` uses for (seed, offset)."""
    u = torch.empty(int(n_inner), dtype=torch.float32, device=device)
    _lib.call('lbt_noise_fill', _lib.ptr(u), int(n_inner), int(seed), int(offset), _lib.ptr(dev_step),
                                         _lib.stream())
    return u


def make_offset(qid, step=0):
    """(quantiser id, step) -> 64-bit Philox offset: low word = id, high word = step."""
    return ((int(step) & 0xFFFFFFFF) << 32) | (int(qid) & 0xFFFFFFFF)

"""CPU restatement of the reference's dynamic-fixed-point path (quantiser + quantised layers).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — never imported by the product path.
PARITY UNPINNED by the reference (it has no tests); pinned here by SURVEY.md App. A.4 KATs.

Everything is fp32 on the CPU, TF layouts (NHWC activations, HWIO conv weights, [in, out] dense
weights), TF1 op semantics:

* ``tf.round`` = round-half-to-even, ``tf.clip_by_value(x, lo, hi)`` = min(max(x, lo), hi),
  ``tf.random_uniform`` in [0, 1);
* ``tf.nn.conv2d`` 'SAME': out = ceil(in / s), pad_total = max((out-1)*s + k - in, 0),
  pad_before = pad_total // 2 (extra pixel on the bottom/right);
* ``tf.nn.moments`` = biased variance; ``tf.nn.max_pool`` 'SAME' ignores padded cells;
* ``tf.nn.dropout(x, keep)`` = x / keep * floor(keep + u);
* ``tf.train.MomentumOptimizer`` (non-Nesterov): a <- mu*a + g ; w <- w - lr*a.

Each function cites the reference lines (``dfxp:N`` = /root/reference/dynamic_fixed_point.py:N) it
restates.  Decisions on the reference's undefined behaviour:

* range read/update race (trainer.py:157) -> **read-then-update**: a quantiser call at step t uses
  the range left by step t-1, then applies the controller's +-1;
* ``2 ** (bits - integer_bits - 1)`` is an int32 ``tf.pow`` in the reference (dfxp:27) and overflows
  for an exponent >= 31; here the multiplier is an exact fp32 power of two with the exponent clamped
  to [-126, 126] — parity tests stay out of that regime.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import philox

# --------------------------------------------------------------------------------------------
# Range state + noise sources
# --------------------------------------------------------------------------------------------


class Range:
    """A ``*_range`` int32 variable (dfxp:161-171): the integer-bit count, excluding sign."""

    def __init__(self, value=2):
        self.value = int(value)

    def __int__(self):
        return self.value

    def __repr__(self):
        return 'Range(%d)' % self.value


class NumpyNoise:
    """Unstructured noise source (stands in for the unseeded tf.random_uniform, dfxp:36)."""

    def __init__(self, seed=0):
        self.rng = np.random.default_rng(seed)

    def __call__(self, qid, shape):
        # fp32 uniform in [0, 1): 24 random bits, like the CUDA stream
        r = self.rng.integers(0, 1 << 24, size=shape, dtype=np.uint32)
        return (r.astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


class PhiloxNoise:
    """Counter-based noise keyed by (seed, quantiser id, step) — what lbt_b200 generates in-kernel."""

    def __init__(self, seed=0):
        self.seed = seed
        self.step = 0

    def __call__(self, qid, shape):
        n = int(np.prod(shape)) if len(shape) else 1
        u = philox.noise(n, self.seed, philox.make_offset(qid, self.step))
        return u.reshape(shape)


class Context:
    """Per-model bookkeeping the TF graph did implicitly: quantiser ids, noise source, step count."""

    def __init__(self, noise=None, exact=False):
        self.noise = noise if noise is not None else NumpyNoise(0)
        # exact=False: fp32 accumulation everywhere, like the reference's cuDNN/cuBLAS/Eigen kernels.
        # exact=True : conv/matmul/batch-moment accumulations are carried out in fp64 (exact for DFXP
        #   mantissa products) and rounded ONCE to fp32 — the arithmetic of the integer tensor-core
        #   path, which must then match bit for bit.
        self.exact = exact
        self.n_quant = 0
        self.last_counts = {}       # qid -> (n_over, n_over_half, numel) of the latest call

    def new_qid(self):
        q = self.n_quant
        self.n_quant += 1
        return q


# --------------------------------------------------------------------------------------------
# The quantiser (dfxp:4-94) — numpy fp32
# --------------------------------------------------------------------------------------------


def _multiplier_limit(bits, integer_bits):
    """(m, L) of dfxp:27-28 / 34-35 / 60-61 as exact fp32 values."""
    f = int(bits) - int(integer_bits) - 1
    f = max(-126, min(126, f))                      # App. E-3 divergence, documented above
    m = np.float32(math.ldexp(1.0, f))
    L = np.float32(math.ldexp(1.0, int(bits) - 1))
    return m, L


def quantize_nearest(x, bits, integer_bits):
    """``identity`` of dfxp:25-30: round_half_even(clip(x*m, -L, L-1)) / m.  Returns (q, mantissa)."""
    x = np.asarray(x, dtype=np.float32)
    m, L = _multiplier_limit(bits, integer_bits)
    hi = np.float32(L - np.float32(1))
    y = x * m                                                   # dfxp:29  X * multiplier
    k = np.rint(np.minimum(np.maximum(y, -L), hi))              # clip then round (half to even)
    return (k / m).astype(np.float32), k


def quantize_stochastic(x, bits, integer_bits, u):
    """``stochastic_identity`` of dfxp:32-38: floor(clip(x*m + u, -L, L-1)) / m.

    ``u`` has shape ``x.shape[1:]`` and broadcasts over dim 0 (dfxp:36).  Returns (q, mantissa)."""
    x = np.asarray(x, dtype=np.float32)
    u = np.asarray(u, dtype=np.float32)
    assert tuple(u.shape) == tuple(x.shape[1:]), (u.shape, x.shape)
    m, L = _multiplier_limit(bits, integer_bits)
    hi = np.float32(L - np.float32(1))
    y = (x * m).astype(np.float32) + u                          # two fp32 roundings, dfxp:36
    k = np.floor(np.minimum(np.maximum(y, -L), hi))
    return (k / m).astype(np.float32), k


def overflow_counts(x, bits, integer_bits):
    """Integer numerators of ``overflow_rate`` (dfxp:60-67): (#overflow at L, #overflow at L/2)."""
    x = np.asarray(x, dtype=np.float32)
    m, L = _multiplier_limit(bits, integer_bits)
    y = x * m                                                   # dfxp:62
    n1 = int(np.count_nonzero(y >= L)) + int(np.count_nonzero(y < -L))          # dfxp:63-64
    h = np.float32(L / np.float32(2))
    n2 = int(np.count_nonzero(y >= h)) + int(np.count_nonzero(y < -h))          # dfxp:65-66
    return n1, n2


def overflow_rate(x, bits, integer_bits):
    """dfxp:48-67 — (overflow_rate(X), overflow_rate(2X)) as fp32 means."""
    n1, n2 = overflow_counts(x, bits, integer_bits)
    n = max(1, int(np.asarray(x).size))
    return np.float32(np.float32(n1) / np.float32(n)), np.float32(np.float32(n2) / np.float32(n))


def range_delta(n1, n2, numel, target_overflow_rate):
    """The tf.cond tree of dfxp:84-92 on integer counts.

    r > t  <=>  n > t * numel ; evaluated in fp32 like the reference's reduce_mean comparison for
    t == 0 (the only value the reference uses) this is exactly n > 0 / n == 0."""
    t = np.float32(target_overflow_rate)
    n = np.float32(max(1, numel))
    r1 = np.float32(np.float32(n1) / n)
    r2 = np.float32(np.float32(n2) / n)
    if r1 > t:
        return 1
    if r2 <= t:
        return -1
    return 0


def update_range(x, target_overflow_rate, bits, integer_bits):
    """dfxp:70-94.  ``integer_bits`` is a Range; assigns min(bits-1, ib+delta) and returns it."""
    n1, n2 = overflow_counts(x, bits, int(integer_bits))
    d = range_delta(n1, n2, int(np.asarray(x).size), target_overflow_rate)
    integer_bits.value = min(int(bits) - 1, int(integer_bits) + d)              # dfxp:94
    return integer_bits.value


def weight_quantization(x, target_overflow_rate, bits, integer_bits, stochastic=False, noise=None):
    """dfxp:4-45 on numpy arrays.  ``integer_bits`` is a Range (read, then updated).

    Returns the fake-quantised fp32 array.  bits == 32 is a pass-through with no range update
    (dfxp:22-23)."""
    assert 1 <= bits <= 32, 'invalid value for bits: %d' % bits                 # dfxp:21
    if bits == 32:
        return np.asarray(x, dtype=np.float32)
    ib = int(integer_bits)
    if not stochastic:
        q, _ = quantize_nearest(x, bits, ib)
    else:
        assert noise is not None, 'stochastic rounding needs the explicit noise tensor'
        q, _ = quantize_stochastic(x, bits, ib, noise)
    update_range(x, target_overflow_rate, bits, integer_bits)                   # dfxp:40-41
    return q


def mantissa(q, bits, integer_bits):
    """Integer mantissa of a fake-quant value: q * 2^f (exact)."""
    m, _ = _multiplier_limit(bits, integer_bits)
    return np.asarray(q, dtype=np.float32) * m


# --------------------------------------------------------------------------------------------
# torch glue: quantiser with straight-through gradient (dfxp:30, 38: grad = dy, no clip mask)
# --------------------------------------------------------------------------------------------


class _STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, q):
        return q.clone()

    @staticmethod
    def backward(ctx, dy):
        return dy, None


class Quantizer:
    """One ``weight_quantization`` call site with its own range variable (App. C census)."""

    def __init__(self, ctx, bits, init_range=2, target_overflow_rate=0.0, stochastic=True):
        self.ctx = ctx
        self.qid = ctx.new_qid()
        self.bits = int(bits)
        self.range = Range(init_range)
        self.t = float(target_overflow_rate)
        self.stochastic = stochastic
        self.last_noise = None
        self.last_range_used = None

    def __call__(self, x):
        """x: torch fp32 tensor -> fake-quant tensor with STE gradient; updates the range."""
        if self.bits == 32:
            return x
        xn = x.detach().contiguous().numpy()
        ib = int(self.range)
        self.last_range_used = ib
        if self.stochastic:
            u = self.ctx.noise(self.qid, tuple(xn.shape[1:]))
            self.last_noise = u
            q, _ = quantize_stochastic(xn, self.bits, ib, u)
        else:
            q, _ = quantize_nearest(xn, self.bits, ib)
        n1, n2 = overflow_counts(xn, self.bits, ib)
        self.ctx.last_counts[self.qid] = (n1, n2, xn.size)
        d = range_delta(n1, n2, xn.size, self.t)
        self.range.value = min(self.bits - 1, ib + d)
        return _STE.apply(x, torch.from_numpy(q))


# --------------------------------------------------------------------------------------------
# TF op restatements
# --------------------------------------------------------------------------------------------


def same_pad(in_size, k, s):
    """TF 'SAME' padding: (out, pad_before, pad_after)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2, total - total // 2


def tf_conv2d(x_nhwc, w_hwio, strides, padding):
    """tf.nn.conv2d (NHWC x HWIO) via F.conv2d on permuted views."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    sh, sw = strides[1], strides[2]
    x = x_nhwc.permute(0, 3, 1, 2)
    if padding == 'SAME':
        _, pt, pb = same_pad(x.shape[2], kh, sh)
        _, pl, pr = same_pad(x.shape[3], kw, sw)
        x = F.pad(x, (pl, pr, pt, pb))
    else:
        assert padding == 'VALID'
    w = w_hwio.permute(3, 2, 0, 1)
    y = F.conv2d(x, w, stride=(sh, sw))
    return y.permute(0, 2, 3, 1)


def tf_max_pool(x_nhwc, ksize, strides, padding):
    kh, kw, sh, sw = ksize[1], ksize[2], strides[1], strides[2]
    x = x_nhwc.permute(0, 3, 1, 2)
    if padding == 'SAME':
        _, pt, pb = same_pad(x.shape[2], kh, sh)
        _, pl, pr = same_pad(x.shape[3], kw, sw)
        x = F.pad(x, (pl, pr, pt, pb), value=float('-inf'))
    y = F.max_pool2d(x, (kh, kw), (sh, sw))
    return y.permute(0, 2, 3, 1)


def tf_avg_pool(x_nhwc, ksize, strides, padding):
    assert padding == 'VALID', 'only VALID average pooling is used by the reference models'
    kh, kw, sh, sw = ksize[1], ksize[2], strides[1], strides[2]
    y = F.avg_pool2d(x_nhwc.permute(0, 3, 1, 2), (kh, kw), (sh, sw))
    return y.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------------------------
# Layers (dfxp:97-1053).  forward(X) / backward(grad) / grads_and_vars(), TF-style ctors.
# --------------------------------------------------------------------------------------------


class Layer_q:
    """dfxp:97-126.  Default backward = autodiff of this layer's y w.r.t. its X (dfxp:113)."""

    def forward(self, X):
        self.X = X
        self.y = self.X
        return self.y

    def backward(self, grad, stochastic=True):
        if self.y is self.X:
            return grad
        return torch.autograd.grad(self.y, self.X, grad, retain_graph=True)[0]

    def grads_and_vars(self):
        return []

    def variables(self):
        """Trainable variables in grads_and_vars() order (available before any backward)."""
        return []

    def quantizers(self):
        return []


def _leaf(x):
    """Each layer differentiates only its own sub-graph (tf.gradients(self.y, self.X, g))."""
    return x.detach().requires_grad_(True)


class Conv2d_q(Layer_q):
    """dfxp:224-316 (Conv2d_pq dfxp:129-221 is a copy)."""

    def __init__(self, ctx, name, bits, ksize, strides, padding, use_bias=True, weight_decay=0,
                 target_overflow_rate=0, input_range=2, weight_range=2, bias_range=2, grad_range=2,
                 rng=None, grad_bits=None):
        h, w, Cin, Cout = self.ksize = list(ksize)
        self.strides, self.padding, self.name, self.use_bias = strides, padding, name, use_bias
        limit = (3 / (h * w * Cin)) ** 0.5                                        # dfxp:247-248
        rng = rng if rng is not None else np.random.default_rng(0)
        self.W = torch.from_numpy(rng.uniform(-limit, limit, size=self.ksize).astype(np.float32)).requires_grad_(True)
        self.bits = bits
        self.weight_decay = weight_decay
        # creation order == quantiser id order: X, W, [b], grad  (forward order then backward)
        self.qX = Quantizer(ctx, bits + 1, input_range, target_overflow_rate)    # dfxp:287-288 bits+1
        self.qW = Quantizer(ctx, bits, weight_range, target_overflow_rate)       # dfxp:289-290
        if use_bias:
            self.b = torch.zeros(Cout, requires_grad=True)                        # dfxp:264
            self.qb = Quantizer(ctx, bits, bias_range, target_overflow_rate)     # dfxp:294-295
        self.qG = Quantizer(ctx, grad_bits or bits, grad_range, target_overflow_rate)  # dfxp:300-301

    def forward(self, X):
        self.X = _leaf(X)
        self.Xq = self.qX(self.X)
        self.Wq = self.qW(self.W)
        if self.qX.ctx.exact:   # exactly-rounded accumulation (see Context)
            self.y = tf_conv2d(self.Xq.double(), self.Wq.double(), self.strides, self.padding).float()
        else:
            self.y = tf_conv2d(self.Xq, self.Wq, self.strides, self.padding)     # dfxp:291
        if self.use_bias:
            self.bq = self.qb(self.b)
            self.y = self.y + self.bq                                             # dfxp:296
        return self.y.detach()

    def backward(self, grad, stochastic=True):
        self.gradq = self.qG(grad).detach()                                       # dfxp:300
        wrt = [self.X, self.W] + ([self.b] if self.use_bias else [])
        g = torch.autograd.grad(self.y, wrt, self.gradq)
        self.dW = g[1] + 2 * self.weight_decay * self.W.detach()                  # dfxp:302
        if self.use_bias:
            self.db = g[2]                                                        # dfxp:304
        return g[0]                                                               # dfxp:305

    def grads_and_vars(self):
        r = [(self.dW, self.W)]
        if self.use_bias:
            r.append((self.db, self.b))
        return r

    def variables(self):
        return [self.W] + ([self.b] if self.use_bias else [])

    def quantizers(self):
        return [self.qX, self.qW] + ([self.qb] if self.use_bias else []) + [self.qG]


Conv2d_pq = Conv2d_q


class Dense_q(Layer_q):
    """dfxp:319-395, 441-470 (the eps/accu-value accumulator at :397-451 is dead code)."""

    def __init__(self, ctx, name, bits, in_units, units, use_bias=True, weight_decay=0,
                 target_overflow_rate=0, input_range=2, weight_range=2, bias_range=2, grad_range=2,
                 rng=None, grad_bits=None):
        limit = (6 / (in_units + units)) ** 0.5                                   # dfxp:338
        rng = rng if rng is not None else np.random.default_rng(0)
        self.name, self.use_bias = name, use_bias
        self.W = torch.from_numpy(rng.uniform(-limit, limit, size=[in_units, units]).astype(np.float32)).requires_grad_(True)
        self.bits, self.weight_decay = bits, weight_decay
        self.qX = Quantizer(ctx, bits, input_range, target_overflow_rate)        # dfxp:384 (bits, not bits+1)
        self.qW = Quantizer(ctx, bits, weight_range, target_overflow_rate)
        if use_bias:
            self.b = torch.zeros(units, requires_grad=True)
            self.qb = Quantizer(ctx, bits, bias_range, target_overflow_rate)
        self.qG = Quantizer(ctx, grad_bits or bits, grad_range, target_overflow_rate)

    def forward(self, X):
        self.X = _leaf(X)
        self.Xq = self.qX(self.X)
        self.Wq = self.qW(self.W)
        if self.qX.ctx.exact:
            self.y = (self.Xq.double() @ self.Wq.double()).float()
        else:
            self.y = self.Xq @ self.Wq                                            # dfxp:388
        if self.use_bias:
            self.bq = self.qb(self.b)
            self.y = self.y + self.bq
        return self.y.detach()

    def backward(self, grad, stochastic=True):
        self.gradq = self.qG(grad).detach()                                       # dfxp:453
        wrt = [self.X, self.W] + ([self.b] if self.use_bias else [])
        g = torch.autograd.grad(self.y, wrt, self.gradq)
        self.dW = g[1] + 2 * self.weight_decay * self.W.detach()                  # dfxp:457
        if self.use_bias:
            self.db = g[2]
        return g[0]

    def grads_and_vars(self):
        r = [(self.dW, self.W)]
        if self.use_bias:
            r.append((self.db, self.b))
        return r

    def variables(self):
        return [self.W] + ([self.b] if self.use_bias else [])

    def quantizers(self):
        return [self.qX, self.qW] + ([self.qb] if self.use_bias else []) + [self.qG]


class Sequential_q(Layer_q):
    """dfxp:512-536."""

    def __init__(self, *layers):
        self.layers = list(layers)

    def forward(self, X):
        for layer in self.layers:
            X = layer.forward(X)
        return X

    def backward(self, grad, stochastic=True):
        for layer in reversed(self.layers):
            grad = layer.backward(grad, stochastic)
        return grad

    def grads_and_vars(self):
        r = []
        for layer in self.layers:
            r += layer.grads_and_vars()
        return r

    def quantizers(self):
        r = []
        for layer in self.layers:
            r += layer.quantizers()
        return r

    def variables(self):
        r = []
        for layer in self.layers:
            r += layer.variables()
        return r


class Normalization_q(Layer_q):
    """dfxp:539-623: quantise, batch moments of the *quantised* x (biased var), normalise."""

    def __init__(self, ctx, name, bits, num_features, training=True, momentum=0.999, eps=1e-5,
                 target_overflow_rate=0, input_range=2, grad_range=2, grad_bits=None):
        self.name, self.train, self.eps, self.momentum, self.bits = name, training, eps, momentum, bits
        self.X_mean_running = torch.zeros(num_features)
        self.X_var_running = torch.ones(num_features)
        self.qX = Quantizer(ctx, bits, input_range, target_overflow_rate)        # dfxp:584
        self.qG = Quantizer(ctx, grad_bits or bits, grad_range, target_overflow_rate)  # dfxp:621

    def forward(self, X):
        self.X = _leaf(X)
        self.Xq = self.qX(self.X)
        axes = list(range(self.Xq.dim() - 1))
        if self.qX.ctx.exact:
            x64 = self.Xq.double()
            m64 = x64.mean(dim=axes)
            mean_b, var_b = m64.float(), ((x64 - m64) ** 2).mean(dim=axes).float()
        else:
            mean_b = self.Xq.mean(dim=axes)                                       # dfxp:588
            var_b = ((self.Xq - mean_b) ** 2).mean(dim=axes)
        if self.train:
            mean, var = mean_b, var_b
            with torch.no_grad():                                                 # dfxp:602-612
                self.X_mean_running = self.momentum * self.X_mean_running + (1 - self.momentum) * mean_b
                self.X_var_running = self.momentum * self.X_var_running + (1 - self.momentum) * var_b
        else:
            mean, var = self.X_mean_running, self.X_var_running
        self.mean, self.var = mean.detach(), var.detach()
        self.y = (self.Xq - mean) / ((var + self.eps) ** 0.5)                     # dfxp:616
        return self.y.detach()

    def backward(self, grad, stochastic=True):
        self.gradq = self.qG(grad).detach()
        if self.qX.ctx.exact and self.train:
            return self._backward_exact()
        return torch.autograd.grad(self.y, self.X, self.gradq)[0]                 # dfxp:623

    def _backward_exact(self):
        """The VJP of dfxp:616 through the batch mean and (biased) variance, tf.gradients' result in closed form:
        dx = (g - mean(g) - xhat * mean(g * xhat)) / sqrt(var + eps).  Exact mode: the two batch sums are accumulated
        exactly (fp64 holds sums of DFXP mantissa products) and rounded once, like every other accumulation of this mode;
        the elementwise part is plain fp32, one rounding per operation, in the order written here."""
        g, xq = self.gradq, self.Xq.detach()
        axes = list(range(g.dim() - 1))
        n = float(g.numel() // g.shape[-1])
        den = torch.sqrt(self.var + self.eps)
        sg = g.double().sum(dim=axes)
        sgx = (g.double() * xq.double()).sum(dim=axes)
        mg = (sg / n).float()
        mgx = (((sgx - self.mean.double() * sg) / den.double()) / n).float()      # mean(g * xhat)
        xhat = (xq - self.mean) / den
        return ((g - mg) - xhat * mgx) / den

    def quantizers(self):
        return [self.qX, self.qG]


class Rescale_q(Layer_q):
    """dfxp:626-694."""

    def __init__(self, ctx, name, bits, num_features, weight_decay=0, target_overflow_rate=0,
                 input_range=2, gamma_range=2, beta_range=2, grad_range=2, grad_bits=None):
        self.name, self.bits, self.weight_decay = name, bits, weight_decay
        self.gamma = torch.ones(num_features, requires_grad=True)
        self.beta = torch.zeros(num_features, requires_grad=True)
        self.qX = Quantizer(ctx, bits, input_range, target_overflow_rate)        # dfxp:677
        self.qg = Quantizer(ctx, bits, gamma_range, target_overflow_rate)        # dfxp:679
        self.qb = Quantizer(ctx, bits, beta_range, target_overflow_rate)         # dfxp:681
        self.qG = Quantizer(ctx, grad_bits or bits, grad_range, target_overflow_rate)  # dfxp:687

    def forward(self, X):
        self.X = _leaf(X)
        self.Xq = self.qX(self.X)
        self.gq = self.qg(self.gamma)
        self.bq = self.qb(self.beta)
        self.y = self.Xq * self.gq + self.bq                                      # dfxp:683
        return self.y.detach()

    def backward(self, grad, stochastic=True):
        self.gradq = self.qG(grad).detach()
        if self.qX.ctx.exact:    # the two reductions over (batch, pixels) accumulate exactly and round once
            axes = list(range(self.gradq.dim() - 1))
            g64 = self.gradq.double()
            dgamma = (g64 * self.Xq.detach().double()).sum(dim=axes).float()
            self.dgamma = dgamma + 2 * self.weight_decay * self.gamma.detach()    # dfxp:689
            self.dbeta = g64.sum(dim=axes).float()                                # dfxp:690
            return self.gradq * self.gq.detach()                                  # dfxp:691
        g = torch.autograd.grad(self.y, [self.X, self.gamma, self.beta], self.gradq)
        self.dgamma = g[1] + 2 * self.weight_decay * self.gamma.detach()          # dfxp:689
        self.dbeta = g[2]                                                         # dfxp:690
        return g[0]                                                               # dfxp:691

    def grads_and_vars(self):
        return [(self.dgamma, self.gamma), (self.dbeta, self.beta)]

    def variables(self):
        return [self.gamma, self.beta]

    def quantizers(self):
        return [self.qX, self.qg, self.qb, self.qG]


class BatchNorm_q(Sequential_q):
    """dfxp:697-743: Normalization_q then Rescale_q (whose input_range is hard-coded 2, dfxp:735)."""

    def __init__(self, ctx, name, bits, num_features, training=True, momentum=0.999, eps=1e-5,
                 weight_decay=0, target_overflow_rate=0, input_range=2, gamma_range=2, beta_range=2,
                 grad_range=2, grad_bits=None):
        super().__init__(
            Normalization_q(ctx, name + '-norm', bits, num_features, training, momentum, eps,
                            target_overflow_rate, input_range, grad_range, grad_bits=grad_bits),
            Rescale_q(ctx, name + '-rescale', bits, num_features, weight_decay, target_overflow_rate,
                      2, gamma_range, beta_range, grad_range, grad_bits=grad_bits))


class ReLU_q(Layer_q):
    """dfxp:983-990."""

    def forward(self, X):
        self.X = _leaf(X)
        # tf.maximum(0.0, X): TF's _MaximumGrad sends the gradient to X only where NOT (0.0 >= X), i.e.
        # strictly X > 0 (at X == 0 it goes to the constant).  torch.relu has exactly that backward;
        # torch.clamp_min would pass the gradient at X == 0, which quantised activations hit often.
        self.y = torch.relu(self.X)
        return self.y.detach()


class MaxPool_q(Layer_q):
    """dfxp:993-1006."""

    def __init__(self, ksize, strides, padding):
        self.ksize, self.strides, self.padding = ksize, strides, padding

    def forward(self, X):
        self.X = _leaf(X)
        self.y = tf_max_pool(self.X, self.ksize, self.strides, self.padding)
        return self.y.detach()


class AvgPool_q(Layer_q):
    """dfxp:1009-1022."""

    def __init__(self, ksize, strides, padding):
        self.ksize, self.strides, self.padding = ksize, strides, padding

    def forward(self, X):
        self.X = _leaf(X)
        self.y = tf_avg_pool(self.X, self.ksize, self.strides, self.padding)
        return self.y.detach()


class Dropout_q(Layer_q):
    """dfxp:1025-1040.  ``keep_prob`` is the KEEP probability; mask = floor(keep + u)."""

    def __init__(self, keep_prob, training=True, uniform_fn=None):
        self.keep_prob, self.train = keep_prob, training
        self.uniform_fn = uniform_fn or (lambda shape: torch.rand(shape))

    def forward(self, X):
        self.X = _leaf(X)
        # tf.nn.dropout returns x untouched when keep_prob is the constant 1 ("Do nothing if we know keep_prob == 1");
        # evaluating the formula instead would double an element whenever u = 1 - 2^-24 (1.0 + u rounds to 2.0 in fp32)
        if self.train and self.keep_prob != 1:
            u = self.uniform_fn(tuple(self.X.shape))
            self.mask = torch.floor(self.keep_prob + u)
            self.y = self.X / self.keep_prob * self.mask
        else:
            self.y = self.X
        return self.y.detach()


class GradientBuffer_q(Layer_q):
    """dfxp:473-509: quantise ``pad(grad) + buffer`` (always stochastic, :500), keep the residual (:503)."""

    def __init__(self, ctx, bits, shape, target_overflow_rate=0.0, grad_range=2):
        self.buffer = torch.zeros(*shape)
        self.qG = Quantizer(ctx, bits, grad_range, target_overflow_rate)

    def forward(self, X):
        self.X = X
        self.y = X
        return X

    def backward(self, grad, stochastic=True):
        pad = []
        for have, want in zip(reversed(grad.shape), reversed(self.buffer.shape)):              # :495-498 pads every dim at the end
            pad += [0, want - have]
        total = torch.nn.functional.pad(grad, pad) + self.buffer                               # :499
        gradq = self.qG(total).detach()                                                        # :500
        self.buffer = total - gradq                                                            # :503
        return gradq[:grad.shape[0]]                                                           # :506

    def quantizers(self):
        return [self.qG]


class Flatten_q(Layer_q):
    """dfxp:1043-1053."""

    def __init__(self, dim):
        self.dim = dim

    def forward(self, X):
        self.X = _leaf(X)
        self.y = self.X.reshape(-1, self.dim)
        return self.y.detach()


class ResidualBlock_q(Layer_q):
    """dfxp:746-875."""
    expansion = 1

    def __init__(self, ctx, name, bits, in_channels, channels, stride, training=True, batch_norm=True,
                 weight_decay=0, target_overflow_rate=0, input_range=2, weight_range=2, bias_range=2,
                 grad_range=2, rng=None, grad_bits=None):
        kw = dict(use_bias=not batch_norm, weight_decay=weight_decay, input_range=input_range,
                  weight_range=weight_range, bias_range=bias_range, grad_range=grad_range, rng=rng,
                  grad_bits=grad_bits)
        bn = lambda n, c: (BatchNorm_q(ctx, n, bits, c, training, weight_decay=weight_decay,
                                       target_overflow_rate=target_overflow_rate, input_range=input_range,
                                       grad_range=grad_range, grad_bits=grad_bits)
                           if batch_norm else Layer_q())
        self.residual = Sequential_q(
            Conv2d_q(ctx, name + '-1', bits, [3, 3, in_channels, channels], [1, stride, stride, 1], 'SAME', **kw),
            bn(name + '-bn1', channels),
            ReLU_q(),
            Conv2d_q(ctx, name + '-2', bits, [3, 3, channels, channels], [1, 1, 1, 1], 'SAME', **kw),
            bn(name + '-bn2', channels))
        self._build_shortcut(ctx, name, bits, in_channels, channels, stride, bn, kw, target_overflow_rate)
        self.relu = ReLU_q()

    def _build_shortcut(self, ctx, name, bits, in_channels, channels, stride, bn, kw, t):
        if stride == 1 and in_channels == self.expansion * channels:              # dfxp:828-829
            self.shortcut = Sequential_q()
        else:
            self.shortcut = Sequential_q(
                Conv2d_q(ctx, name + '-shortcut', bits, [1, 1, in_channels, self.expansion * channels],
                         [1, stride, stride, 1], 'SAME', target_overflow_rate=t, **kw),
                bn(name + '-shortcut-bn', self.expansion * channels))

    def forward(self, X):
        self.y1 = self.residual.forward(X)                                        # dfxp:860
        self.y2 = self.shortcut.forward(X)                                        # dfxp:861
        return self.relu.forward(self.y1 + self.y2)                               # dfxp:862

    def backward(self, grad, stochastic=True):
        grad = self.relu.backward(grad, stochastic)                               # dfxp:866
        grad1 = self.residual.backward(grad, stochastic)
        grad2 = self.shortcut.backward(grad, stochastic)
        return grad1 + grad2                                                      # dfxp:869

    def grads_and_vars(self):
        return self.residual.grads_and_vars() + self.shortcut.grads_and_vars()

    def variables(self):
        return self.residual.variables() + self.shortcut.variables()

    def quantizers(self):
        return self.residual.quantizers() + self.shortcut.quantizers()


class ResidualBottleneck_q(ResidualBlock_q):
    """dfxp:878-980 — 1x1, 3x3 (carries the stride, dfxp:929-934), 1x1 x4."""
    expansion = 4

    def __init__(self, ctx, name, bits, in_channels, channels, stride, training=True, batch_norm=True,
                 weight_decay=0, target_overflow_rate=0, input_range=2, weight_range=2, bias_range=2,
                 grad_range=2, rng=None, grad_bits=None):
        kw = dict(use_bias=not batch_norm, weight_decay=weight_decay, input_range=input_range,
                  weight_range=weight_range, bias_range=bias_range, grad_range=grad_range, rng=rng,
                  grad_bits=grad_bits)
        bn = lambda n, c: (BatchNorm_q(ctx, n, bits, c, training, weight_decay=weight_decay,
                                       target_overflow_rate=target_overflow_rate, input_range=input_range,
                                       grad_range=grad_range, grad_bits=grad_bits)
                           if batch_norm else Layer_q())
        out_channels = 4 * channels
        self.residual = Sequential_q(
            Conv2d_q(ctx, name + '-1', bits, [1, 1, in_channels, channels], [1, 1, 1, 1], 'SAME', **kw),
            bn(name + '-bn1', channels), ReLU_q(),
            Conv2d_q(ctx, name + '-2', bits, [3, 3, channels, channels], [1, stride, stride, 1], 'SAME', **kw),
            bn(name + '-bn2', channels), ReLU_q(),
            Conv2d_q(ctx, name + '-3', bits, [1, 1, channels, out_channels], [1, 1, 1, 1], 'SAME', **kw),
            bn(name + '-bn3', out_channels))
        self._build_shortcut(ctx, name, bits, in_channels, channels, stride, bn, kw, target_overflow_rate)
        self.relu = ReLU_q()


# --------------------------------------------------------------------------------------------
# Models (models.py) and the training step (trainer.py:79-84, 157-160)
# --------------------------------------------------------------------------------------------


class Model:
    """models.py:7-54: forward chain, mean sparse-softmax-xent loss, manual reverse chain."""

    def __init__(self, bits, dropout=0.5, weight_decay=0.0, noise=None, seed=0, grad_bits=None, exact=False):
        self.bits, self.dropout, self.weight_decay, self.grad_bits = bits, dropout, weight_decay, grad_bits
        self.ctx = Context(noise, exact=exact)
        self.rng = np.random.default_rng(seed)
        self.layers = self.get_layers()
        self.velocity = None

    def get_layers(self):
        return []

    def forward(self, X):
        for layer in self.layers:                                                 # models.py:22-24
            X = layer.forward(X)
        self.logits = X
        return X

    def loss_and_grad(self, labels):
        logits = self.logits.detach().requires_grad_(True)
        loss = F.cross_entropy(logits, labels.long(), reduction='mean')           # models.py:30-32
        g = torch.autograd.grad(loss, logits)[0]
        return loss.detach(), g

    def backward(self, labels):
        self.loss, grad = self.loss_and_grad(labels)
        for layer in reversed(self.layers):                                       # models.py:47-51
            grad = layer.backward(grad, True)
        return grad

    def grads_and_vars(self):
        r = []
        for layer in self.layers:
            r += layer.grads_and_vars()
        return r

    def quantizers(self):
        r = []
        for layer in self.layers:
            r += layer.quantizers()
        return r

    def variables(self):
        r = []
        for layer in self.layers:
            r += layer.variables()
        return r

    def ranges(self):
        return [int(q.range) for q in self.quantizers()]

    def train_step(self, X, labels, lr=1e-2, momentum=0.9):
        """One ``sess.run([train_op, update_range_op])`` (trainer.py:157): returns the loss."""
        self.forward(X)
        self.backward(labels)
        gv = self.grads_and_vars()
        if self.velocity is None:
            self.velocity = [torch.zeros_like(v) for _, v in gv]
        with torch.no_grad():
            for (g, v), a in zip(gv, self.velocity):                              # trainer.py:81-82
                a.mul_(momentum).add_(g)
                v.sub_(lr * a)
        if hasattr(self.ctx.noise, 'step'):
            self.ctx.noise.step += 1
        return float(self.loss)


class CIFAR10_Model(Model):
    """models.py:155-234."""

    def get_layers(self):
        c, b, wd, r, gb = self.ctx, self.bits, self.weight_decay, self.rng, self.grad_bits
        pool = lambda: MaxPool_q([1, 3, 3, 1], [1, 2, 2, 1], 'SAME')
        drop = lambda: Dropout_q(self.dropout, True, getattr(self, 'dropout_uniform', None))
        return [
            Conv2d_q(c, 'conv1', b, [5, 5, 3, 64], [1, 1, 1, 1], 'SAME', weight_decay=wd, rng=r, grad_bits=gb),
            ReLU_q(), pool(),
            drop(),
            Conv2d_q(c, 'conv2', b, [5, 5, 64, 128], [1, 1, 1, 1], 'SAME', weight_decay=wd, rng=r, grad_bits=gb),
            ReLU_q(), pool(),
            drop(),
            Conv2d_q(c, 'conv3', b, [5, 5, 128, 128], [1, 1, 1, 1], 'SAME', weight_decay=wd, rng=r, grad_bits=gb),
            ReLU_q(), pool(),
            Flatten_q(128 * 4 * 4),
            drop(),
            Dense_q(c, 'dense1', b, 128 * 4 * 4, 400, weight_decay=wd, rng=r, grad_bits=gb),
            ReLU_q(),
            drop(),
            Dense_q(c, 'softmax', b, 400, 10, weight_decay=wd, rng=r, grad_bits=gb),
        ]


class PI_MNIST_Model(Model):
    """models.py:57-88: 784 -> 1024 -> 1024 -> 10 dense network (permutation-invariant MNIST)."""

    def get_layers(self):
        c, b, wd, r, gb = self.ctx, self.bits, self.weight_decay, self.rng, self.grad_bits
        drop = lambda: Dropout_q(self.dropout, True, getattr(self, 'dropout_uniform', None))
        return [
            Dense_q(c, 'dense1', b, 784, 1024, weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(), drop(),
            Dense_q(c, 'dense2', b, 1024, 1024, weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(), drop(),
            Dense_q(c, 'softmax', b, 1024, 10, weight_decay=wd, rng=r, grad_bits=gb),
        ]


class MNIST_Model(Model):
    """models.py:91-152: LeNet-5 (5x5 SAME conv 1->6, 2x2 VALID pool, 5x5 VALID 6->16, pool, 5x5 VALID 16->120, 84, 10)."""

    def get_layers(self):
        c, b, wd, r, gb = self.ctx, self.bits, self.weight_decay, self.rng, self.grad_bits
        pool = lambda: MaxPool_q([1, 2, 2, 1], [1, 2, 2, 1], 'VALID')
        drop = lambda: Dropout_q(self.dropout, True, getattr(self, 'dropout_uniform', None))
        return [
            Conv2d_q(c, 'conv1', b, [5, 5, 1, 6], [1, 1, 1, 1], 'SAME', weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(), pool(),
            Conv2d_q(c, 'conv2', b, [5, 5, 6, 16], [1, 1, 1, 1], 'VALID', weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(), pool(),
            Conv2d_q(c, 'conv3', b, [5, 5, 16, 120], [1, 1, 1, 1], 'VALID', weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(),
            Flatten_q(120), drop(),
            Dense_q(c, 'dense1', b, 120, 84, weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(), drop(),
            Dense_q(c, 'softmax', b, 84, 10, weight_decay=wd, rng=r, grad_bits=gb),
        ]


class CIFAR10_VGG_Model(Model):
    """models.py:237-368: 2x(3x3 128) pool 2x(3x3 256) pool 2x(3x3 512) pool, 1024, 1024, 10."""

    def get_layers(self):
        c, b, wd, r, gb = self.ctx, self.bits, self.weight_decay, self.rng, self.grad_bits
        pool = lambda: MaxPool_q([1, 3, 3, 1], [1, 2, 2, 1], 'SAME')
        drop = lambda: Dropout_q(self.dropout, True, getattr(self, 'dropout_uniform', None))
        conv = lambda n, ci, co: Conv2d_q(c, n, b, [3, 3, ci, co], [1, 1, 1, 1], 'SAME', weight_decay=wd, rng=r, grad_bits=gb)
        return [
            conv('conv1-1', 3, 128), ReLU_q(), conv('conv1-2', 128, 128), ReLU_q(), pool(),
            drop(), conv('conv2-1', 128, 256), ReLU_q(), conv('conv2-2', 256, 256), ReLU_q(), pool(),
            drop(), conv('conv3-1', 256, 512), ReLU_q(), conv('conv3-2', 512, 512), ReLU_q(), pool(),
            Flatten_q(512 * 4 * 4),
            drop(), Dense_q(c, 'dense1', b, 512 * 4 * 4, 1024, weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(),
            drop(), Dense_q(c, 'dense2', b, 1024, 1024, weight_decay=wd, rng=r, grad_bits=gb), ReLU_q(),
            drop(), Dense_q(c, 'softmax', b, 1024, 10, weight_decay=wd, rng=r, grad_bits=gb),
        ]


class CIFAR10_Resnet(Model):
    """models.py:371-450."""

    def __init__(self, bits, num_blocks, block=ResidualBlock_q, **kw):
        self.num_blocks, self.block = num_blocks, block
        super().__init__(bits, **kw)

    def _build_blocks(self, channels, num_blocks, stride):
        blocks = []
        for i in range(1, 1 + num_blocks):
            blocks.append(self.block(self.ctx, 'block%d-%d' % (channels, i), self.bits, self.channels, channels,
                                     1 if i > 1 else stride, True, weight_decay=self.weight_decay,
                                     rng=self.rng, grad_bits=self.grad_bits))
            self.channels = channels * self.block.expansion
        return blocks

    def get_layers(self):
        self.channels = 16
        c, b, wd = self.ctx, self.bits, self.weight_decay
        return [
            Conv2d_pq(c, 'conv1', b, [3, 3, 3, 16], [1, 1, 1, 1], 'SAME', use_bias=False, weight_decay=wd,
                      rng=self.rng, grad_bits=self.grad_bits),
            BatchNorm_q(c, 'conv1-bn', b, 16, True, weight_decay=wd, grad_bits=self.grad_bits),
            ReLU_q(),
        ] + self._build_blocks(16, self.num_blocks[0], 1) \
          + self._build_blocks(32, self.num_blocks[1], 2) \
          + self._build_blocks(64, self.num_blocks[2], 2) + [
            AvgPool_q([1, 8, 8, 1], [1, 1, 1, 1], 'VALID'),
            Flatten_q(64),
            Dense_q(c, 'softmax', b, 64, 10, use_bias=False, weight_decay=wd, rng=self.rng,
                    grad_bits=self.grad_bits),
        ]


def CIFAR10_Resnet20(bits, **kw):
    """models.py:453-455."""
    return CIFAR10_Resnet(bits, [3, 3, 3], ResidualBlock_q, **kw)


class ImageNet_Resnet(Model):
    """Composed from the reference's blocks: stem 7x7/2 + BN + ReLU + MaxPool 3x3/2,
    four stages, global AvgPool, Dense -> num_classes.  Not in models.py; whole-model parity is vs
    this composition."""

    def __init__(self, bits, num_blocks, block, image=224, num_classes=1000, **kw):
        self.num_blocks, self.block, self.image, self.num_classes = num_blocks, block, image, num_classes
        super().__init__(bits, **kw)

    _build_blocks = CIFAR10_Resnet._build_blocks

    def get_layers(self):
        self.channels = 64
        c, b, wd = self.ctx, self.bits, self.weight_decay
        final = -(-self.image // 32)
        layers = [
            Conv2d_q(c, 'conv1', b, [7, 7, 3, 64], [1, 2, 2, 1], 'SAME', use_bias=False, weight_decay=wd,
                     rng=self.rng, grad_bits=self.grad_bits),
            BatchNorm_q(c, 'conv1-bn', b, 64, True, weight_decay=wd, grad_bits=self.grad_bits),
            ReLU_q(),
            MaxPool_q([1, 3, 3, 1], [1, 2, 2, 1], 'SAME'),
        ]
        for ch, n, s in zip((64, 128, 256, 512), self.num_blocks, (1, 2, 2, 2)):
            layers += self._build_blocks(ch, n, s)
        layers += [
            AvgPool_q([1, final, final, 1], [1, 1, 1, 1], 'VALID'),
            Flatten_q(self.channels),
            Dense_q(c, 'softmax', b, self.channels, self.num_classes, use_bias=False, weight_decay=wd,
                    rng=self.rng, grad_bits=self.grad_bits),
        ]
        return layers


def Resnet18(bits, **kw):
    return ImageNet_Resnet(bits, [2, 2, 2, 2], ResidualBlock_q, **kw)


def Resnet50(bits, **kw):
    return ImageNet_Resnet(bits, [3, 4, 6, 3], ResidualBottleneck_q, **kw)

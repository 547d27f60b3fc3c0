// Fused dynamic-fixed-point quantiser for B200 (sm_100a): quantise + overflow statistics + range
// controller in ONE pass over the tensor.  Replaces weight_quantization / overflow_rate /
// update_range of /root/reference/dynamic_fixed_point.py:4-94 (≈25 TF elementwise + reduce launches,
// ≈100 B/elem of HBM traffic) with 4 B/elem read + (4 | 1 | 2) B/elem written.
//
// HBM-bound by construction.  Layout: the tensor is [n_outer, n_inner]; a CTA tile is 256 float4
// columns x `rows_per_group` rows, so one noise vector (loaded or Philox-generated once per thread)
// is reused down the rows exactly as tf.random_uniform(X.shape[1:]) broadcasts over dim 0.  The
// grid is persistent (a multiple of the SM count) so the statistics cost two 64-bit atomics per CTA,
// not per tile; the last CTA (atomic ticket) applies the controller on the device.
#include "common.cuh"

namespace lbt {

namespace {

constexpr int kThreads = 256;
constexpr int kRowsUnroll = 8;

// Defaults from the round-1 sweep on B200 (benchmarks/quantize_bench.py --tune, 2^26 elements).
int g_blocks_per_sm = 8;
int g_rows_per_group = 32;

struct QParams {
  const float* x;
  size_t n_outer, n_inner;
  int bits;
  int32_t* ib;
  float target;
  const float* noise;
  uint64_t seed, offset;
  const uint64_t* dev_step;
  float* out;
  void* mant;
  int mant_kind;
  unsigned long long* counters;
  int update_range;
  // tiling (vector path)
  uint32_t n_vec, chunks, rows_per_group;
  uint64_t total_tiles;
};

struct QConst {
  float m, inv_m, L, hi, half;
};

__device__ __forceinline__ QConst make_const(int bits, int ib) {
  QConst c;
  int f = bits - ib - 1;
  f = max(-126, min(126, f));
  c.m = exp2i(f);
  c.inv_m = exp2i(-f);
  c.L = exp2i(bits - 1);
  c.hi = c.L - 1.0f;
  c.half = c.L * 0.5f;
  return c;
}

template <int MODE>
__device__ __forceinline__ float quant1(float x, float u, const QConst& c, uint32_t& n1, uint32_t& n2) {
  const float y = __fmul_rn(x, c.m);  // dfxp:29/36/62  X * multiplier
  n1 += (uint32_t)(y >= c.L) + (uint32_t)(y < -c.L);
  n2 += (uint32_t)(y >= c.half) + (uint32_t)(y < -c.half);
  float k;
  if (MODE == LBT_ROUND_NEAREST) {
    k = rintf(fminf(fmaxf(y, -c.L), c.hi));  // clip then round-half-even
  } else {
    const float t = __fadd_rn(y, u);  // noise before the clip; separate rounding (no FMA contraction)
    k = floorf(fminf(fmaxf(t, -c.L), c.hi));
  }
  return k;
}

// Same arithmetic with min/max tracking instead of the four comparisons + adds of the exact counters
// (LBT_STATS_MINMAX): enough for the controller whenever target_overflow_rate == 0, the only value the
// reference ever uses (n1 > 0 <=> max >= L or min < -L; n2 == 0 <=> max < L/2 and min >= -L/2).
template <int MODE>
__device__ __forceinline__ float quant1_mm(float x, float u, const QConst& c, float& mx, float& mn) {
  const float y = __fmul_rn(x, c.m);
  mx = fmaxf(mx, y);
  mn = fminf(mn, y);
  if (MODE == LBT_ROUND_NEAREST) return rintf(fminf(fmaxf(y, -c.L), c.hi));
  return floorf(fminf(fmaxf(__fadd_rn(y, u), -c.L), c.hi));
}

// Tail of every quantiser kernel: block-reduce the two counters, publish, ticket, controller.
__device__ __forceinline__ void finish_stats(uint32_t n1, uint32_t n2, const QParams& p, int ib_at_launch) {
  if (p.counters == nullptr) return;
  __shared__ uint32_t s1[kThreads / 32], s2[kThreads / 32];
  n1 = warp_sum(n1);
  n2 = warp_sum(n2);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    s1[w] = n1;
    s2[w] = n2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t b1 = 0, b2 = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) {
      b1 += s1[i];
      b2 += s2[i];
    }
    unsigned long long* c = p.counters;
    if (b1) atomicAdd(c + LBT_CNT_OVER, (unsigned long long)b1);
    if (b2) atomicAdd(c + LBT_CNT_OVER_HALF, (unsigned long long)b2);
    __threadfence();
    const unsigned long long t = atomicAdd(c + LBT_CNT_TICKET, 1ull);
    if (t == (unsigned long long)gridDim.x - 1ull) {
      __threadfence();
      const unsigned long long numel =
          atomicAdd(c + LBT_CNT_NUMEL, 0ull) + (unsigned long long)p.n_outer * (unsigned long long)p.n_inner;
      if (p.update_range) {
        const unsigned long long c1 = atomicAdd(c + LBT_CNT_OVER, 0ull);
        const unsigned long long c2 = atomicAdd(c + LBT_CNT_OVER_HALF, 0ull);
        const float n = (float)(numel ? numel : 1ull);
        const float r1 = __fdiv_rn((float)c1, n), r2 = __fdiv_rn((float)c2, n);
        const int delta = (r1 > p.target) ? 1 : ((r2 <= p.target) ? -1 : 0);  // dfxp:84-92
        *p.ib = min(p.bits - 1, ib_at_launch + delta);                         // dfxp:94
        c[LBT_CNT_OVER] = 0ull;
        c[LBT_CNT_OVER_HALF] = 0ull;
        c[LBT_CNT_NUMEL] = 0ull;
      } else {
        c[LBT_CNT_NUMEL] = numel;
      }
      c[LBT_CNT_TICKET] = 0ull;
    }
  }
}

__device__ __forceinline__ void store_mant4(void* mant, int kind, size_t idx, float k0, float k1, float k2, float k3) {
  const int i0 = __float2int_rn(k0), i1 = __float2int_rn(k1), i2 = __float2int_rn(k2), i3 = __float2int_rn(k3);
  if (kind == LBT_MANT_S16) {
    uint2 v;
    v.x = (uint32_t)(i0 & 0xffff) | ((uint32_t)(i1 & 0xffff) << 16);
    v.y = (uint32_t)(i2 & 0xffff) | ((uint32_t)(i3 & 0xffff) << 16);
    *reinterpret_cast<uint2*>(reinterpret_cast<int16_t*>(mant) + idx) = v;
  } else {  // S8 / U8: same low byte
    const uint32_t v = (uint32_t)(i0 & 0xff) | ((uint32_t)(i1 & 0xff) << 8) | ((uint32_t)(i2 & 0xff) << 16) |
                       ((uint32_t)(i3 & 0xff) << 24);
    *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(mant) + idx) = v;
  }
}

template <int MODE, bool MM>
__global__ void __launch_bounds__(kThreads) quantize_vec_kernel(const QParams p) {
  pdl_trigger();
  pdl_wait();
  const int ib = *reinterpret_cast<volatile const int32_t*>(p.ib);
  const QConst c = make_const(p.bits, ib);
  uint64_t off = p.offset;
  if (MODE == LBT_ROUND_STOCHASTIC_PHILOX && p.dev_step) off += (*p.dev_step) << 32;
  uint32_t n1 = 0, n2 = 0;
  float mx = -INFINITY, mn = INFINITY;

  for (uint64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const uint32_t rg = (uint32_t)(tile / p.chunks), ch = (uint32_t)(tile % p.chunks);
    const uint32_t v = ch * kThreads + threadIdx.x;
    if (v >= p.n_vec) continue;
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == LBT_ROUND_STOCHASTIC_NOISE) u = __ldg(reinterpret_cast<const float4*>(p.noise) + v);
    if (MODE == LBT_ROUND_STOCHASTIC_PHILOX) u = philox_noise4(v, p.seed, off);
    const size_t r0 = (size_t)rg * p.rows_per_group;
    const size_t r1 = min(r0 + (size_t)p.rows_per_group, p.n_outer);
    for (size_t r = r0; r < r1; r += kRowsUnroll) {
      float4 xv[kRowsUnroll];
#pragma unroll
      for (int i = 0; i < kRowsUnroll; ++i)
        if (r + i < r1) xv[i] = __ldcs(reinterpret_cast<const float4*>(p.x + (r + i) * p.n_inner) + v);
#pragma unroll
      for (int i = 0; i < kRowsUnroll; ++i) {
        if (r + i < r1) {
          float k0, k1, k2, k3;
          if (MM) {
            k0 = quant1_mm<MODE>(xv[i].x, u.x, c, mx, mn);
            k1 = quant1_mm<MODE>(xv[i].y, u.y, c, mx, mn);
            k2 = quant1_mm<MODE>(xv[i].z, u.z, c, mx, mn);
            k3 = quant1_mm<MODE>(xv[i].w, u.w, c, mx, mn);
          } else {
            k0 = quant1<MODE>(xv[i].x, u.x, c, n1, n2);
            k1 = quant1<MODE>(xv[i].y, u.y, c, n1, n2);
            k2 = quant1<MODE>(xv[i].z, u.z, c, n1, n2);
            k3 = quant1<MODE>(xv[i].w, u.w, c, n1, n2);
          }
          const size_t idx = (r + i) * p.n_inner + 4 * (size_t)v;
          if (p.out)
            *reinterpret_cast<float4*>(p.out + idx) =
                make_float4(k0 * c.inv_m, k1 * c.inv_m, k2 * c.inv_m, k3 * c.inv_m);
          if (p.mant) store_mant4(p.mant, p.mant_kind, idx, k0, k1, k2, k3);
        }
      }
    }
  }
  if (MM) {  // per-thread indicators: the controller only needs "any" / "none"
    n1 = (mx >= c.L || mn < -c.L) ? 1u : 0u;
    n2 = (mx >= c.half || mn < -c.half) ? 1u : 0u;
  }
  finish_stats(n1, n2, p, ib);
}

// Any shape / alignment: one element per thread-iteration, identical arithmetic.
template <int MODE>
__global__ void __launch_bounds__(kThreads) quantize_scalar_kernel(const QParams p) {
  const int ib = *reinterpret_cast<volatile const int32_t*>(p.ib);
  const QConst c = make_const(p.bits, ib);
  uint64_t off = p.offset;
  if (MODE == LBT_ROUND_STOCHASTIC_PHILOX && p.dev_step) off += (*p.dev_step) << 32;
  uint32_t n1 = 0, n2 = 0;
  const size_t n = p.n_outer * p.n_inner;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    const size_t col = i % p.n_inner;
    float u = 0.f;
    if (MODE == LBT_ROUND_STOCHASTIC_NOISE) u = __ldg(p.noise + col);
    if (MODE == LBT_ROUND_STOCHASTIC_PHILOX) {
      const float4 u4 = philox_noise4(col >> 2, p.seed, off);
      const int l = (int)(col & 3);
      u = l == 0 ? u4.x : (l == 1 ? u4.y : (l == 2 ? u4.z : u4.w));
    }
    const float k = quant1<MODE>(p.x[i], u, c, n1, n2);
    if (p.out) p.out[i] = k * c.inv_m;
    if (p.mant) {
      const int ki = __float2int_rn(k);
      if (p.mant_kind == LBT_MANT_S16)
        reinterpret_cast<int16_t*>(p.mant)[i] = (int16_t)ki;
      else
        reinterpret_cast<uint8_t*>(p.mant)[i] = (uint8_t)(ki & 0xff);
    }
  }
  finish_stats(n1, n2, p, ib);
}

// 3-channel input (an image): one thread per pixel, output 16 s8 bytes {hi x3, hi x3, lo x3, 0 x7} (LBT_MANT_S9C3).
template <int MODE>
__global__ void __launch_bounds__(kThreads) quantize_c3_kernel(const QParams p) {
  const int ib = *reinterpret_cast<volatile const int32_t*>(p.ib);
  const QConst c = make_const(p.bits, ib);
  uint64_t off = p.offset;
  if (MODE == LBT_ROUND_STOCHASTIC_PHILOX && p.dev_step) off += (*p.dev_step) << 32;
  uint32_t n1 = 0, n2 = 0;
  const size_t ppr = p.n_inner / 3, npix = p.n_outer * ppr;
  for (size_t pix = (size_t)blockIdx.x * kThreads + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * kThreads) {
    const size_t row = pix / ppr, j0 = (pix % ppr) * 3;
    int k[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const size_t col = j0 + ch;
      float u = 0.f;
      if (MODE == LBT_ROUND_STOCHASTIC_NOISE) u = __ldg(p.noise + col);
      if (MODE == LBT_ROUND_STOCHASTIC_PHILOX) {
        const float4 u4 = philox_noise4(col >> 2, p.seed, off);
        const int l = (int)(col & 3);
        u = l == 0 ? u4.x : (l == 1 ? u4.y : (l == 2 ? u4.z : u4.w));
      }
      k[ch] = __float2int_rn(quant1<MODE>(p.x[row * p.n_inner + col], u, c, n1, n2));
    }
    uint32_t w[4] = {0, 0, 0, 0};
    uint8_t b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) b[i] = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      b[ch] = b[3 + ch] = (uint8_t)((k[ch] >> 1) & 0xff);  // hi = floor(k / 2), in [-128, 127]
      b[6 + ch] = (uint8_t)(k[ch] & 1);                     // lo in {0, 1}
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i >> 2] |= (uint32_t)b[i] << (8 * (i & 3));
    reinterpret_cast<uint4*>(p.mant)[pix] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  finish_stats(n1, n2, p, ib);
}

// The same for n_inner % 12 == 0 and 16-byte aligned tensors (every image batch): one thread owns FOUR pixels = 12 consecutive
// inner elements = three float4 loads and three Philox groups, and walks the batch rows of its row group with that noise in
// registers (the noise is broadcast over dim 0) — the per-pixel kernel above ran three Philox blocks and a 64-bit division per
// pixel (108 instructions per element, ncu) and read 4 bytes per load.  Same quant1 arithmetic: bit-identical output.
template <int MODE>
__global__ void __launch_bounds__(kThreads) quantize_c3v_kernel(const QParams p) {
  pdl_trigger();
  pdl_wait();
  const int ib = *reinterpret_cast<volatile const int32_t*>(p.ib);
  const QConst c = make_const(p.bits, ib);
  uint64_t off = p.offset;
  if (MODE == LBT_ROUND_STOCHASTIC_PHILOX && p.dev_step) off += (*p.dev_step) << 32;
  uint32_t n1 = 0, n2 = 0;
  const uint32_t nv4 = (uint32_t)(p.n_inner / 4);   // float4 per row
  for (uint64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const uint32_t rg = (uint32_t)(tile / p.chunks), v = (uint32_t)(tile % p.chunks) * kThreads + threadIdx.x;
    if (v >= p.n_vec) continue;   // n_vec = n_inner / 12
    float u[12];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == LBT_ROUND_STOCHASTIC_NOISE) t = __ldg(reinterpret_cast<const float4*>(p.noise) + 3 * (size_t)v + g);
      if (MODE == LBT_ROUND_STOCHASTIC_PHILOX) t = philox_noise4(3ull * v + g, p.seed, off);
      u[4 * g] = t.x; u[4 * g + 1] = t.y; u[4 * g + 2] = t.z; u[4 * g + 3] = t.w;
    }
    const uint32_t r0 = rg * p.rows_per_group, r1 = min(r0 + p.rows_per_group, (uint32_t)p.n_outer);
    for (uint32_t r = r0; r < r1; ++r) {
      const float4* xr = reinterpret_cast<const float4*>(p.x) + ((size_t)r * nv4 + 3 * (size_t)v);
      const float4 a = __ldcs(xr), b = __ldcs(xr + 1), d = __ldcs(xr + 2);
      const float xs[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
      int k[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) k[i] = __float2int_rn(quant1<MODE>(xs[i], u[i], c, n1, n2));
      uint4* o = reinterpret_cast<uint4*>(p.mant) + ((size_t)r * (nv4 / 3) * 4 + 4 * (size_t)v);   // pixel index = r * n_inner/3 + 4v
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t h0 = (uint32_t)(k[3 * q] >> 1) & 0xffu, h1 = (uint32_t)(k[3 * q + 1] >> 1) & 0xffu,
                       h2 = (uint32_t)(k[3 * q + 2] >> 1) & 0xffu;   // hi = floor(k / 2), in [-128, 127]
        const uint32_t l0 = (uint32_t)k[3 * q] & 1u, l1 = (uint32_t)k[3 * q + 1] & 1u, l2 = (uint32_t)k[3 * q + 2] & 1u;
        o[q] = make_uint4(h0 | (h1 << 8) | (h2 << 16) | (h0 << 24), h1 | (h2 << 8) | (l0 << 16) | (l1 << 24), l2, 0u);
      }
    }
  }
  finish_stats(n1, n2, p, ib);
}

// GradientBuffer_q (dynamic_fixed_point.py:473-509), one pass: total = pad(grad) + buffer; q = Q_stochastic(total);
// buffer <- total - q; out = q[:n_grad_rows].  One element per thread-iteration (the layer is disabled in the reference's
// models: correctness and a single pass matter here, not the last 20 % of bandwidth).
template <int MODE>
__global__ void __launch_bounds__(kThreads) quantize_residual_kernel(const QParams p, float* __restrict__ buffer, size_t n_grad_rows) {
  const int ib = *reinterpret_cast<volatile const int32_t*>(p.ib);
  const QConst c = make_const(p.bits, ib);
  uint64_t off = p.offset;
  if (MODE == LBT_ROUND_STOCHASTIC_PHILOX && p.dev_step) off += (*p.dev_step) << 32;
  uint32_t n1 = 0, n2 = 0;
  const size_t n = p.n_outer * p.n_inner, n_grad = n_grad_rows * p.n_inner;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
    const size_t col = i % p.n_inner;
    float u = 0.f;
    if (MODE == LBT_ROUND_STOCHASTIC_NOISE) u = __ldg(p.noise + col);
    if (MODE == LBT_ROUND_STOCHASTIC_PHILOX) {
      const float4 u4 = philox_noise4(col >> 2, p.seed, off);
      const int l = (int)(col & 3);
      u = l == 0 ? u4.x : (l == 1 ? u4.y : (l == 2 ? u4.z : u4.w));
    }
    const float total = __fadd_rn(i < n_grad ? p.x[i] : 0.f, buffer[i]);       // dfxp:499 tf.pad(grad) + buffer
    const float q = quant1<MODE>(total, u, c, n1, n2) * c.inv_m;              // dfxp:500
    buffer[i] = __fsub_rn(total, q);                                           // dfxp:503
    if (i < n_grad) p.out[i] = q;                                              // dfxp:506
  }
  finish_stats(n1, n2, p, ib);
}

__global__ void noise_fill_kernel(float* u, size_t n_inner, uint64_t seed, uint64_t offset, const uint64_t* dev_step) {
  uint64_t off = offset;
  if (dev_step) off += (*dev_step) << 32;
  const size_t ng = (n_inner + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (size_t)gridDim.x * blockDim.x) {
    const float4 v = philox_noise4(g, seed, off);
    const size_t j = 4 * g;
    if (j + 0 < n_inner) u[j + 0] = v.x;
    if (j + 1 < n_inner) u[j + 1] = v.y;
    if (j + 2 < n_inner) u[j + 2] = v.z;
    if (j + 3 < n_inner) u[j + 3] = v.w;
  }
}

__global__ void update_ranges_kernel(int32_t* ranges, unsigned long long* counters, const int32_t* bits,
                                     const float* target, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long* c = counters + i * LBT_CNT_WORDS;
  const unsigned long long numel = c[LBT_CNT_NUMEL];
  if (numel == 0ull) return;  // quantiser did not run this step: leave its range alone
  const float nn = (float)numel;
  const float r1 = __fdiv_rn((float)c[LBT_CNT_OVER], nn), r2 = __fdiv_rn((float)c[LBT_CNT_OVER_HALF], nn);
  const float t = target ? target[i] : 0.f;
  const int delta = (r1 > t) ? 1 : ((r2 <= t) ? -1 : 0);
  ranges[i] = min(bits[i] - 1, ranges[i] + delta);
  c[LBT_CNT_OVER] = 0ull;
  c[LBT_CNT_OVER_HALF] = 0ull;
  c[LBT_CNT_NUMEL] = 0ull;
  c[LBT_CNT_TICKET] = 0ull;
}

__global__ void step_advance_kernel(uint64_t* s) { *s += 1; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

}  // namespace lbt

using namespace lbt;

extern "C" int lbt_quantize_tune(int blocks_per_sm, int rows_per_group) {
  if (blocks_per_sm > 0) g_blocks_per_sm = blocks_per_sm;
  if (rows_per_group > 0) g_rows_per_group = rows_per_group;
  return LBT_OK;
}

extern "C" int lbt_quantize(const float* x, size_t n_outer, size_t n_inner, int bits, int32_t* integer_bits,
                            float target_overflow_rate, int mode, const float* noise, uint64_t seed, uint64_t offset,
                            const uint64_t* dev_step, float* out_fp32, void* out_mant, int mant_kind,
                            uint64_t* counters, int update_range, void* stream) {
  if (!x || !integer_bits) return LBT_EINVAL;
  if (bits < 1 || bits > 31) return LBT_EINVAL;   // dfxp:21 accepts 1..32; 32 is the caller's pass-through (:22-23)
  // LBT_STATS_MINMAX: min/max tracking instead of exact overflow counts (valid for target_overflow_rate == 0)
  const bool minmax = (mode & LBT_STATS_MINMAX) != 0 && target_overflow_rate == 0.0f && counters != nullptr;
  mode &= ~LBT_STATS_MINMAX;
  if (mode < 0 || mode > 2) return LBT_EINVAL;
  if (mode == LBT_ROUND_STOCHASTIC_NOISE && !noise) return LBT_EINVAL;
  if (!out_fp32 && !out_mant && !counters) return LBT_EINVAL;  // nothing to produce
  if (out_mant && (mant_kind < LBT_MANT_S8 || mant_kind > LBT_MANT_S9C3)) return LBT_EINVAL;
  if (mant_kind == LBT_MANT_S9C3 && (bits > 9 || n_inner % 3 != 0 || out_fp32 || (reinterpret_cast<uintptr_t>(out_mant) & 15)))
    return LBT_EINVAL;
  if (!out_mant) mant_kind = LBT_MANT_NONE;
  if (mant_kind == LBT_MANT_S8 && bits > 8) return LBT_EINVAL;
  if (mant_kind == LBT_MANT_U8 && bits > 9) return LBT_EINVAL;
  if (mant_kind == LBT_MANT_S16 && bits > 16) return LBT_EINVAL;
  if (update_range && !counters) return LBT_EINVAL;
  if (n_outer == 0 || n_inner == 0) return LBT_OK;  // empty tensor: nothing to do, range untouched
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  QParams p{};
  p.x = x;
  p.n_outer = n_outer;
  p.n_inner = n_inner;
  p.bits = bits;
  p.ib = integer_bits;
  p.target = target_overflow_rate;
  p.noise = noise;
  p.seed = seed;
  p.offset = offset;
  p.dev_step = dev_step;
  p.out = out_fp32;
  p.mant = out_mant;
  p.mant_kind = mant_kind;
  p.counters = reinterpret_cast<unsigned long long*>(counters);
  p.update_range = update_range;

  const bool vec = (n_inner % 4 == 0) && (n_inner / 4 < 0xffffffffull) && aligned16(x) && (!out_fp32 || aligned16(out_fp32)) &&
                   (!out_mant || aligned16(out_mant)) && (mode != LBT_ROUND_STOCHASTIC_NOISE || aligned16(noise));
  const uint64_t cap = (uint64_t)di.sm_count * (uint64_t)g_blocks_per_sm;
  if (mant_kind == LBT_MANT_S9C3 && vec && n_inner % 12 == 0 && n_outer < (1ull << 31)) {
    p.n_vec = (uint32_t)(n_inner / 12);
    p.chunks = (p.n_vec + kThreads - 1) / kThreads;
    // row groups: enough tiles for ~8 CTAs per SM, at least 4 rows each so that the Philox blocks amortise
    uint64_t groups = ((uint64_t)di.sm_count * 8 + p.chunks - 1) / p.chunks;
    if (groups > (n_outer + 3) / 4) groups = (n_outer + 3) / 4;
    if (groups < 1) groups = 1;
    const uint64_t rpg = (n_outer + groups - 1) / groups;
    p.rows_per_group = (uint32_t)rpg;
    p.total_tiles = (uint64_t)p.chunks * ((n_outer + rpg - 1) / rpg);
    const unsigned grid = (unsigned)(p.total_tiles < cap ? p.total_tiles : cap);
    switch (mode) {
      case LBT_ROUND_NEAREST: launch_pdl(quantize_c3v_kernel<0>, grid, kThreads, 0, st, p); break;
      case LBT_ROUND_STOCHASTIC_NOISE: launch_pdl(quantize_c3v_kernel<1>, grid, kThreads, 0, st, p); break;
      default: launch_pdl(quantize_c3v_kernel<2>, grid, kThreads, 0, st, p); break;
    }
    return check_launch("lbt_quantize(c3, 4 pixels per thread)");
  }
  if (mant_kind == LBT_MANT_S9C3) {
    const uint64_t npix = (uint64_t)n_outer * (n_inner / 3);
    const uint64_t blocks = (npix + kThreads - 1) / kThreads;
    const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    switch (mode) {
      case LBT_ROUND_NEAREST: quantize_c3_kernel<0><<<grid, kThreads, 0, st>>>(p); break;
      case LBT_ROUND_STOCHASTIC_NOISE: quantize_c3_kernel<1><<<grid, kThreads, 0, st>>>(p); break;
      default: quantize_c3_kernel<2><<<grid, kThreads, 0, st>>>(p); break;
    }
    return check_launch("lbt_quantize(c3)");
  }
  if (vec) {
    p.n_vec = (uint32_t)(n_inner / 4);
    p.chunks = (p.n_vec + kThreads - 1) / kThreads;
    void (*kern)(const QParams) = nullptr;
    switch (mode) {
      case LBT_ROUND_NEAREST: kern = minmax ? quantize_vec_kernel<0, true> : quantize_vec_kernel<0, false>; break;
      case LBT_ROUND_STOCHASTIC_NOISE: kern = minmax ? quantize_vec_kernel<1, true> : quantize_vec_kernel<1, false>; break;
      default: kern = minmax ? quantize_vec_kernel<2, true> : quantize_vec_kernel<2, false>; break;
    }
    // One balanced wave of resident CTAs: a grid of 1.15 waves costs two CTA latencies on the few-MB tensors of a
    // training step, and a persistent grid larger than the resident set leaves the last wave part empty.
    static int occ_cache[16][3][2] = {};
    int& occ = occ_cache[di.device][mode][minmax ? 1 : 0];
    if (occ == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kThreads, 0) != cudaSuccess || n < 1) n = 1;
      occ = n;
    }
    const uint64_t wave = (uint64_t)di.sm_count * (uint64_t)(g_blocks_per_sm < occ ? g_blocks_per_sm : occ);
    uint64_t rpg, grid64;
    const uint64_t big_tiles = (uint64_t)p.chunks * ((n_outer + g_rows_per_group - 1) / g_rows_per_group);
    if (big_tiles >= 4 * wave) {
      // large tensor: many tiles per CTA; a grid of several CTA generations per SM evens out the tail (measured)
      rpg = (uint64_t)g_rows_per_group;
      grid64 = (uint64_t)di.sm_count * (uint64_t)g_blocks_per_sm;
    } else {
      uint64_t groups = wave / p.chunks;
      if (groups < 1) groups = 1;
      if (groups > n_outer) groups = n_outer;
      rpg = (n_outer + groups - 1) / groups;
      grid64 = wave;
    }
    p.rows_per_group = (uint32_t)rpg;
    p.total_tiles = (uint64_t)p.chunks * ((n_outer + rpg - 1) / rpg);
    const unsigned grid = (unsigned)(p.total_tiles < grid64 ? p.total_tiles : grid64);
    launch_pdl(kern, grid, kThreads, 0, st, p);
  } else {
    const uint64_t n = (uint64_t)n_outer * n_inner;
    const uint64_t blocks = (n + kThreads - 1) / kThreads;
    const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    switch (mode) {
      case LBT_ROUND_NEAREST: quantize_scalar_kernel<0><<<grid, kThreads, 0, st>>>(p); break;
      case LBT_ROUND_STOCHASTIC_NOISE: quantize_scalar_kernel<1><<<grid, kThreads, 0, st>>>(p); break;
      default: quantize_scalar_kernel<2><<<grid, kThreads, 0, st>>>(p); break;
    }
  }
  return check_launch("lbt_quantize");
}

extern "C" int lbt_quantize_residual(const float* grad, size_t n_grad_rows, float* buffer, size_t n_outer, size_t n_inner, int bits,
                                     int32_t* integer_bits, int mode, const float* noise, uint64_t seed, uint64_t offset,
                                     const uint64_t* dev_step, float* out, uint64_t* counters, void* stream) {
  if (!buffer || !integer_bits || (n_grad_rows && (!grad || !out))) return LBT_EINVAL;
  if (bits < 1 || bits > 31 || mode < 0 || mode > 2 || n_grad_rows > n_outer) return LBT_EINVAL;
  if (mode == LBT_ROUND_STOCHASTIC_NOISE && !noise) return LBT_EINVAL;
  if (n_outer == 0 || n_inner == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  QParams p{};
  p.x = grad;
  p.n_outer = n_outer;
  p.n_inner = n_inner;
  p.bits = bits;
  p.ib = integer_bits;
  p.noise = noise;
  p.seed = seed;
  p.offset = offset;
  p.dev_step = dev_step;
  p.out = out;
  p.counters = reinterpret_cast<unsigned long long*>(counters);
  const uint64_t n = (uint64_t)n_outer * n_inner, blocks = (n + kThreads - 1) / kThreads;
  const uint64_t cap = (uint64_t)di.sm_count * 8;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (mode) {
    case LBT_ROUND_NEAREST: quantize_residual_kernel<0><<<grid, kThreads, 0, st>>>(p, buffer, n_grad_rows); break;
    case LBT_ROUND_STOCHASTIC_NOISE: quantize_residual_kernel<1><<<grid, kThreads, 0, st>>>(p, buffer, n_grad_rows); break;
    default: quantize_residual_kernel<2><<<grid, kThreads, 0, st>>>(p, buffer, n_grad_rows); break;
  }
  return check_launch("lbt_quantize_residual");
}

extern "C" int lbt_noise_fill(float* u, size_t n_inner, uint64_t seed, uint64_t offset, const uint64_t* dev_step,
                              void* stream) {
  if (!u) return LBT_EINVAL;
  if (n_inner == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  const size_t ng = (n_inner + 3) / 4;
  const unsigned grid = (unsigned)((ng + 255) / 256 < 4096 ? (ng + 255) / 256 : 4096);
  noise_fill_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(u, n_inner, seed, offset, dev_step);
  return check_launch("lbt_noise_fill");
}

extern "C" int lbt_update_ranges(int32_t* ranges, uint64_t* counters, const int32_t* bits, const float* target,
                                 size_t n, void* stream) {
  if (!ranges || !counters || !bits) return LBT_EINVAL;
  if (n == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  update_ranges_kernel<<<(unsigned)((n + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ranges, reinterpret_cast<unsigned long long*>(counters), bits, target, n);
  return check_launch("lbt_update_ranges");
}

extern "C" int lbt_step_advance(uint64_t* dev_step, void* stream) {
  if (!dev_step) return LBT_EINVAL;
  LBT_REQUIRE_ARCH();
  step_advance_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dev_step);
  return check_launch("lbt_step_advance");
}

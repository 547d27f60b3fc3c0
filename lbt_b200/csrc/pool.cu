// Max pooling on NHWC fp32 activations, tf.nn.max_pool with 'SAME' (out-of-range taps ignored) or 'VALID' padding
// (/root/reference/dynamic_fixed_point.py:993-1006), forward + the gradient tf.gradients routes through it.
// HBM-bound: forward reads each input once (k*k/s^2 re-reads hit L1/L2), writes the pooled tensor and one byte per
// output (the winning tap); backward is a GATHER over the <= ceil(k/s)^2 windows that cover an input pixel, so it needs
// no atomics and writes dX exactly once.  Ties go to the first maximum in (row, column) scan order.
#include <cstdlib>

#include "common.cuh"

namespace lbt {
namespace {

constexpr int kThreads = 256;

struct PoolParams {
  int N, H, W, C, k, s, pt, pl, OH, OW;
  uint32_t c4;  // C / 4
  FastDiv d_c4, d_W, d_H, d_OW, d_OH, d_s, d_lanes;   // d_lanes: c4 / J of the multi-group backward kernel
};

// V = 4: a thread owns 4 consecutive channels (C % 4 == 0, 16-byte aligned tensors); V = 1: one channel (any C — the
// LeNet-style 6-channel layers of models.py:91-152).  p.c4 holds C / V.
template <int V>
__device__ __forceinline__ void ld_vec(const float* p, float (&v)[V]) {
  if constexpr (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ void st_vec(float* p, const float (&v)[V]) {
  if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else p[0] = v[0];
}
template <int V>
__device__ __forceinline__ void ld_idx(const uint8_t* p, unsigned char (&w)[V]) {
  if constexpr (V == 4) {
    const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(p));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  } else {
    w[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ void st_idx(uint8_t* p, const unsigned char (&w)[V]) {
  if constexpr (V == 4) *reinterpret_cast<uchar4*>(p) = make_uchar4(w[0], w[1], w[2], w[3]);
  else p[0] = w[0];
}

template <int V>
__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                uint8_t* __restrict__ idx, const PoolParams p, size_t total) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - pix * p.c4;
    uint32_t t = fastdiv(pix, p.d_OW);
    const int ow = (int)(pix - t * (uint32_t)p.OW);
    const int n = (int)fastdiv(t, p.d_OH);
    const int oh = (int)(t - (uint32_t)n * (uint32_t)p.OH);
    float m[V];
    unsigned char w[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      m[j] = -INFINITY;
      w[j] = 0;
    }
    const int h0 = oh * p.s - p.pt, w0 = ow * p.s - p.pl;
    for (int r = 0; r < p.k; ++r) {
      const int ih = h0 + r;
      if ((unsigned)ih >= (unsigned)p.H) continue;
      for (int q = 0; q < p.k; ++q) {
        const int iw = w0 + q;
        if ((unsigned)iw >= (unsigned)p.W) continue;
        float v[V];
        ld_vec<V>(x + (((size_t)n * p.H + ih) * p.W + iw) * p.C + (size_t)V * c, v);
        const unsigned char tt = (unsigned char)(r * p.k + q);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (v[j] > m[j] || v[j] != v[j]) {
            m[j] = v[j];
            w[j] = tt;
          }
      }
    }
    st_vec<V>(out + (size_t)V * i, m);
    st_idx<V>(idx + (size_t)V * i, w);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const float* __restrict__ g, const uint8_t* __restrict__ idx,
                                                                float* __restrict__ dx, const PoolParams p, size_t total) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - pix * p.c4;
    uint32_t t = fastdiv(pix, p.d_W);
    const int iw = (int)(pix - t * (uint32_t)p.W);
    const int n = (int)fastdiv(t, p.d_H);
    const int ih = (int)(t - (uint32_t)n * (uint32_t)p.H);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    // windows (oh, ow) with oh*s - pt <= ih < oh*s - pt + k
    const int th = ih + p.pt, tw = iw + p.pl;
    // ceil((th - k + 1) / s) for th - k + 1 >= 0, else 0
    const int oh_lo = th - p.k + 1 <= 0 ? 0 : (int)fastdiv((uint32_t)(th - p.k + p.s), p.d_s);
    const int ow_lo = tw - p.k + 1 <= 0 ? 0 : (int)fastdiv((uint32_t)(tw - p.k + p.s), p.d_s);
    const int oh_hi = min(p.OH - 1, (int)fastdiv((uint32_t)th, p.d_s)), ow_hi = min(p.OW - 1, (int)fastdiv((uint32_t)tw, p.d_s));
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      const int r = th - oh * p.s;
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int q = tw - ow * p.s;
        const unsigned char tt = (unsigned char)(r * p.k + q);
        const size_t o = ((((size_t)n * p.OH + oh) * p.OW + ow) * p.c4 + c) * V;
        unsigned char w[V];
        float gv[V];
        ld_idx<V>(idx + o, w);
        ld_vec<V>(g + o, gv);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (w[j] == tt) acc[j] += gv[j];
      }
    }
    st_vec<V>(dx + (size_t)V * i, acc);
  }
}

// Fast paths for the common windows (K = 2, 3): every tap of the window is LOADED before the first comparison, so a thread
// has K*K (forward) / up to 4 (backward) requests in flight instead of one — the generic kernels below issue a load, wait,
// compare, and only then the next load, which holds ResNet-18's 3x3/2 pool (1.08 GB of traffic) at 2.8 TB/s.  Same scan
// order, tie-breaking and NaN handling as the generic kernels.
template <int K>
__global__ void __launch_bounds__(kThreads) maxpool_fwd_k_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                  uint8_t* __restrict__ idx, const PoolParams p, size_t total) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - pix * p.c4;
    uint32_t t = fastdiv(pix, p.d_OW);
    const int ow = (int)(pix - t * (uint32_t)p.OW);
    const int n = (int)fastdiv(t, p.d_OH);
    const int oh = (int)(t - (uint32_t)n * (uint32_t)p.OH);
    const int h0 = oh * p.s - p.pt, w0 = ow * p.s - p.pl;
    float4 v[K * K];
    bool ok[K * K];
#pragma unroll
    for (int r = 0; r < K; ++r)
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const int ih = h0 + r, iw = w0 + q;
        ok[r * K + q] = (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
        if (ok[r * K + q]) v[r * K + q] = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * p.H + ih) * p.W + iw) * p.C) + c);
      }
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 w = make_uchar4(0, 0, 0, 0);
#pragma unroll
    for (int tq = 0; tq < K * K; ++tq)
      if (ok[tq]) {
        const unsigned char tt = (unsigned char)tq;
        const float4 a = v[tq];
        if (a.x > m.x || a.x != a.x) { m.x = a.x; w.x = tt; }
        if (a.y > m.y || a.y != a.y) { m.y = a.y; w.y = tt; }
        if (a.z > m.z || a.z != a.z) { m.z = a.z; w.z = tt; }
        if (a.w > m.w || a.w != a.w) { m.w = a.w; w.w = tt; }
      }
    reinterpret_cast<float4*>(out)[i] = m;
    reinterpret_cast<uchar4*>(idx)[i] = w;
  }
}

// Backward for windows with ceil(K / s) <= 2 per axis (K <= 2 * s): at most 2 x 2 windows cover an input pixel.
// J: float4 channel groups per thread.  The window geometry (a dozen fast divisions) depends on the pixel only, so a thread that
// owns J groups of one pixel pays it once (ncu of the J = 1 version on ResNet-18's stem: 57 instructions per element, issue-bound
// at 0.35 of the HBM roofline); the c4 / J lanes of a pixel read neighbouring 16-byte groups, i.e. whole sectors.
template <int K, int J>
__global__ void __launch_bounds__(kThreads, J == 1 ? 1 : 2) maxpool_bwd_k_kernel(const float* __restrict__ g, const uint8_t* __restrict__ idx,
                                                                  float* __restrict__ dx, const PoolParams p, size_t total) {
  pdl_trigger();
  pdl_wait();
  const uint32_t lanes = p.c4 / J;   // threads per pixel; thread `cq` owns groups cq, cq + lanes, ...
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = J == 1 ? fastdiv((uint32_t)i, p.d_c4) : fastdiv((uint32_t)i, p.d_lanes);
    const uint32_t cq = (uint32_t)i - pix * lanes;
    uint32_t t = fastdiv(pix, p.d_W);
    const int iw = (int)(pix - t * (uint32_t)p.W);
    const int n = (int)fastdiv(t, p.d_H);
    const int ih = (int)(t - (uint32_t)n * (uint32_t)p.H);
    const int th = ih + p.pt, tw = iw + p.pl;
    const int oh_lo = th - K + 1 <= 0 ? 0 : (int)fastdiv((uint32_t)(th - K + p.s), p.d_s);
    const int ow_lo = tw - K + 1 <= 0 ? 0 : (int)fastdiv((uint32_t)(tw - K + p.s), p.d_s);
    const int oh_hi = min(p.OH - 1, (int)fastdiv((uint32_t)th, p.d_s)), ow_hi = min(p.OW - 1, (int)fastdiv((uint32_t)tw, p.d_s));
    bool ok[4];
    unsigned char tap[4];
    size_t o[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int oh = oh_lo + a, ow = ow_lo + b;
        ok[2 * a + b] = oh <= oh_hi && ow <= ow_hi;
        tap[2 * a + b] = (unsigned char)((th - oh * p.s) * K + (tw - ow * p.s));
        o[2 * a + b] = ok[2 * a + b] ? ((((size_t)n * p.OH + oh) * p.OW + ow) * p.c4 + cq) : 0;
      }
    uchar4 wv[J][4];
    float4 gv[J][4];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (ok[q]) {
          wv[j][q] = __ldg(reinterpret_cast<const uchar4*>(idx) + o[q] + j * lanes);
          gv[j][q] = __ldg(reinterpret_cast<const float4*>(g) + o[q] + j * lanes);
        }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q)       // (oh, ow) ascending: the generic kernel's order of additions
        if (ok[q]) {
          if (wv[j][q].x == tap[q]) acc.x += gv[j][q].x;
          if (wv[j][q].y == tap[q]) acc.y += gv[j][q].y;
          if (wv[j][q].z == tap[q]) acc.z += gv[j][q].z;
          if (wv[j][q].w == tap[q]) acc.w += gv[j][q].w;
        }
      __stcs(reinterpret_cast<float4*>(dx) + ((size_t)pix * p.c4 + cq + j * lanes), acc);
    }
  }
}

// Stride 2, K <= 4: one thread owns a 2 x 2 block of input pixels (padded coordinates th = 2a + {0, 1}, tw = 2b + {0, 1}) of one
// channel group.  Exactly the windows (a - 1 .. a) x (b - 1 .. b) cover the block, so four gradient / index loads serve four
// pixels (the per-pixel kernels fetch up to four windows for EACH pixel), the window geometry is computed once per block, and the
// tap a window must have recorded for a pixel — r = py + 2 - 2 wy, q = px + 2 - 2 wx — is a compile-time constant.  Additions in
// (oh, ow) ascending order like the other kernels: bit-identical.  ~10 instructions per element instead of 57 (3x3 / 2).
template <int K>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_s2_kernel(const float* __restrict__ g, const uint8_t* __restrict__ idx,
                                                                   float* __restrict__ dx, const PoolParams p, uint32_t nBH, uint32_t nBW,
                                                                   FastDiv d_nBW, FastDiv d_nBH, size_t total) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t t = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - t * p.c4;
    uint32_t t2 = fastdiv(t, d_nBW);
    const int b = (int)(t - t2 * nBW);
    const int n = (int)fastdiv(t2, d_nBH);
    const int a = (int)(t2 - (uint32_t)n * nBH);
    uchar4 wv[4];
    float4 gv[4];
    bool ok[4];
#pragma unroll
    for (int wy = 0; wy < 2; ++wy)
#pragma unroll
      for (int wx = 0; wx < 2; ++wx) {
        const int oh = a - 1 + wy, ow = b - 1 + wx;
        ok[2 * wy + wx] = (unsigned)oh < (unsigned)p.OH && (unsigned)ow < (unsigned)p.OW;
        if (ok[2 * wy + wx]) {
          const size_t o = ((((size_t)n * p.OH + oh) * p.OW + ow) * p.c4 + c);
          wv[2 * wy + wx] = __ldg(reinterpret_cast<const uchar4*>(idx) + o);
          gv[2 * wy + wx] = __ldg(reinterpret_cast<const float4*>(g) + o);
        }
      }
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const int ih = 2 * a + py - p.pt, iw = 2 * b + px - p.pl;
        if ((unsigned)ih >= (unsigned)p.H || (unsigned)iw >= (unsigned)p.W) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int wy = 0; wy < 2; ++wy)
#pragma unroll
          for (int wx = 0; wx < 2; ++wx) {
            constexpr int kNone = 255;
            const int r = py + 2 - 2 * wy, q = px + 2 - 2 * wx;
            const int tap = (r < K && q < K) ? r * K + q : kNone;   // compile-time after unrolling
            if (tap != kNone && ok[2 * wy + wx]) {
              const uchar4 w = wv[2 * wy + wx];
              const float4 v = gv[2 * wy + wx];
              if (w.x == tap) acc.x += v.x;
              if (w.y == tap) acc.y += v.y;
              if (w.z == tap) acc.z += v.z;
              if (w.w == tap) acc.w += v.w;
            }
          }
        __stcs(reinterpret_cast<float4*>(dx) + ((((size_t)n * p.H + ih) * p.W + iw) * p.c4 + c), acc);
      }
  }
}

void fill_divs(PoolParams& p) {
  p.d_c4 = make_fastdiv(p.c4);
  p.d_W = make_fastdiv((uint32_t)p.W);
  p.d_H = make_fastdiv((uint32_t)p.H);
  p.d_OW = make_fastdiv((uint32_t)p.OW);
  p.d_OH = make_fastdiv((uint32_t)p.OH);
  p.d_s = make_fastdiv((uint32_t)p.s);
}

int check(const PoolParams& p) {
  if (p.N <= 0 || p.H <= 0 || p.W <= 0 || p.C <= 0 || p.k <= 0 || p.s <= 0 || p.OH <= 0 || p.OW <= 0) return LBT_EINVAL;
  if (p.k > 15 || p.pt < 0 || p.pl < 0) return LBT_EUNSUPPORTED;
  return LBT_OK;
}

// vector width of a launch: 4 channels per thread when C % 4 == 0 and every tensor is 16-byte (index bytes: 4-byte) aligned
inline int pool_vec(int C, const void* a, const void* b, const void* idx) {
  const bool al = !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) && !(reinterpret_cast<uintptr_t>(idx) & 3);
  return ((C & 3) == 0 && al) ? 4 : 1;
}

}  // namespace
}  // namespace lbt

using namespace lbt;

extern "C" int lbt_maxpool_fwd(const float* x, int N, int H, int W, int C, int k, int s, int pad_top, int pad_left, int OH, int OW,
                               float* out, uint8_t* idx, void* stream) {
  if (!x || !out || !idx) return LBT_EINVAL;
  PoolParams p{};
  const int V = pool_vec(C, x, out, idx);
  p.N = N; p.H = H; p.W = W; p.C = C; p.k = k; p.s = s; p.pt = pad_top; p.pl = pad_left; p.OH = OH; p.OW = OW; p.c4 = (uint32_t)(C / V);
  int rc = check(p);
  if (rc) return rc;
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)N * OH * OW * p.c4;
  if (total >= (1ull << 31)) return LBT_EUNSUPPORTED;
  fill_divs(p);
  const size_t blocks = (total + kThreads - 1) / kThreads;
  const size_t cap = (size_t)device_info().sm_count * 8;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (V == 1) launch_pdl(maxpool_fwd_kernel<1>, grid, kThreads, 0, st, x, out, idx, p, total);
  else if (k == 3) launch_pdl(maxpool_fwd_k_kernel<3>, grid, kThreads, 0, st, x, out, idx, p, total);
  else if (k == 2) launch_pdl(maxpool_fwd_k_kernel<2>, grid, kThreads, 0, st, x, out, idx, p, total);
  else launch_pdl(maxpool_fwd_kernel<4>, grid, kThreads, 0, st, x, out, idx, p, total);
  return check_launch("lbt_maxpool_fwd");
}

extern "C" int lbt_maxpool_bwd(const float* g, const uint8_t* idx, int N, int H, int W, int C, int k, int s, int pad_top,
                               int pad_left, int OH, int OW, float* dx, void* stream) {
  if (!g || !idx || !dx) return LBT_EINVAL;
  PoolParams p{};
  const int V = pool_vec(C, g, dx, idx);
  p.N = N; p.H = H; p.W = W; p.C = C; p.k = k; p.s = s; p.pt = pad_top; p.pl = pad_left; p.OH = OH; p.OW = OW; p.c4 = (uint32_t)(C / V);
  int rc = check(p);
  if (rc) return rc;
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)N * H * W * p.c4;
  if (total >= (1ull << 31)) return LBT_EUNSUPPORTED;
  fill_divs(p);
  const size_t blocks = (total + kThreads - 1) / kThreads;
  const size_t cap = (size_t)device_info().sm_count * 8;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const int pool_blk = [] {   // A/B: LBT_POOL_BLOCK=0 keeps the per-pixel kernels for stride 2
    const char* e = std::getenv("LBT_POOL_BLOCK");
    return e ? std::atoi(e) : 1;
  }();
  if (V == 4 && s == 2 && (k == 2 || k == 3 || k == 4) && pad_top < 2 && pad_left < 2 && pool_blk) {
    // blocks of padded rows th = 2a + {0,1}: a from 0 (pad < 2) to (pad + H - 1) / 2
    const uint32_t nBH = (uint32_t)((pad_top + H - 1) / 2 + 1), nBW = (uint32_t)((pad_left + W - 1) / 2 + 1);
    const size_t totalb = (size_t)N * nBH * nBW * p.c4;
    if (totalb < (1ull << 31)) {
      const size_t blocksb = (totalb + kThreads - 1) / kThreads;
      const unsigned gridb = (unsigned)(blocksb < cap ? blocksb : cap);
      const FastDiv dW = make_fastdiv(nBW), dH = make_fastdiv(nBH);
      if (k == 2) launch_pdl(maxpool_bwd_s2_kernel<2>, gridb, kThreads, 0, st, g, idx, dx, p, nBH, nBW, dW, dH, totalb);
      else if (k == 3) launch_pdl(maxpool_bwd_s2_kernel<3>, gridb, kThreads, 0, st, g, idx, dx, p, nBH, nBW, dW, dH, totalb);
      else launch_pdl(maxpool_bwd_s2_kernel<4>, gridb, kThreads, 0, st, g, idx, dx, p, nBH, nBW, dW, dH, totalb);
      return check_launch("lbt_maxpool_bwd");
    }
  }
  static const int pool_j = [] {   // channel groups per thread (A/B: LBT_POOL_J = 1, 2 or 4)
    const char* e = std::getenv("LBT_POOL_J");
    return e ? std::atoi(e) : 2;
  }();
  const int J = (pool_j >= 4 && p.c4 % 4 == 0) ? 4 : ((pool_j >= 2 && p.c4 % 2 == 0) ? 2 : 1);
  if (V == 4 && (k == 2 || k == 3) && k <= 2 * s && J > 1) {   // several channel groups per thread
    const size_t totalj = total / J, blocksj = (totalj + kThreads - 1) / kThreads;
    const unsigned gridj = (unsigned)(blocksj < cap ? blocksj : cap);
    p.d_lanes = make_fastdiv(p.c4 / J);
    if (k == 3 && J == 4) launch_pdl(maxpool_bwd_k_kernel<3, 4>, gridj, kThreads, 0, st, g, idx, dx, p, totalj);
    else if (k == 3) launch_pdl(maxpool_bwd_k_kernel<3, 2>, gridj, kThreads, 0, st, g, idx, dx, p, totalj);
    else if (J == 4) launch_pdl(maxpool_bwd_k_kernel<2, 4>, gridj, kThreads, 0, st, g, idx, dx, p, totalj);
    else launch_pdl(maxpool_bwd_k_kernel<2, 2>, gridj, kThreads, 0, st, g, idx, dx, p, totalj);
    return check_launch("lbt_maxpool_bwd");
  }
  if (V == 1) launch_pdl(maxpool_bwd_kernel<1>, grid, kThreads, 0, st, g, idx, dx, p, total);
  else if (k == 3 && k <= 2 * s) launch_pdl(maxpool_bwd_k_kernel<3, 1>, grid, kThreads, 0, st, g, idx, dx, p, total);
  else if (k == 2 && k <= 2 * s) launch_pdl(maxpool_bwd_k_kernel<2, 1>, grid, kThreads, 0, st, g, idx, dx, p, total);
  else launch_pdl(maxpool_bwd_kernel<4>, grid, kThreads, 0, st, g, idx, dx, p, total);
  return check_launch("lbt_maxpool_bwd");
}

// ------------------------------------------------------------------------------------------------------------------
// tf.nn.avg_pool 'VALID' on NHWC fp32 (`AvgPool_q`, dynamic_fixed_point.py:1009-1022) and its gradient, and the mean
// sparse softmax cross-entropy of models.py:30-32 with its gradient w.r.t. the logits.  Small kernels that keep the last
// few operations of a training step inside the library instead of a handful of framework launches.
// ------------------------------------------------------------------------------------------------------------------
namespace lbt {
namespace {

template <int V>
__global__ void __launch_bounds__(kThreads) avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                const PoolParams p, size_t total) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - pix * p.c4;
    uint32_t t = fastdiv(pix, p.d_OW);
    const int ow = (int)(pix - t * (uint32_t)p.OW);
    const int n = (int)fastdiv(t, p.d_OH);
    const int oh = (int)(t - (uint32_t)n * (uint32_t)p.OH);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    // the window is summed in (r, q) order like the oracle; the loads of 16 taps are issued together so a global 8x8
    // pool costs 4 memory round trips per thread instead of 64
    const int taps = p.k * p.k;
    const float* base = x + (((size_t)n * p.H + oh * p.s) * p.W + ow * p.s) * p.C + (size_t)V * c;
    for (int t0 = 0; t0 < taps; t0 += 16) {
      float v[16][V];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int tt = t0 + j;
        if (tt < taps) {
          const int r = tt / p.k, q = tt - r * p.k;
          ld_vec<V>(base + ((size_t)r * p.W + q) * p.C, v[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (t0 + j < taps) {
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = __fadd_rn(acc[e], v[j][e]);
        }
    }
    const float d = (float)(p.k * p.k);   // the window sum is DIVIDED by k*k (same arithmetic as the oracle's restatement)
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = __fdiv_rn(acc[e], d);
    st_vec<V>(out + (size_t)V * i, acc);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads) avgpool_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, const PoolParams p,
                                                                size_t total) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  const float d = (float)(p.k * p.k);
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - pix * p.c4;
    uint32_t t = fastdiv(pix, p.d_W);
    const int iw = (int)(pix - t * (uint32_t)p.W);
    const int n = (int)fastdiv(t, p.d_H);
    const int ih = (int)(t - (uint32_t)n * (uint32_t)p.H);
    const int oh_lo = ih - p.k + 1 <= 0 ? 0 : (int)fastdiv((uint32_t)(ih - p.k + p.s), p.d_s);
    const int ow_lo = iw - p.k + 1 <= 0 ? 0 : (int)fastdiv((uint32_t)(iw - p.k + p.s), p.d_s);
    const int oh_hi = min(p.OH - 1, (int)fastdiv((uint32_t)ih, p.d_s)), ow_hi = min(p.OW - 1, (int)fastdiv((uint32_t)iw, p.d_s));
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        float gv[V];
        ld_vec<V>(g + ((((size_t)n * p.OH + oh) * p.OW + ow) * p.c4 + c) * V, gv);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = __fadd_rn(acc[e], __fdiv_rn(gv[e], d));
      }
    st_vec<V>(dx + (size_t)V * i, acc);
  }
}

// ONE CTA of 32 warps (the logits are a few hundred KB at most): warp w takes rows w, w+32, ...; p = softmax(logits) is
// kept for the backward; the mean is summed in a FIXED order (per warp, then over the 32 warps) so the loss is
// bit-reproducible from run to run.
constexpr int kXentStage = 8192;   // floats of shared memory: logits of small problems are staged with ONE coalesced pass

__global__ void __launch_bounds__(1024) xent_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C,
                                                        float* __restrict__ probs, float* __restrict__ loss) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  __shared__ float s_part[32];
  __shared__ float s_x[kXentStage];
  __shared__ int s_y[1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool staged = (size_t)B * C <= (size_t)kXentStage && B <= 1024;
  if (staged) {   // one memory round trip for the whole problem instead of four dependent ones per row
    for (int i = threadIdx.x; i < B * C; i += 1024) s_x[i] = logits[i];
    if ((int)threadIdx.x < B) {
      const long long y = labels[threadIdx.x];
      s_y[threadIdx.x] = (y >= 0 && y < C) ? (int)y : -1;
    }
    __syncthreads();
  }
  float part = 0.f;
  for (int row = warp; row < B; row += 32) {
    const float* x = staged ? s_x + (size_t)row * C : logits + (size_t)row * C;
    const long long y = staged ? (long long)s_y[row] : labels[row];
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(x[c] - m);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float ls = logf(s);
    for (int c = lane; c < C; c += 32) probs[(size_t)row * C + c] = expf(x[c] - m - ls);
    if (y >= 0 && y < C) part += -(x[y] - m - ls);
  }
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += s_part[w];
    *loss = t / (float)B;
  }
}

// The same computation spread over a thread-block cluster: one CTA was the whole 80 us of a 256 x 1000 problem.  CTA r of the
// cluster takes the 32-row groups g with g % kXentCluster == r; every row's loss goes to CTA 0's shared memory (distributed
// shared memory), and CTA 0 adds them in exactly the single-CTA kernel's order (warp w: rows w, w + 32, ... from 0.0f; then
// the 32 warp partials in order), so the loss is bit-identical.  Rows never interact elsewhere.
constexpr int kXentCluster = 8;
constexpr int kXentMaxRows = 4096;
__global__ void __cluster_dims__(kXentCluster, 1, 1) __launch_bounds__(1024)
xent_fwd_cluster_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C, float* __restrict__ probs,
                        float* __restrict__ loss) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_row[kXentMaxRows];   // used in CTA 0 only
  __shared__ float s_part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint32_t row0_addr;   // CTA 0's s_row in the cluster's shared window
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(row0_addr) : "r"((uint32_t)__cvta_generic_to_shared(s_row)), "r"(0));
  for (int row = warp + 32 * (int)rank; row < B; row += 32 * kXentCluster) {
    const float* x = logits + (size_t)row * C;
    const long long y = labels[row];
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(x[c] - m);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float ls = logf(s);
    for (int c = lane; c < C; c += 32) probs[(size_t)row * C + c] = expf(x[c] - m - ls);
    if (lane == 0) {
      const float l = (y >= 0 && y < C) ? -(x[y] - m - ls) : 0.f;   // + 0.0f leaves the running sum as the skipped row did
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(row0_addr + 4u * (uint32_t)row), "f"(l) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank != 0) return;
  if (lane == 0) {
    float part = 0.f;
    for (int row = warp; row < B; row += 32) part += s_row[row];
    s_part[warp] = part;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += s_part[w];
    *loss = t / (float)B;
  }
}

__global__ void __launch_bounds__(256) xent_bwd_kernel(const float* __restrict__ probs, const long long* __restrict__ labels,
                                                       const float* __restrict__ gout, int B, int C, float* __restrict__ dlogits) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  const float scale = (gout ? *gout : 1.0f) / (float)B;
  const size_t total = (size_t)B * C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int row = (int)(i / C), c = (int)(i % C);
    dlogits[i] = (probs[i] - (labels[row] == c ? 1.0f : 0.0f)) * scale;
  }
}

}  // namespace
}  // namespace lbt

extern "C" int lbt_avgpool_fwd(const float* x, int N, int H, int W, int C, int k, int s, int OH, int OW, float* out, void* stream) {
  if (!x || !out) return LBT_EINVAL;
  PoolParams p{};
  const int V = pool_vec(C, x, out, nullptr);
  p.N = N; p.H = H; p.W = W; p.C = C; p.k = k; p.s = s; p.pt = 0; p.pl = 0; p.OH = OH; p.OW = OW; p.c4 = (uint32_t)(C / V);
  int rc = check(p);
  if (rc) return rc;
  if ((OH - 1) * s + k > H || (OW - 1) * s + k > W) return LBT_EINVAL;   // 'VALID': every window inside the input
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)N * OH * OW * p.c4;
  if (total >= (1ull << 31)) return LBT_EUNSUPPORTED;
  fill_divs(p);
  const size_t blocks = (total + kThreads - 1) / kThreads, cap = (size_t)device_info().sm_count * 8;
  if (V == 4) launch_pdl(avgpool_fwd_kernel<4>, (unsigned)(blocks < cap ? blocks : cap), kThreads, 0, reinterpret_cast<cudaStream_t>(stream), x, out, p, total);
  else launch_pdl(avgpool_fwd_kernel<1>, (unsigned)(blocks < cap ? blocks : cap), kThreads, 0, reinterpret_cast<cudaStream_t>(stream), x, out, p, total);
  return check_launch("lbt_avgpool_fwd");
}

extern "C" int lbt_avgpool_bwd(const float* g, int N, int H, int W, int C, int k, int s, int OH, int OW, float* dx, void* stream) {
  if (!g || !dx) return LBT_EINVAL;
  PoolParams p{};
  const int V = pool_vec(C, g, dx, nullptr);
  p.N = N; p.H = H; p.W = W; p.C = C; p.k = k; p.s = s; p.pt = 0; p.pl = 0; p.OH = OH; p.OW = OW; p.c4 = (uint32_t)(C / V);
  int rc = check(p);
  if (rc) return rc;
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)N * H * W * p.c4;
  if (total >= (1ull << 31)) return LBT_EUNSUPPORTED;
  fill_divs(p);
  const size_t blocks = (total + kThreads - 1) / kThreads, cap = (size_t)device_info().sm_count * 8;
  if (V == 4) launch_pdl(avgpool_bwd_kernel<4>, (unsigned)(blocks < cap ? blocks : cap), kThreads, 0, reinterpret_cast<cudaStream_t>(stream), g, dx, p, total);
  else launch_pdl(avgpool_bwd_kernel<1>, (unsigned)(blocks < cap ? blocks : cap), kThreads, 0, reinterpret_cast<cudaStream_t>(stream), g, dx, p, total);
  return check_launch("lbt_avgpool_bwd");
}

extern "C" int lbt_softmax_xent_fwd(const float* logits, const int64_t* labels, int B, int C, float* probs, float* loss, void* stream) {
  if (!logits || !labels || !probs || !loss || B <= 0 || C <= 0) return LBT_EINVAL;
  LBT_REQUIRE_ARCH();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B > 32 && B <= kXentMaxRows && (size_t)B * C > (size_t)kXentStage)   // small problems: one CTA with staged logits
    launch_pdl(xent_fwd_cluster_kernel, kXentCluster, 1024, 0, st, logits, reinterpret_cast<const long long*>(labels), B, C, probs, loss);
  else
    launch_pdl(xent_fwd_kernel, 1, 1024, 0, st, logits, reinterpret_cast<const long long*>(labels), B, C, probs, loss);
  return check_launch("lbt_softmax_xent_fwd");
}

extern "C" int lbt_softmax_xent_bwd(const float* probs, const int64_t* labels, const float* grad_loss, int B, int C, float* dlogits,
                                    void* stream) {
  if (!probs || !labels || !dlogits || B <= 0 || C <= 0) return LBT_EINVAL;
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)B * C, blocks = (total + 255) / 256, cap = (size_t)device_info().sm_count * 8;
  launch_pdl(xent_bwd_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      probs, reinterpret_cast<const long long*>(labels), grad_loss, B, C, dlogits);
  return check_launch("lbt_softmax_xent_bwd");
}

// ------------------------------------------------------------------------------------------------------------------
// ReLU_q (tf.maximum(0.0, X), dynamic_fixed_point.py:983-990) and Dropout_q (tf.nn.dropout(x, keep) = x / keep *
// floor(keep + u), :1025-1040) for the models without batch-norm (where they are not folded into a BN kernel).  The
// dropout uniforms are an explicit tensor (parity tests) or the Philox stream (seed, offset + (*dev_step << 32)); the
// backward pass recomputes the same mask from the same stream, so no mask tensor is stored.
// ------------------------------------------------------------------------------------------------------------------
namespace lbt {
namespace {

__global__ void __launch_bounds__(256) relu_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ out,
                                                   size_t n4, size_t n) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  // g == NULL: out = max(0, x);  else: out = g where x > 0 (x = the forward OUTPUT or input: same sign test), else 0
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 o;
    if (g) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
      o = make_float4(v.x > 0.f ? gv.x : 0.f, v.y > 0.f ? gv.y : 0.f, v.z > 0.f ? gv.z : 0.f, v.w > 0.f ? gv.w : 0.f);
    } else {
      o = make_float4(fmaxf(0.f, v.x), fmaxf(0.f, v.y), fmaxf(0.f, v.z), fmaxf(0.f, v.w));
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    out[i] = g ? (x[i] > 0.f ? g[i] : 0.f) : fmaxf(0.f, x[i]);
}

__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, const float* __restrict__ u, float keep, uint64_t seed,
                                                      uint64_t offset, const uint64_t* dev_step, float* __restrict__ out, size_t n) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  const uint64_t off = offset + (dev_step ? ((*dev_step) << 32) : 0ull);
  const size_t ng = (n + 3) / 4;
  for (size_t gidx = (size_t)blockIdx.x * 256 + threadIdx.x; gidx < ng; gidx += (size_t)gridDim.x * 256) {
    float4 r;
    const size_t e = 4 * gidx;
    if (u) {
      r.x = e + 0 < n ? u[e + 0] : 0.f;
      r.y = e + 1 < n ? u[e + 1] : 0.f;
      r.z = e + 2 < n ? u[e + 2] : 0.f;
      r.w = e + 3 < n ? u[e + 3] : 0.f;
    } else {
      r = philox_noise4(gidx, seed, off);
    }
    const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (e + j < n) out[e + j] = __fmul_rn(__fdiv_rn(x[e + j], keep), floorf(__fadd_rn(keep, rr[j])));
  }
}

}  // namespace
}  // namespace lbt

extern "C" int lbt_relu(const float* x, const float* g, float* out, size_t n, void* stream) {
  if (!x || !out) return LBT_EINVAL;
  if (n == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  const bool al = !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | (g ? reinterpret_cast<uintptr_t>(g) : 0)) & 15);
  const size_t n4 = al ? n / 4 : 0;
  const size_t blocks = ((n4 ? n4 : n) + 255) / 256, cap = (size_t)device_info().sm_count * 8;
  launch_pdl(relu_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, g, out, n4, n);
  return check_launch("lbt_relu");
}

extern "C" int lbt_dropout(const float* x, const float* u, float keep_prob, uint64_t seed, uint64_t offset, const uint64_t* dev_step,
                           float* out, size_t n, void* stream) {
  if (!x || !out || !(keep_prob > 0.0f)) return LBT_EINVAL;
  if (n == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  const size_t blocks = ((n + 3) / 4 + 255) / 256, cap = (size_t)device_info().sm_count * 8;
  launch_pdl(dropout_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream), x, u, keep_prob, seed, offset,
                                                                                                       dev_step, out, n);
  return check_launch("lbt_dropout");
}

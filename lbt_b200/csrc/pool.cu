// Max pooling on NHWC fp32 activations, tf.nn.max_pool with 'SAME' (out-of-range taps ignored) or 'VALID' padding
// (/root/reference/dynamic_fixed_point.py:993-1006), forward + the gradient tf.gradients routes through it.
// HBM-bound: forward reads each input once (k*k/s^2 re-reads hit L1/L2), writes the pooled tensor and one byte per
// output (the winning tap); backward is a GATHER over the <= ceil(k/s)^2 windows that cover an input pixel, so it needs
// no atomics and writes dX exactly once.  Ties go to the first maximum in (row, column) scan order.
#include "common.cuh"

namespace lbt {
namespace {

constexpr int kThreads = 256;

struct PoolParams {
  int N, H, W, C, k, s, pt, pl, OH, OW;
  uint32_t c4;  // C / 4
  FastDiv d_c4, d_W, d_H, d_OW, d_OH, d_s;
};

__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                uint8_t* __restrict__ idx, const PoolParams p, size_t total) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - pix * p.c4;
    uint32_t t = fastdiv(pix, p.d_OW);
    const int ow = (int)(pix - t * (uint32_t)p.OW);
    const int n = (int)fastdiv(t, p.d_OH);
    const int oh = (int)(t - (uint32_t)n * (uint32_t)p.OH);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 w = make_uchar4(0, 0, 0, 0);
    const int h0 = oh * p.s - p.pt, w0 = ow * p.s - p.pl;
    for (int r = 0; r < p.k; ++r) {
      const int ih = h0 + r;
      if ((unsigned)ih >= (unsigned)p.H) continue;
      for (int q = 0; q < p.k; ++q) {
        const int iw = w0 + q;
        if ((unsigned)iw >= (unsigned)p.W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * p.H + ih) * p.W + iw) * p.C) + c);
        const unsigned char t = (unsigned char)(r * p.k + q);
        if (v.x > m.x || v.x != v.x) { m.x = v.x; w.x = t; }
        if (v.y > m.y || v.y != v.y) { m.y = v.y; w.y = t; }
        if (v.z > m.z || v.z != v.z) { m.z = v.z; w.z = t; }
        if (v.w > m.w || v.w != v.w) { m.w = v.w; w.w = t; }
      }
    }
    reinterpret_cast<float4*>(out)[i] = m;
    reinterpret_cast<uchar4*>(idx)[i] = w;
  }
}

__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const float* __restrict__ g, const uint8_t* __restrict__ idx,
                                                                float* __restrict__ dx, const PoolParams p, size_t total) {
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kThreads) {
    uint32_t pix = fastdiv((uint32_t)i, p.d_c4);
    const uint32_t c = (uint32_t)i - pix * p.c4;
    uint32_t t = fastdiv(pix, p.d_W);
    const int iw = (int)(pix - t * (uint32_t)p.W);
    const int n = (int)fastdiv(t, p.d_H);
    const int ih = (int)(t - (uint32_t)n * (uint32_t)p.H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
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
        const unsigned char t = (unsigned char)(r * p.k + q);
        const size_t o = ((((size_t)n * p.OH + oh) * p.OW + ow) * p.c4 + c);
        const uchar4 w = __ldg(reinterpret_cast<const uchar4*>(idx) + o);
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + o);
        if (w.x == t) acc.x += gv.x;
        if (w.y == t) acc.y += gv.y;
        if (w.z == t) acc.z += gv.z;
        if (w.w == t) acc.w += gv.w;
      }
    }
    reinterpret_cast<float4*>(dx)[i] = acc;
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
  if ((p.C & 3) || p.k > 15 || p.pt < 0 || p.pl < 0) return LBT_EUNSUPPORTED;
  return LBT_OK;
}

}  // namespace
}  // namespace lbt

using namespace lbt;

extern "C" int lbt_maxpool_fwd(const float* x, int N, int H, int W, int C, int k, int s, int pad_top, int pad_left, int OH, int OW,
                               float* out, uint8_t* idx, void* stream) {
  if (!x || !out || !idx) return LBT_EINVAL;
  PoolParams p{};
  p.N = N; p.H = H; p.W = W; p.C = C; p.k = k; p.s = s; p.pt = pad_top; p.pl = pad_left; p.OH = OH; p.OW = OW; p.c4 = (uint32_t)(C / 4);
  int rc = check(p);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(idx) & 3))
    return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)N * OH * OW * p.c4;
  if (total >= (1ull << 31)) return LBT_EUNSUPPORTED;
  fill_divs(p);
  const size_t blocks = (total + kThreads - 1) / kThreads;
  const size_t cap = (size_t)device_info().sm_count * 8;
  maxpool_fwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, idx, p,
                                                                                                                  total);
  return check_launch("lbt_maxpool_fwd");
}

extern "C" int lbt_maxpool_bwd(const float* g, const uint8_t* idx, int N, int H, int W, int C, int k, int s, int pad_top,
                               int pad_left, int OH, int OW, float* dx, void* stream) {
  if (!g || !idx || !dx) return LBT_EINVAL;
  PoolParams p{};
  p.N = N; p.H = H; p.W = W; p.C = C; p.k = k; p.s = s; p.pt = pad_top; p.pl = pad_left; p.OH = OH; p.OW = OW; p.c4 = (uint32_t)(C / 4);
  int rc = check(p);
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(dx) & 15) || (reinterpret_cast<uintptr_t>(idx) & 3))
    return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)N * H * W * p.c4;
  if (total >= (1ull << 31)) return LBT_EUNSUPPORTED;
  fill_divs(p);
  const size_t blocks = (total + kThreads - 1) / kThreads;
  const size_t cap = (size_t)device_info().sm_count * 8;
  maxpool_bwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, idx, dx, p,
                                                                                                                  total);
  return check_launch("lbt_maxpool_bwd");
}

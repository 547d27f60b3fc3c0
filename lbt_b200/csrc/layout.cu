// Data-layout kernels around the int8 GEMM: im2col gathers of NHWC mantissas (fprop and the
// transposed gather of dgrad), byte-matrix transpose, and the bias-gradient column sum.
// All are HBM/L2-bound byte movers: 16-byte vector accesses, grids sized in multiples of the SM count.
//
// They stand in for what cuDNN did inside tf.nn.conv2d / Conv2DBackpropInput / Conv2DBackpropFilter
// (/root/reference/dynamic_fixed_point.py:291, 302-305) with TF's NHWC x HWIO layout and 'SAME' padding
// (pad_before = total/2, the extra pixel on the bottom/right).
#include "common.cuh"

namespace lbt {
namespace {

struct Im2colParams {
  const void* src;
  int src_kind;  // LBT_MANT_S8 / U8 (bytes) or LBT_MANT_S16 (split into hi|hi|lo segments)
  int N, H, W, C;        // source tensor [N, H, W, C]
  int OH, OW;            // spatial grid the rows iterate over
  int kh, kw, sh, sw, pt, pl;
  int transposed;        // 0: fprop gather, 1: dgrad (transposed-conv) gather
  uint8_t* out;
  size_t ld;             // row pitch of out in bytes
  int K;                 // kh*kw*C
  size_t M;              // N*OH*OW rows
  FastDiv d_kvecs, d_cvecs, d_kw, d_OW, d_OH, d_sh, d_sw;   // vec16 kernel: no integer divisions in the loop
};

// Source coordinate of output row (n, oh, ow), tap (r, s).  Returns false when the tap reads padding.
__device__ __forceinline__ bool tap_coord(const Im2colParams& p, int oh, int ow, int r, int s, int& ih, int& iw) {
  if (!p.transposed) {
    ih = oh * p.sh - p.pt + r;  // fprop: input pixel under filter tap (r, s) of output pixel (oh, ow)
    iw = ow * p.sw - p.pl + s;
  } else {
    // dgrad: rows iterate over INPUT pixels (oh, ow); the source is the output-gradient map; tap (r, s)
    // contributes iff some output pixel (ih, iw) has ih*sh - pt + r == oh.
    const int th = oh + p.pt - r, tw = ow + p.pl - s;
    if (th < 0 || tw < 0 || (th % p.sh) != 0 || (tw % p.sw) != 0) return false;
    ih = th / p.sh;
    iw = tw / p.sw;
  }
  return ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
}

// C % 16 == 0: one thread moves 16 channels of one tap (one 16-byte load, one 16-byte store).  total < 2^31.
__global__ void __launch_bounds__(256) im2col_vec16_kernel(const Im2colParams p) {
  const uint32_t kvecs = (uint32_t)p.K >> 4;
  const uint32_t cvecs = (uint32_t)p.C >> 4;
  const uint32_t total = (uint32_t)p.M * kvecs;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t m = fastdiv(i, p.d_kvecs);
    const uint32_t kv = i - m * kvecs;
    const uint32_t tap = fastdiv(kv, p.d_cvecs), cv = kv - tap * cvecs;
    const int r = (int)fastdiv(tap, p.d_kw), s = (int)(tap - (uint32_t)r * (uint32_t)p.kw);
    const uint32_t t = fastdiv(m, p.d_OW);
    const int ow = (int)(m - t * (uint32_t)p.OW);
    const int n = (int)fastdiv(t, p.d_OH);
    const int oh = (int)(t - (uint32_t)n * (uint32_t)p.OH);
    int ih, iw;
    bool ok;
    if (!p.transposed) {
      ih = oh * p.sh - p.pt + r;
      iw = ow * p.sw - p.pl + s;
      ok = true;
    } else {
      const int th = oh + p.pt - r, tw = ow + p.pl - s;
      ok = th >= 0 && tw >= 0;
      ih = ok ? (int)fastdiv((uint32_t)th, p.d_sh) : 0;
      iw = ok ? (int)fastdiv((uint32_t)tw, p.d_sw) : 0;
      ok = ok && ih * p.sh == th && iw * p.sw == tw;
    }
    ok = ok && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ok)
      v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.src) +
                                               (((size_t)n * p.H + ih) * p.W + iw) * p.C) + cv);
    *reinterpret_cast<uint4*>(p.out + m * p.ld + ((size_t)kv << 4)) = v;
  }
}

// Any C, any source kind: one thread per output byte.  S16 sources (signed 9..16-bit mantissas k) are
// written as three K-segments [hi | hi | lo] with k = 2*hi + lo, hi = k >> 1 (s8), lo = k & 1, so that
// an s8 GEMM against [W | W | W] reproduces sum(k * w) exactly.
__global__ void __launch_bounds__(256) im2col_scalar_kernel(const Im2colParams p) {
  const int segs = p.src_kind == LBT_MANT_S16 ? 3 : 1;
  const int Kout = p.K * segs;
  const size_t total = p.M * (size_t)Kout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / Kout;
    const int kk = (int)(i % Kout);
    const int seg = kk / p.K, k = kk % p.K;
    const int c = k % p.C, tap = k / p.C;
    const int r = tap / p.kw, s = tap % p.kw;
    const int ow = (int)(m % p.OW);
    const size_t t = m / p.OW;
    const int oh = (int)(t % p.OH), n = (int)(t / p.OH);
    int ih, iw;
    int v = 0;
    if (tap_coord(p, oh, ow, r, s, ih, iw)) {
      const size_t si = (((size_t)n * p.H + ih) * p.W + iw) * p.C + c;
      if (p.src_kind == LBT_MANT_S16) {
        const int kv = reinterpret_cast<const int16_t*>(p.src)[si];
        v = seg < 2 ? (kv >> 1) : (kv & 1);
      } else {
        v = reinterpret_cast<const uint8_t*>(p.src)[si];
      }
    }
    p.out[m * p.ld + kk] = (uint8_t)(v & 0xff);
  }
}

// out[c, r] = in[r, c] for byte matrices, 64x64 tiles through shared memory (both sides coalesced).
__global__ void __launch_bounds__(256) transpose_i8_kernel(const uint8_t* __restrict__ in, size_t R, size_t C, size_t ld_in,
                                                           uint8_t* __restrict__ out, size_t ld_out) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  __shared__ uint8_t tile[64][64 + 4];
  const size_t tiles_c = (C + 63) / 64, tiles_r = (R + 63) / 64;
  for (size_t t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const size_t r0 = (t / tiles_c) * 64, c0 = (t % tiles_c) * 64;
    // load: 64 rows x 64 bytes; thread -> (row = tid/4 + 0, 16-byte chunk = tid%4) when aligned, else bytes
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
      const int rr = i >> 6, cc = i & 63;
      tile[rr][cc] = (r0 + rr < R && c0 + cc < C) ? in[(r0 + rr) * ld_in + c0 + cc] : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
      const int cc = i >> 6, rr = i & 63;
      if (c0 + cc < C && r0 + rr < R) out[(c0 + cc) * ld_out + r0 + rr] = tile[rr][cc];
    }
    __syncthreads();
  }
}

// out[c] = 2^e * sum_r in[r, c] (s8 or s16 mantissas), exact integer sum.  Bias gradient
// tf.gradients(y, b, gradq) = reduce_sum(gradq, [N, H, W]) (dynamic_fixed_point.py:304, 459).
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ in, size_t R, size_t C, const int32_t* ib,
                                                     int exp_const, long long* __restrict__ acc) {
  // grid.x over column blocks of 32, grid.y over row slabs; one warp-wide coalesced row segment per step
  const size_t c = (size_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int rlane = threadIdx.x >> 5, rstep = (blockDim.x >> 5) * gridDim.y;
  long long s = 0;
  if (c < C)
    for (size_t r = (size_t)blockIdx.y * (blockDim.x >> 5) + rlane; r < R; r += rstep) s += (long long)in[r * C + c];
  __shared__ long long sm[8][32];
  sm[rlane][threadIdx.x & 31] = s;
  __syncthreads();
  if (rlane == 0 && c < C) {
    for (int i = 1; i < 8; ++i) s += sm[i][threadIdx.x & 31];
    if (s) atomicAdd(reinterpret_cast<unsigned long long*>(acc) + c, (unsigned long long)s);
  }
}

}  // namespace
}  // namespace lbt

using namespace lbt;

extern "C" int lbt_im2col_i8(const void* src, int src_kind, int N, int H, int W, int C, int OH, int OW, int kh, int kw,
                             int sh, int sw, int pad_top, int pad_left, int transposed, void* out, size_t ld, void* stream) {
  if (!src || !out) return LBT_EINVAL;
  if (src_kind != LBT_MANT_S8 && src_kind != LBT_MANT_U8 && src_kind != LBT_MANT_S16) return LBT_EINVAL;
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || OH <= 0 || OW <= 0 || kh <= 0 || kw <= 0 || sh <= 0 || sw <= 0) return LBT_EINVAL;
  const int segs = src_kind == LBT_MANT_S16 ? 3 : 1;
  const size_t K = (size_t)kh * kw * C;
  if (ld < K * segs) return LBT_EINVAL;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  Im2colParams p{};
  p.src = src;
  p.src_kind = src_kind;
  p.N = N; p.H = H; p.W = W; p.C = C;
  p.OH = OH; p.OW = OW;
  p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.pt = pad_top; p.pl = pad_left;
  p.transposed = transposed;
  p.out = reinterpret_cast<uint8_t*>(out);
  p.ld = ld;
  p.K = (int)K;
  p.M = (size_t)N * OH * OW;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = segs == 1 && (C % 16 == 0) && (ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && p.M * (K / 16) < (1ull << 31);
  const size_t work = vec ? p.M * (K / 16) : p.M * K * segs;
  const size_t blocks = (work + 255) / 256;
  const size_t cap = (size_t)di.sm_count * 16;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  if (vec)
    {
      p.d_kvecs = make_fastdiv((uint32_t)p.K >> 4);
      p.d_cvecs = make_fastdiv((uint32_t)p.C >> 4);
      p.d_kw = make_fastdiv((uint32_t)p.kw);
      p.d_OW = make_fastdiv((uint32_t)p.OW);
      p.d_OH = make_fastdiv((uint32_t)p.OH);
      p.d_sh = make_fastdiv((uint32_t)p.sh);
      p.d_sw = make_fastdiv((uint32_t)p.sw);
        im2col_vec16_kernel<<<grid, 256, 0, st>>>(p);
    }
  else
    im2col_scalar_kernel<<<grid, 256, 0, st>>>(p);
  return check_launch("lbt_im2col_i8");
}

extern "C" int lbt_transpose_i8(const void* in, size_t R, size_t C, size_t ld_in, void* out, size_t ld_out, void* stream) {
  if (!in || !out) return LBT_EINVAL;
  if (R == 0 || C == 0) return LBT_OK;
  if (ld_in < C || ld_out < R) return LBT_EINVAL;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  const size_t tiles = ((R + 63) / 64) * ((C + 63) / 64);
  const size_t cap = (size_t)di.sm_count * 8;
  launch_pdl(transpose_i8_kernel, (unsigned)(tiles < cap ? tiles : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint8_t*>(in), R, C, ld_in, reinterpret_cast<uint8_t*>(out), ld_out);
  return check_launch("lbt_transpose_i8");
}

extern "C" int lbt_colsum_i(const void* in, int kind, size_t R, size_t C, int64_t* acc64, void* stream) {
  if (!in || !acc64) return LBT_EINVAL;
  if (kind != LBT_MANT_S8 && kind != LBT_MANT_S16) return LBT_EINVAL;
  if (R == 0 || C == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  dim3 grid((unsigned)((C + 31) / 32), 1);
  size_t slabs = (R + 255) / 256;
  const size_t want = (size_t)di.sm_count * 4 / grid.x + 1;
  grid.y = (unsigned)(slabs < want ? (slabs ? slabs : 1) : want);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (kind == LBT_MANT_S8)
    colsum_kernel<int8_t><<<grid, 256, 0, st>>>(reinterpret_cast<const int8_t*>(in), R, C, nullptr, 0,
                                                 reinterpret_cast<long long*>(acc64));
  else
    colsum_kernel<int16_t><<<grid, 256, 0, st>>>(reinterpret_cast<const int16_t*>(in), R, C, nullptr, 0,
                                                  reinterpret_cast<long long*>(acc64));
  return check_launch("lbt_colsum_i");
}

// ---- 16-bit mantissas as two 8-bit tensor-core operands: k = 256 * hi + lo, hi = k >> 8 (s8), lo = k & 255 (u8) ----
namespace lbt {
namespace {
__global__ void __launch_bounds__(256) split_s16_kernel(const int16_t* __restrict__ in, size_t n, int8_t* __restrict__ hi,
                                                        uint8_t* __restrict__ lo) {
  const size_t nv = n / 8;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nv; i += (size_t)gridDim.x * 256) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t h[2] = {0, 0}, l[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t k = (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu;
      h[j >> 2] |= (k >> 8) << ((j & 3) * 8);
      l[j >> 2] |= (k & 0xffu) << ((j & 3) * 8);
    }
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(h[0], h[1]);
    reinterpret_cast<uint2*>(lo)[i] = make_uint2(l[0], l[1]);
  }
  for (size_t i = nv * 8 + (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const int k = in[i];
    hi[i] = (int8_t)(k >> 8);
    lo[i] = (uint8_t)(k & 0xff);
  }
}
}  // namespace
}  // namespace lbt

extern "C" int lbt_split_s16(const int16_t* in, size_t n, int8_t* hi, uint8_t* lo, void* stream) {
  if (!in || !hi || !lo) return LBT_EINVAL;
  if (n == 0) return LBT_OK;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(hi) & 7) || (reinterpret_cast<uintptr_t>(lo) & 7))
    return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  const size_t blocks = (n / 8 + 255) / 256 + 1;
  const size_t cap = (size_t)lbt::device_info().sm_count * 8;
  lbt::split_s16_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, n, hi, lo);
  return lbt::check_launch("lbt_split_s16");
}

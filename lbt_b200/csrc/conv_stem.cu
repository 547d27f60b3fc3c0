// Weight gradient of a FIRST-LAYER convolution on the 3-channel 9-bit image, stride 2 (the 7x7/2 ImageNet stem) on B200.
//
// tf.gradients(y, W, gradq) (dynamic_fixed_point.py:207, 302) for the layer whose input is the image: dW[r, s, c, co] = sum over
// output pixels (n, oh, ow) of k[n, 2 oh + r - pt, 2 ow + s - pl, c] * g[n, oh, ow, co], k = 2 * hi + lo the 9-bit mantissa.
//
// The general implicit-GEMM wgrad (conv_i8.cu) runs this layer on the 16-byte LBT_MANT_S9C3 pixels: 49 taps x 16 pseudo-channels
// = 784 GEMM rows of which 147 are real (hi is stored twice for fprop), gathered by im2col-mode TMA as 16-byte rows at a 32-byte
// stride, with the gradient block re-read once per 128-row tile: 946 us on ResNet-18 (ncu: tensor pipe 12 %, L2 54 %), the
// largest kernel of the step and fully exposed at its end.  Here:
//   * lbt_stem_pack8 rewrites the image once per step as 8-byte pixels {hi0, hi1, hi2, 0, lo0, lo1, lo2, 0} with a zero margin of
//     pad_left pixels before and >= 6 behind every row ([N, H, Wp, 8], Wp = 2 OW + 6);
//   * for a filter row r, the 8 x 8 = 64 bytes that output pixel ow multiplies start at byte 16 ow of input row 2 oh + r - pt: a
//     TILED tensor map whose "ow" dimension has a 16-byte stride over 64-byte rows (overlapping windows) delivers, in ONE load per
//     filter row, the [8 oh x 16 ow] x 64 B block that is the MN-major A operand of 128 pixels x (8 taps x 8 bytes); rows above /
//     below the image come back as zeros from the TMA unit (input rows split as (parity, half-row) so that stride 2 is a plain
//     coordinate); left / right padding is the physical margin;
//   * a CTA keeps ALL FOUR 128-row accumulator tiles (2 filter rows each; 4 x Cout tensor-memory columns) across its whole share
//     of the patches: the gradient block is loaded once per patch, 64 KB + 8 KB per 128 pixels instead of 114 KB + 57 KB, 16 MMAs
//     instead of 28, and the int64 atomics run once per CTA.
// acc8[((r * 8 + s) * 8 + b), co] (int64, zeroed by the caller) receives the sums of byte b of tap (r, s); the caller combines
// dW[r, s, c] = 2 * acc8[r, s, c] + acc8[r, s, 4 + c].  Exact (s32 per CTA, <= 65536 pixels; int64 across CTAs).
#include <atomic>

#include "conv_internal.h"
#include "tcgen05.cuh"

namespace lbt {
namespace {

using namespace tc;

constexpr int kPix = 128;          // pixels per patch: 8 output rows x 16 output columns
constexpr int kPatchOH = 8, kPatchOW = 16;
constexpr int kRowBytes = 64;      // 8 input pixels x 8 bytes under one output pixel and one filter row
constexpr int kBlock = kPix * kRowBytes;   // one [128 pixels][64 B] operand block
constexpr int kTapRows = 8;        // filter rows carried (kh <= 8); two per 128-row accumulator tile
constexpr int kStemStages = 3;
constexpr int kStemThreads = 192;  // producer, MMA issuer, 4 epilogue warps

__device__ int g_stem_error = 0;

struct StemParams {
  uint32_t patches, per_cta;       // all patches; contiguous share of one CTA
  uint32_t tiles_x, tiles_img;     // patches per output row block / per image
  FastDiv d_tiles_x, d_tiles_img;
  int pt;
  uint32_t kh, kw, N;              // filter rows / columns, output channels
  long long* acc8;
  int alpha;                       // acc8 += alpha * sum (256 | 1: the byte planes of a 16-bit gradient)
  uint32_t idesc;
};

__device__ __forceinline__ void tma_load_tiled_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_tiled_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// image as 8-byte pixels with zero margins: one thread per output pixel (margins included)
__global__ void __launch_bounds__(256) stem_pack8_kernel(const uint4* __restrict__ x16, int H, int W, int Wp, int ml, size_t total,
                                                         uint2* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const size_t row = i / (size_t)Wp;   // n * H + h
    const int wp = (int)(i - row * (size_t)Wp), w = wp - ml;
    uint2 o = make_uint2(0u, 0u);
    if (w >= 0 && w < W) {
      const uint4 v = __ldcs(x16 + row * (size_t)W + w);   // {hi0 hi1 hi2 hi0 | hi1 hi2 lo0 lo1 | lo2 0 0 0 | 0}
      o.x = v.x & 0x00ffffffu;
      o.y = (v.y >> 16) | ((v.z & 0xffu) << 16);
    }
    out[i] = o;
  }
}

template <int BN>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const StemParams p) {
  constexpr int kStageBytes = kTapRows * kBlock + kBlock;   // 8 filter-row blocks + the gradient block (BN <= 64: 64-byte rows)
  constexpr int kTmemCols = 4 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kStemStages];
  __shared__ __align__(8) uint64_t empty_bar[kStemStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStemStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    s_abort = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
  pdl_trigger();
  fence_before();
  __syncthreads();
  fence_after();
  pdl_wait();   // everything above touched only shared / tensor memory and kernel parameters
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  volatile int* abort_flag = &s_abort;

  const uint32_t p0 = blockIdx.x * p.per_cta, p1 = min(p0 + p.per_cta, p.patches);
  const bool has_work = p0 < p1;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      for (uint32_t pa = p0; pa < p1 && ok; ++pa) {
        if (!(ok = mbar_wait(&empty_bar[stage], phase ^ 1, abort_flag, &g_stem_error))) break;
        uint8_t* sa = smem + (size_t)stage * kStageBytes;
        uint8_t* sb = sa + kTapRows * kBlock;
        const uint32_t img = fastdiv(pa, p.d_tiles_img), t2 = pa - img * p.tiles_img;
        const uint32_t ty = fastdiv(t2, p.d_tiles_x), tx = t2 - ty * p.tiles_x;
        const int oh0 = (int)ty * kPatchOH, ow0 = (int)tx * kPatchOW;
        mbar_expect_tx(&full_bar[stage], (p.kh + 1) * (uint32_t)kBlock);
        for (uint32_t r = 0; r < p.kh; ++r) {
          const int d = (int)r - p.pt;   // input row = 2 oh + d = 2 (oh + floor(d / 2)) + (d & 1)
          tma_load_tiled_5d(&tmX, &full_bar[stage], sa + r * kBlock, 0, ow0, d & 1, oh0 + (d >> 1), (int)img);
        }
        tma_load_tiled_4d(&tmG, &full_bar[stage], sb, 0, ow0, oh0, (int)img);
        if (++stage == kStemStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && has_work) {
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      for (uint32_t pa = p0; pa < p1 && ok; ++pa) {
        if (!(ok = mbar_wait(&full_bar[stage], phase, abort_flag, &g_stem_error))) break;
        fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * kStageBytes);
        const uint32_t sb = sa + kTapRows * kBlock;
#pragma unroll
        for (uint32_t t = 0; t < 4; ++t) {
          if (2 * t >= p.kh) break;
#pragma unroll
          for (uint32_t kk = 0; kk < kPix / 32; ++kk) {
            // 32 pixels (K) per instruction: 32 rows of 64 bytes further down both blocks; the tile's second filter row is the
            // next block (group stride kBlock)
            umma_i8(tmem_base + t * BN, make_desc_mnmajor(sa + 2 * t * kBlock + kk * 32 * kRowBytes, 2, kBlock),
                    make_desc_mnmajor(sb + kk * 32 * kRowBytes, 2, kBlock), p.idesc, (pa > p0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kStemStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (ok) umma_commit(&done_bar);
    }
  } else if (has_work) {
    const uint32_t quad = warp & 3;
    bool ok = mbar_wait(&done_bar, 0, abort_flag, &g_stem_error);
    ok = __all_sync(0xffffffffu, ok);
    if (ok) {
      fence_after();
      const uint32_t m = quad * 32 + lane;          // row of the tile: (filter row & 1, column s, byte b)
      const uint32_t b = m & 7u, s = (m >> 3) & 7u;
      for (uint32_t t = 0; t < 4; ++t) {
        const uint32_t r = 2 * t + (m >> 6);
        const bool real = r < p.kh && s < p.kw && (b & 3u) != 3u;   // bytes 3 and 7 of a pixel are padding
        const uint32_t taddr = tmem_base + t * BN + ((quad * 32u) << 16);
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
          if (real && (uint32_t)c < p.N) {
            unsigned long long* o = reinterpret_cast<unsigned long long*>(p.acc8) + (size_t)(t * 128 + m) * p.N + c;
            const uint32_t ncol = min(16u, p.N - (uint32_t)c);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < (int)ncol && v[j] != 0u) atomicAdd(o + j, (unsigned long long)((long long)(int)v[j] * (long long)p.alpha));
          }
        }
      }
    }
  }

  fence_before();
  __syncthreads();
  fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* stem_driver_fn(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return f;
}

}  // namespace
}  // namespace lbt

using namespace lbt;

extern "C" size_t lbt_stem_pack8_bytes(int N, int H, int OW) {
  if (N <= 0 || H <= 0 || OW <= 0) return 0;
  return (size_t)N * H * (size_t)(2 * OW + 6) * 8;
}

extern "C" int lbt_stem_pack8(const int8_t* x16, int N, int H, int W, int OW, int pad_left, int8_t* work8, void* stream) {
  if (!x16 || !work8) return LBT_EINVAL;
  if (N <= 0 || H <= 0 || W <= 0 || OW <= 0 || pad_left < 0) return LBT_EINVAL;
  const int Wp = 2 * OW + 6;
  if (pad_left + W > Wp) return LBT_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x16) & 15) || (reinterpret_cast<uintptr_t>(work8) & 15)) return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  const size_t total = (size_t)N * H * Wp;
  const size_t blocks = (total + 255) / 256, cap = (size_t)device_info().sm_count * 16;
  launch_pdl(stem_pack8_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream),
             reinterpret_cast<const uint4*>(x16), H, W, Wp, pad_left, total, reinterpret_cast<uint2*>(work8));
  return check_launch("lbt_stem_pack8");
}

extern "C" int lbt_conv_i8_wgrad_c3(const int8_t* x16, int N, int H, int W, const void* g, int g_kind, int Cout, int kh, int kw,
                                    int pad_top, int pad_left, int OH, int OW, int8_t* work8, int repack, int64_t* acc8, int alpha,
                                    void* stream) {
  if (!x16 || !g || !work8 || !acc8) return LBT_EINVAL;
  if (g_kind != LBT_MANT_S8 && g_kind != LBT_MANT_U8) return LBT_EINVAL;
  if (N <= 0 || H <= 0 || W <= 0 || Cout <= 0 || kh <= 0 || kw <= 0 || OH <= 0 || OW <= 0 || pad_top < 0 || pad_left < 0) return LBT_EINVAL;
  // shapes: stride 2 (implied), <= 8 x 8 taps, 64 output channels, an even number of input rows (parity split), and every
  // window inside the padded row
  if (kh > 8 || kw > 8 || Cout != 64 || (H & 1)) return LBT_EUNSUPPORTED;
  const int Wp = 2 * OW + 6;
  if (pad_left + W > Wp || 2 * (OW - 1) + kw > Wp) return LBT_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x16) & 15) || (reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(work8) & 15))
    return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  static EncodeTiledFn enc_tiled = reinterpret_cast<EncodeTiledFn>(stem_driver_fn("cuTensorMapEncodeTiled"));
  if (!enc_tiled) return LBT_ECUDA;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  StemParams p{};
  p.tiles_x = (uint32_t)((OW + kPatchOW - 1) / kPatchOW);
  const uint32_t tiles_y = (uint32_t)((OH + kPatchOH - 1) / kPatchOH);
  p.tiles_img = p.tiles_x * tiles_y;
  if ((uint64_t)N * p.tiles_img >= (1ull << 31)) return LBT_EUNSUPPORTED;
  p.patches = (uint32_t)N * p.tiles_img;
  const uint32_t ctas = (uint32_t)di.sm_count;
  p.per_cta = (p.patches + ctas - 1) / ctas;
  if (p.per_cta > 65536 / kPix) return LBT_EUNSUPPORTED;   // s32 exactness bound per CTA
  p.d_tiles_x = make_fastdiv(p.tiles_x);
  p.d_tiles_img = make_fastdiv(p.tiles_img);
  p.pt = pad_top;
  p.kh = (uint32_t)kh;
  p.kw = (uint32_t)kw;
  p.N = (uint32_t)Cout;
  p.acc8 = reinterpret_cast<long long*>(acc8);
  p.alpha = alpha;
  p.idesc = tc::make_idesc_i8(true, g_kind == LBT_MANT_S8, true, true, 64, 128);

  if (repack) {  // the 8-byte image (repack == 0: work8 already holds it — lbt_stem_pack8 or the previous call of this step)
    int rc = lbt_stem_pack8(x16, N, H, W, OW, pad_left, work8, stream);
    if (rc) return rc;
  }
  CUtensorMap tx, tg;
  const size_t pitch = (size_t)Wp * 8;
  {
    // (64 bytes under an output pixel) x (ow, 16-byte stride: overlapping windows) x (row parity) x (half row) x (image)
    cuuint64_t gdim[5] = {64, (cuuint64_t)OW, 2, (cuuint64_t)(H / 2), (cuuint64_t)N};
    cuuint64_t gstr[4] = {16, (cuuint64_t)pitch, (cuuint64_t)(2 * pitch), (cuuint64_t)((size_t)H * pitch)};
    cuuint32_t box[5] = {64, (cuuint32_t)kPatchOW, 1, (cuuint32_t)kPatchOH, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc_tiled(&tx, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, work8, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LBT_EUNSUPPORTED;   // the caller falls back to the general kernel
  }
  {
    cuuint64_t gdim[4] = {64, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t gstr[3] = {64, (cuuint64_t)OW * 64, (cuuint64_t)OH * OW * 64};
    cuuint32_t box[4] = {64, (cuuint32_t)kPatchOW, (cuuint32_t)kPatchOH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc_tiled(&tg, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(g), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LBT_EUNSUPPORTED;
  }
  constexpr size_t smem = (size_t)kStemStages * (kTapRows + 1) * kBlock + 1024;
  static bool attr_done[16] = {};
  if (!attr_done[di.device]) {
    cudaError_t e = cudaFuncSetAttribute(stem_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_cuda_error(e, "cudaFuncSetAttribute(stem_wgrad_kernel)");
      return LBT_ECUDA;
    }
    attr_done[di.device] = true;
  }
  const unsigned grid = (p.patches + p.per_cta - 1) / p.per_cta;
  launch_pdl(stem_wgrad_kernel<64>, grid, kStemThreads, smem, st, tx, tg, p);
  return check_launch("lbt_conv_i8_wgrad_c3");
}

namespace lbt {
int conv_stem_debug_error() {
  int v = 0, zero = 0;
  if (cudaMemcpyFromSymbol(&v, g_stem_error, sizeof(int)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_stem_error, &zero, sizeof(int));
  return v;
}
}  // namespace lbt

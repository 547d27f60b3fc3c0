// Implicit-GEMM convolution for NARROW channel counts (C in {16, 32, 64}, Cout <= 128) on B200 (sm_100a).
//
// Measured on B200 (benchmarks/tma_probe.cu): the TMA engine retires one [pixel x C] row every ~3 clocks per SM,
// im2col mode or tiled, whatever the row width — 5 B/clk/SM at C = 16, eight times below what the HBM-bound
// first-stage convolutions of the CIFAR / ImageNet ResNets need.  Here the A operand (the im2col rows) is gathered
// by four loader warps with 16-byte cp.async (512 B per warp instruction, zero-fill for the padding), straight into
// the tensor core's un-swizzled K-major layout; the whole filter bank stays resident in shared memory for the life
// of the persistent CTA; tcgen05.mma.kind::i8 accumulates in tensor memory and the epilogue is the same as the
// TMA-fed kernel's (fp32 rescale, or the fused re-quantise + batch statistics of qsite.cuh).
//
//   gather == 0  fprop:  row (n, oh, ow), tap (r, s) reads src[n, oh*sh - pt + r, ow*sw - pl + s, :]
//                         (tf.nn.conv2d, dynamic_fixed_point.py:291; with a rot180 filter also the stride-1 dgrad)
//   gather == 1  dgrad:  row (n, h, w) of dX, tap (r, s) reads g[n, (h + pt - r)/sh, (w + pl - s)/sw, :] where
//                         divisible (tf.gradients(y, X, gradq), dynamic_fixed_point.py:305, any stride) — the
//                         strided transposed convolution without an im2col matrix in HBM.
//
// Shared-memory layouts (SWIZZLE_NONE, "interleaved" 8x16B core matrices):
//   A stage : [8 K-chunks][128 rows][16 B]   chunk kc of the tile = (tap, 16-channel group), kc = tap * (C/16) + cc
//   B       : [KCp K-chunks][BN rows][16 B]  resident; KCp = chunk count rounded up to even (zero chunk)
// One K=32 MMA covers two consecutive chunks (leading byte offset = chunk stride).
#include <atomic>

#include "conv_internal.h"
#include "qsite.cuh"
#include "tcgen05.cuh"

namespace lbt {
namespace {

using namespace tc;

constexpr int kBlockM = 128;
constexpr int kLoaderWarps = 4;
constexpr int kThreads = 32 * (kLoaderWarps + 1 + 4);  // loaders, MMA issuer, epilogue
constexpr int kChunksPerStage = 8;
constexpr int kStageBytes = kChunksPerStage * kBlockM * 16;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kMaxAccStages = 4;

__device__ int g_ldg_error = 0;

struct LdgParams {
  const uint8_t* src;            // NHWC mantissas of the gathered tensor
  const uint8_t* wp;             // packed filter [Cout, taps*C] K-major
  uint32_t ldw;
  uint32_t M, N;                 // output rows (pixels), output channels
  uint32_t SH, SW, C;            // gathered tensor: height, width, channels
  uint32_t OW, OHW;              // output grid width, height*width
  int sh, sw, pt, pl;
  uint32_t kw, taps, cpp;        // filter width, kh*kw, C/16
  uint32_t KC, KCp;              // 16-byte K chunks (real, padded to even)
  uint32_t stages_per_tile;      // ceil(KCp / 8)
  uint32_t nstages;              // ring depth
  uint32_t m_tiles;
  int gather;                    // 0 fprop, 1 transposed (dgrad)
  const int32_t* ibA;
  const int32_t* ibB;
  int exp_const;
  const float* bias;
  float* out;
  size_t ldc;
  uint32_t idesc;
  BnqParams bnq;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int BN>
__global__ void __launch_bounds__(kThreads, 2) conv_ldg_kernel(const LdgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  constexpr int kAccStages = BN <= 64 ? 4 : 2;   // two CTAs per SM must fit the 512 TMEM columns
  __shared__ __align__(8) uint64_t tmem_full_bar[kMaxAccStages];
  __shared__ __align__(8) uint64_t tmem_empty_bar[kMaxAccStages];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort;
  __shared__ int s_stat[4][2 * BN];

  constexpr int kTmemCols = (kAccStages * BN) < 32 ? 32 : (kAccStages * BN);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sB = smem;                                        // KCp * BN * 16
  uint8_t* sA = smem + (((size_t)p.KCp * BN * 16 + 127) & ~(size_t)127);

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < p.nstages; ++s) {
      mbar_init(&full_bar[s], kLoaderWarps * 32);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);
    }
    s_abort = 0;
    fence_barrier_init();
  }
  if (warp == kLoaderWarps) tmem_alloc(&tmem_slot, kTmemCols);
  // resident filter bank: chunk kc, output channel n -> 16 bytes (zeros beyond Cout / KC)
  for (uint32_t i = threadIdx.x; i < p.KCp * BN; i += kThreads) {
    const uint32_t kc = i / BN, n = i % BN;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (kc < p.KC && n < p.N) v = __ldg(reinterpret_cast<const uint4*>(p.wp + (size_t)n * p.ldw + (size_t)kc * 16));
    *reinterpret_cast<uint4*>(sB + (size_t)i * 16) = v;
  }
  fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  volatile int* abort_flag = &s_abort;

  if (warp < kLoaderWarps) {
    // ===== loaders: thread t gathers row t of every tile =====
    const uint32_t row = threadIdx.x;
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    for (uint32_t tile = blockIdx.x; tile < p.m_tiles && ok; tile += gridDim.x) {
      const uint32_t m = tile * kBlockM + row;
      const bool row_ok = m < p.M;
      const uint32_t img = m / p.OHW, rem = m % p.OHW;
      const int oy = (int)(rem / p.OW), ox = (int)(rem % p.OW);
      const uint8_t* img_base = p.src + (size_t)img * p.SH * p.SW * p.C;
      uint32_t tap = 0, cc = 0;  // decomposition of the running chunk index
      int r = 0, s = 0;
      for (uint32_t st = 0; st < p.stages_per_tile; ++st) {
        ok = mbar_wait(&empty_bar[stage], phase ^ 1, abort_flag, &g_ldg_error);
        if (!ok) break;
        const uint32_t dst0 = smem_u32(sA + (size_t)stage * kStageBytes) + row * 16;
#pragma unroll
        for (int j = 0; j < kChunksPerStage; ++j) {
          const uint32_t kc = st * kChunksPerStage + j;
          if (kc < p.KCp) {  // uniform
            bool v = row_ok && kc < p.KC;
            int iy, ix;
            if (p.gather == 0) {
              iy = oy * p.sh - p.pt + r;
              ix = ox * p.sw - p.pl + s;
            } else {
              const int ty = oy + p.pt - r, tx = ox + p.pl - s;
              iy = ty / p.sh;
              ix = tx / p.sw;
              v = v && ty >= 0 && tx >= 0 && (ty - iy * p.sh) == 0 && (tx - ix * p.sw) == 0;
            }
            v = v && iy >= 0 && ix >= 0 && iy < (int)p.SH && ix < (int)p.SW;
            const uint8_t* src = v ? img_base + ((size_t)iy * p.SW + ix) * p.C + cc * 16 : p.src;
            cp_async16(dst0 + j * (kBlockM * 16), src, v ? 16u : 0u);
            if (++cc == p.cpp) {
              cc = 0;
              ++tap;
              if (++s == (int)p.kw) {
                s = 0;
                ++r;
              }
            }
          }
        }
        cp_async_arrive_noinc(&full_bar[stage]);
        if (++stage == p.nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == kLoaderWarps) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ok = true;
      const uint32_t sb0 = smem_u32(sB);
      for (uint32_t tile = blockIdx.x; tile < p.m_tiles && ok; tile += gridDim.x) {
        if (!(ok = mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, abort_flag, &g_ldg_error))) break;
        fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t first = 1;
        for (uint32_t st = 0; st < p.stages_per_tile; ++st) {
          if (!(ok = mbar_wait(&full_bar[stage], phase, abort_flag, &g_ldg_error))) break;
          fence_after();
          const uint32_t sa = smem_u32(sA + (size_t)stage * kStageBytes);
          const uint32_t kc0 = st * kChunksPerStage;
          const uint32_t n2 = min((uint32_t)kChunksPerStage, p.KCp - kc0) >> 1;
          for (uint32_t j = 0; j < n2; ++j) {
            umma_i8(d_tmem, make_desc_kmajor(sa + 2 * j * (kBlockM * 16), 0, kBlockM * 16),
                    make_desc_kmajor(sb0 + (kc0 + 2 * j) * (BN * 16), 0, BN * 16), p.idesc, first ? 0u : 1u);
            first = 0;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (!ok) break;
        umma_commit(&tmem_full_bar[acc]);
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===== epilogue: TMEM lane quadrant = warp % 4 =====
    const uint32_t quad = warp & 3;
    int e = p.exp_const;
    if (p.ibA) e += *p.ibA;
    if (p.ibB) e += *p.ibB;
    const float scale = exp2i(e);
    const bool fused = p.bnq.q.bits != 0;
    int* my_stat = s_stat[quad];
    BnqState bst;
    bst.tiles = 0;
    if (fused) {
      bst.init(p.bnq);
      for (int i = lane; i < 2 * BN; i += 32) my_stat[i] = 0;
      __syncwarp();
    }
    uint32_t acc = 0, acc_phase = 0;
    bool ok = true;
    for (uint32_t tile = blockIdx.x; tile < p.m_tiles && ok; tile += gridDim.x) {
      ok = mbar_wait(&tmem_full_bar[acc], acc_phase, abort_flag, &g_ldg_error);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      fence_after();
      const uint32_t row = tile * kBlockM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + ((quad * 32u) << 16);
      if (fused && bst.tiles >= (uint32_t)kBnqFlushTiles) {
        bnq_flush(p.bnq, my_stat, 0, BN, p.N, lane);
        bst.tiles = 0;
      }
      ++bst.tiles;
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        if ((uint32_t)c >= p.N) continue;  // warp-uniform
        const uint32_t ncol = min(16u, p.N - (uint32_t)c);
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          f[j] = __int2float_rn((int)v[j]) * scale;
          if (p.bias && j < (int)ncol) f[j] = __fadd_rn(f[j], __ldg(p.bias + c + j));
        }
        if (fused) {
          bnq_chunk(p.bnq, bst, f, row, row < p.M, (uint32_t)c, ncol, p.N, my_stat, BN, (uint32_t)c, lane);
        } else if (row < p.M) {
          float* o = p.out + (size_t)row * p.ldc + c;
          if (ncol == 16 && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < (int)ncol) o[j] = f[j];
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (fused) {
      bnq_flush(p.bnq, my_stat, 0, BN, p.N, lane);
      bnq_finish(p.bnq, bst, (unsigned long long)p.M * p.N, quad == 0, lane);
    }
  }

  fence_before();
  __syncthreads();
  fence_after();
  if (warp == kLoaderWarps) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN>
int launch_ldg(const LdgParams& p, unsigned grid, size_t smem, cudaStream_t st) {
  static size_t attr_smem[16] = {};
  const int dev = device_info().device;
  if (attr_smem[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_ldg_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_cuda_error(e, "cudaFuncSetAttribute(conv_ldg_kernel)");
      return LBT_ECUDA;
    }
    attr_smem[dev] = smem;
  }
  conv_ldg_kernel<BN><<<grid, kThreads, smem, st>>>(p);
  return check_launch("lbt_conv_i8 (cp.async gather)");
}

}  // namespace

std::atomic<int> g_use_ldg{1};
bool conv_ldg_enabled() { return g_use_ldg.load(std::memory_order_relaxed) != 0; }

// Shapes this kernel takes.
bool conv_ldg_ok(int C, int Cout, int kh, int kw) {
  if (C != 16 && C != 32 && C != 64) return false;
  if (Cout < 1 || Cout > 128) return false;
  const size_t kc = (size_t)kh * kw * (C / 16);
  const size_t kcp = (kc + 1) & ~(size_t)1;
  const int bn = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128));
  return kcp * bn * 16 + 2 * kStageBytes + 1024 <= 200 * 1024 && (size_t)kh * kw * C <= 65536;
}

// Shared by lbt_conv_i8_fprop (gather 0) and lbt_conv_i8_dgrad (gather 1).  (M rows) x (Cout columns); the gathered
// tensor is src[N, SH, SW, C]; the output grid is OH x OW per image.
int conv_ldg_run(const void* src, int src_kind, int N, int SH, int SW, int C, const void* wp, int w_kind, size_t ldw, int Cout,
                 int kh, int kw, int sh, int sw, int pt, int pl, int OH, int OW, int gather, const int32_t* ibA,
                 const int32_t* ibB, int exp_const, const float* bias, float* out, size_t ldc, const lbt_qsite* q_out,
                 int8_t* k_out, int64_t* sums, void* stream) {
  const DeviceInfo& di = device_info();
  const int bn = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128));
  LdgParams p{};
  p.src = reinterpret_cast<const uint8_t*>(src);
  p.wp = reinterpret_cast<const uint8_t*>(wp);
  p.ldw = (uint32_t)ldw;
  p.M = (uint32_t)((size_t)N * OH * OW);
  p.N = (uint32_t)Cout;
  p.SH = (uint32_t)SH;
  p.SW = (uint32_t)SW;
  p.C = (uint32_t)C;
  p.OW = (uint32_t)OW;
  p.OHW = (uint32_t)(OH * OW);
  p.sh = sh;
  p.sw = sw;
  p.pt = pt;
  p.pl = pl;
  p.kw = (uint32_t)kw;
  p.taps = (uint32_t)(kh * kw);
  p.cpp = (uint32_t)C / 16;
  p.KC = p.taps * p.cpp;
  p.KCp = (p.KC + 1) & ~1u;
  p.stages_per_tile = (p.KCp + kChunksPerStage - 1) / kChunksPerStage;
  p.m_tiles = (p.M + kBlockM - 1) / kBlockM;
  p.gather = gather;
  p.ibA = ibA;
  p.ibB = ibB;
  p.exp_const = exp_const;
  p.bias = bias;
  p.out = out;
  p.ldc = ldc;
  p.idesc = tc::make_idesc_i8(src_kind == LBT_MANT_S8, w_kind == LBT_MANT_S8, false, false, bn, kBlockM);
  p.bnq.q = site_from_abi(q_out);
  p.bnq.k = k_out;
  p.bnq.sums = reinterpret_cast<long long*>(sums);
  p.bnq.rows_per_image = (uint32_t)(OH * OW);
  const size_t b_bytes = (((size_t)p.KCp * bn * 16) + 127) & ~(size_t)127;
  // ring depth: two CTAs per SM when the filter bank is small, else one CTA with a deep ring
  size_t budget = (b_bytes <= 24 * 1024 ? 100 * 1024 : 200 * 1024) - b_bytes - 1024;
  uint32_t nst = (uint32_t)(budget / kStageBytes);
  if (nst > (uint32_t)kMaxStages) nst = kMaxStages;
  if (nst < 2) return LBT_EUNSUPPORTED;
  p.nstages = nst;
  const size_t smem = b_bytes + (size_t)nst * kStageBytes + 256;
  const unsigned ctas_per_sm = b_bytes <= 24 * 1024 ? 2u : 1u;
  const uint64_t cap = (uint64_t)di.sm_count * ctas_per_sm;
  const unsigned grid = (unsigned)(p.m_tiles < cap ? p.m_tiles : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 16: return launch_ldg<16>(p, grid, smem, st);
    case 32: return launch_ldg<32>(p, grid, smem, st);
    case 64: return launch_ldg<64>(p, grid, smem, st);
    default: return launch_ldg<128>(p, grid, smem, st);
  }
}

}  // namespace lbt

using namespace lbt;

extern "C" int lbt_conv_i8_dgrad(const void* g, int g_kind, int N, int OH, int OW, int Cout, const void* wp, int w_kind, size_t ldw,
                                 int Cin, int kh, int kw, int sh, int sw, int pad_top, int pad_left, int H, int W,
                                 const int32_t* ib_g, const int32_t* ib_w, int exp_const, float* dx, size_t ldc, void* stream) {
  if (!g || !wp || !dx) return LBT_EINVAL;
  if ((g_kind != LBT_MANT_S8 && g_kind != LBT_MANT_U8) || (w_kind != LBT_MANT_S8 && w_kind != LBT_MANT_U8)) return LBT_EINVAL;
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || kh <= 0 || kw <= 0 || sh <= 0 || sw <= 0 || OH <= 0 || OW <= 0)
    return LBT_EINVAL;
  if (ldc < (size_t)Cin || ldw < (size_t)kh * kw * Cout) return LBT_EINVAL;
  if ((reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(wp) & 15) || (ldw & 15)) return LBT_EUNSUPPORTED;
  if (!conv_ldg_ok(Cout, Cin, kh, kw)) return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  // rows = input pixels (n, h, w); gathered tensor = g[N, OH, OW, Cout]; output channels = Cin
  return conv_ldg_run(g, g_kind, N, OH, OW, Cout, wp, w_kind, ldw, Cin, kh, kw, sh, sw, pad_top, pad_left, H, W, 1, ib_g, ib_w,
                      exp_const, nullptr, dx, ldc, nullptr, nullptr, nullptr, stream);
}

// Test / bench knob (not in lbt.h): 0 routes every convolution through the TMA-im2col kernel, 1 (default) lets the
// narrow-channel shapes take the cp.async-gather kernel.
extern "C" int lbt_conv_set_path(int use_ldg) {
  g_use_ldg.store(use_ldg ? 1 : 0, std::memory_order_relaxed);
  return LBT_OK;
}

extern "C" int lbt_conv_ldg_debug_error(void) {
  int v = 0, zero = 0;
  if (cudaMemcpyFromSymbol(&v, g_ldg_error, sizeof(int)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_ldg_error, &zero, sizeof(int));
  return v;
}

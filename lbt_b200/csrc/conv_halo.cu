// Stride-1 implicit-GEMM convolution for 64- and 128-channel inputs on B200 (sm_100a): the HALO-PATCH loader on the TMA engine.
//
// The im2col-mode kernel of conv_i8.cu moves every input pixel from L2 to shared memory kh*kw times (1 152 TMA rows per
// 128-pixel tile at 3x3) and reloads the filter tile per M tile; measured, it is paced by the TMA engine's ~3 clocks per row
// (ncu: tensor pipe 10 % active on ResNet-18's 64-channel layers).  Here an M tile is an 8-wide x 16-high patch of output
// pixels of ONE image and
//   * ONE tiled 4-D TMA load brings the (16 + kh - 1) x (8 + kw - 1) input patch under it ([halo pixels][C bytes], 64- or
//     128-byte swizzled rows, out-of-image pixels zero-filled by the TMA unit = TF 'SAME' padding): 180 rows instead of 1 152;
//   * the MMA descriptors walk the patch: the 8 rows of a core-matrix group are 8 neighbouring pixels of an output row, the
//     stride-byte-offset is the halo row pitch, and a filter tap (r, s) is just a different START ADDRESS.  tcgen05 applies the
//     swizzle XOR to absolute shared-memory address bits, so a start address that is not aligned to the swizzle's repeating
//     pattern and a row pitch that is not a multiple of it read exactly what the TMA unit wrote (descriptor base-offset field
//     0; checked tap by tap against the CPU by benchmarks/halo_probe.cu on B200 for both swizzle widths);
//   * the whole filter bank (kh*kw*C x Cout <= 144 KB) is loaded once per persistent CTA and stays resident.
// Same fp32 / fused re-quantising epilogues and bit-identical results as the im2col kernel (tests/test_layers_gpu.py).
// Replaces tf.nn.conv2d (dynamic_fixed_point.py:291) and, with the filter rotated by 180 degrees, its stride-1 input
// gradient (:305) for the 3x3 layers of the ImageNet ResNets' first two stages.
#include <atomic>
#include <cstdlib>

#include "conv_internal.h"
#include "qsite.cuh"
#include "tcgen05.cuh"

namespace lbt {
namespace {

using namespace tc;

constexpr int kBlockM = 128;
constexpr int kMaxStages = 6;
constexpr int kPatchW = 8, kPatchH = 16;
constexpr int kMaxMmas = 25 * 4;   // 5 x 5 taps x (128 B / 32 B)

__device__ int g_halo_error = 0;
std::atomic<long long> g_halo_launches{0};

struct HaloParams {
  uint32_t M, N;                 // output pixels, output channels
  uint32_t OH, OW, OHW;
  uint32_t m_tiles;              // images * tiles per image
  FastDiv d_tiles_img, d_tiles_x;
  int pt, pl;
  uint32_t kh, kw;
  uint32_t cb, mode;             // bytes per pixel (= C), log2(cb / 16)
  uint32_t halo_w;               // 8 + kw - 1
  uint32_t stage_bytes;          // one patch, rounded up to 1024
  uint32_t patch_bytes;          // halo pixels * cb: what one TMA load delivers
  uint32_t nstages;
  uint32_t b_block;              // BN * cb: one filter tap
  const int32_t* ibA;
  const int32_t* ibB;
  int exp_const;
  const float* bias;
  const float* addend;
  float* out;
  size_t ldc;
  uint32_t idesc;
  BnqParams bnq;
  int remap;                     // fp32 rows go to out + img * rs_n + oy * rs_y + ox * rs_x (OutRemap) instead of row * ldc
  int sector_f32;                // fp32 epilogue through the .16x256b load shape (full-sector stores): N % 16 == 0, 8-byte aligned rows
  uint32_t idesc2;               // DUAL: instruction descriptor of the low-plane MMAs (A = u8)
  long long rs_n, rs_y, rs_x;
};

__device__ __forceinline__ void tma_load_tiled_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c, int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

// K-major swizzled operand with an explicit stride between 8-row groups (the halo row pitch).
__device__ __forceinline__ uint64_t make_desc_halo(uint32_t smem_addr, int m, uint32_t sbo) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         (desc_layout_bits(m) << 61);
}

// EPI epilogue warps (a multiple of 4: EPI / 4 per TMEM lane quadrant, interleaved over the 16-column chunks).
// FUSED: the re-quantising epilogue (two instantiations per shape: either epilogue compiles without the other's registers).
// DUAL: the source is a 9..16-bit mantissa split into byte planes k = 256 * hi + lo (hi s8 through tmA, lo u8 through tmA2: the
// 16-bit gradients of BASELINE config 5).  Both planes' patches arrive per tile, both are multiplied with the SAME resident filter
// bank into two accumulators (BN columns apart), and the fp32 epilogue rounds 256 * acc_hi + acc_lo ONCE — lbt_gemm_i8_dual's
// arithmetic without the two im2col matrices it needed for a 3x3 layer.  fp32 sector epilogue only (host-checked).
template <int BN, int EPI, bool FUSED, bool DUAL = false>
__global__ void __launch_bounds__(32 * (2 + EPI), EPI == 8 ? 2 : 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const HaloParams p) {
  constexpr int kThreadsH = 32 * (2 + EPI);
  constexpr int kAccW = (DUAL ? 2 : 1) * BN;   // tensor-memory columns of one accumulator stage
  constexpr int kAccStages = DUAL ? (EPI == 8 ? 256 / kAccW : 512 / kAccW) : (EPI == 8 ? (BN <= 64 ? 4 : 2) : (512 / BN > 4 ? 4 : 512 / BN));
  constexpr int kTmemCols = kAccStages * kAccW;
  static_assert(kAccStages >= 1, "tensor memory budget");
  constexpr int kSub = EPI / 4;   // epilogue warps per quadrant
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[kAccStages];
  __shared__ __align__(8) uint64_t tmem_empty_bar[kAccStages];
  __shared__ __align__(8) uint64_t b_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort;
  __shared__ int s_stat[EPI][2 * BN];
  __shared__ unsigned long long s_tot[2 * BN];
  __shared__ uint64_t s_bdesc[kMaxMmas];   // per MMA of a tile: the filter block's descriptor ...
  __shared__ uint32_t s_aoff[kMaxMmas];    // ... and the patch offset (>> 4) of its tap / K slice

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t taps = p.kh * p.kw;
  uint8_t* sB = smem;                                   // [taps][BN rows x cb] swizzled K-major blocks, resident
  uint8_t* sA = smem + (size_t)taps * p.b_block;        // ring of patches (b_block is a multiple of 1024)
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < p.nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], EPI);
    }
    mbar_init(&b_bar, 1);
    s_abort = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    if (DUAL) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
  for (uint32_t i = threadIdx.x; i < 2 * BN; i += kThreadsH) s_tot[i] = 0ull;
  const uint32_t per_tap = p.cb / 32, n_mma = taps * per_tap;
  for (uint32_t j = threadIdx.x; j < n_mma; j += kThreadsH) {   // the single issuing thread then needs two loads and an add per MMA
    const uint32_t tap = j / per_tap, kk = j - tap * per_tap, r = tap / p.kw, sx = tap - r * p.kw;
    s_aoff[j] = ((r * p.halo_w + sx) * p.cb + kk * 32) >> 4;
    s_bdesc[j] = make_desc_kmajor(smem_u32(sB) + tap * p.b_block + kk * 32, (int)p.mode, 16);
  }
  pdl_trigger();
  fence_before();
  __syncthreads();
  fence_after();
  pdl_wait();   // everything above touched only shared / tensor memory and kernel parameters
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  volatile int* abort_flag = &s_abort;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&b_bar, taps * p.b_block);
      for (uint32_t t = 0; t < taps; ++t) tma_load_2d(&tmB, &b_bar, sB + (size_t)t * p.b_block, (int)(t * p.cb), 0);
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      for (uint32_t tile = blockIdx.x; tile < p.m_tiles && ok; tile += gridDim.x) {
        const uint32_t img = fastdiv(tile, p.d_tiles_img), t2 = tile - img * p.d_tiles_img.d;
        const uint32_t ty = fastdiv(t2, p.d_tiles_x), tx = t2 - ty * p.d_tiles_x.d;
        if (!(ok = mbar_wait(&empty_bar[stage], phase ^ 1, abort_flag, &g_halo_error))) break;
        mbar_expect_tx(&full_bar[stage], (DUAL ? 2u : 1u) * p.patch_bytes);
        tma_load_tiled_4d(&tmA, &full_bar[stage], sA + (size_t)stage * p.stage_bytes, 0, (int)(tx * kPatchW) - p.pl,
                          (int)(ty * kPatchH) - p.pt, (int)img);
        if (DUAL)   // the low plane's patch: second half of the stage
          tma_load_tiled_4d(&tmA2, &full_bar[stage], sA + (size_t)stage * p.stage_bytes + p.stage_bytes / 2, 0, (int)(tx * kPatchW) - p.pl,
                            (int)(ty * kPatchH) - p.pt, (int)img);
        if (++stage == p.nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ok = mbar_wait(&b_bar, 0, abort_flag, &g_halo_error);
      const uint32_t sbo = p.halo_w * p.cb;
      for (uint32_t tile = blockIdx.x; tile < p.m_tiles && ok; tile += gridDim.x) {
        if (!(ok = mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, abort_flag, &g_halo_error))) break;
        if (!(ok = mbar_wait(&full_bar[stage], phase, abort_flag, &g_halo_error))) break;
        fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccW;
        const uint32_t sa = smem_u32(sA + (size_t)stage * p.stage_bytes);
        const uint64_t da0 = make_desc_halo(sa, (int)p.mode, sbo);   // + offset >> 4: the start-address field cannot carry out
#pragma unroll 4
        for (uint32_t j = 0; j < n_mma; ++j) umma_i8(d_tmem, da0 + s_aoff[j], s_bdesc[j], p.idesc, j ? 1u : 0u);
        if (DUAL) {   // the low plane against the same filter bank, into the second accumulator
          const uint64_t db0 = make_desc_halo(sa + p.stage_bytes / 2, (int)p.mode, sbo);
#pragma unroll 4
          for (uint32_t j = 0; j < n_mma; ++j) umma_i8(d_tmem + BN, db0 + s_aoff[j], s_bdesc[j], p.idesc2, j ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[acc]);
        if (++stage == p.nstages) {
          stage = 0;
          phase ^= 1;
        }
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const uint32_t quad = warp & 3;
    const uint32_t sub = (uint32_t)(warp - 2) >> 2;   // which of the quadrant's warps: chunks sub, sub + kSub, ...
    int e = p.exp_const;
    if (p.ibA) e += *p.ibA;
    if (p.ibB) e += *p.ibB;
    const float scale = exp2i(e);
    constexpr bool fused = FUSED;   // == (p.bnq.q.bits != 0), chosen by the host
    int* my_stat = s_stat[warp - 2];
    BnqState bst;
    bst.tiles = 0;
    if (fused) {
      bst.init(p.bnq);
      for (int i = lane; i < 2 * BN; i += 32) my_stat[i] = 0;
      __syncwarp();
    }
    // fp32 output in full sectors (see below): needs whole 16-column chunks and 8-byte aligned rows
    const bool sect = !fused && p.sector_f32 != 0;
    uint32_t acc = 0, acc_phase = 0;
    bool ok = true;
    for (uint32_t tile = blockIdx.x; tile < p.m_tiles && ok; tile += gridDim.x) {
      // accumulator row m = 8 * (row of the patch) + column of the patch
      const uint32_t m = quad * 32 + lane;
      const uint32_t img = fastdiv(tile, p.d_tiles_img), t2 = tile - img * p.d_tiles_img.d;
      const uint32_t ty = fastdiv(t2, p.d_tiles_x), tx = t2 - ty * p.d_tiles_x.d;
      const uint32_t oy = ty * kPatchH + (m >> 3), ox = tx * kPatchW + (m & 7);
      const bool rvalid = oy < p.OH && ox < p.OW;
      const uint32_t pix = oy * p.OW + ox;
      const uint32_t row = img * p.OHW + pix;
      // fp32 epilogue: where this row goes (and where its addend comes from)
      const long long obase = fused ? 0ll
                              : (p.remap ? (long long)img * p.rs_n + (long long)oy * p.rs_y + (long long)ox * p.rs_x
                                         : (long long)((size_t)row * p.ldc));
      const bool add_vec = !fused && p.addend != nullptr && rvalid && ((reinterpret_cast<uintptr_t>(p.addend + obase) & 15u) == 0);
      if (fused) {   // this warp's noise lines of the tile: into L1 while the accumulator is still being computed
#pragma unroll 1
        for (int c = 16 * (int)sub; c < BN; c += 16 * kSub) bnq_prefetch(p.bnq, pix, p.N, (uint32_t)c, rvalid && (uint32_t)c < p.N);
      } else if (add_vec) {   // ... and the addend's lines (an fp32 tensor in HBM: a DRAM round trip per chunk otherwise)
#pragma unroll 1
        for (int c = 16 * (int)sub; c < BN; c += 16 * kSub)
          if ((uint32_t)c < p.N) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.addend + obase + c));
      }
      ok = mbar_wait(&tmem_full_bar[acc], acc_phase, abort_flag, &g_halo_error);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      fence_after();
      const uint32_t taddr = tmem_base + acc * kAccW + ((quad * 32u) << 16);
      if (fused && bst.tiles >= (uint32_t)kBnqFlushTiles) {
        bnq_flush(p.bnq, my_stat, 0, BN, p.N, lane);
        bst.tiles = 0;
      }
      ++bst.tiles;
      if (sect) {
        // fp32 rows as full 32-byte sectors: the accumulator-fragment load shape puts 8 contiguous bytes of a row in each lane
        // and a row's 32 bytes in four neighbouring lanes.  This lane's four rows: patch column lane / 4, patch rows
        // 4 * quad + {0, 1, 2, 3} (the 16-lane halves h = 0, 1 of the quadrant x the fragment's row pair k = 0, 1).
        const uint32_t ox2 = tx * kPatchW + ((uint32_t)lane >> 2), oy2 = ty * kPatchH + 4u * quad;
        const long long ystep = p.remap ? (long long)p.rs_y : (long long)((size_t)p.OW * p.ldc);
        const long long ob2 = (p.remap ? (long long)img * p.rs_n + (long long)oy2 * p.rs_y + (long long)ox2 * p.rs_x
                                       : (long long)((size_t)(img * p.OHW + oy2 * p.OW + ox2) * p.ldc)) + 2 * (lane & 3);
        bool rv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rv[j] = ox2 < p.OW && oy2 + (uint32_t)j < p.OH;
#pragma unroll 1
        for (int c = 16 * (int)sub; c < BN; c += 16 * kSub) {
          if ((uint32_t)c >= p.N) break;   // warp-uniform (N % 16 == 0 on this path)
          float2 ad[8];
          if (p.addend) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int n = 0; n < 2; ++n)
                if (rv[j]) ad[2 * j + n] = __ldcs(reinterpret_cast<const float2*>(p.addend + ob2 + j * ystep + c + 8 * n));
          }
          uint32_t va[8], vb[8], wa[8], wb[8];
          tmem_ld_16x256b_x2(taddr + c, va);                 // lanes 0..15 of the quadrant: patch rows 4 * quad + {0, 1}
          tmem_ld_16x256b_x2(taddr + c + (16u << 16), vb);   // lanes 16..31:                             + {2, 3}
          if (DUAL) {
            tmem_ld_16x256b_x2(taddr + BN + c, wa);
            tmem_ld_16x256b_x2(taddr + BN + c + (16u << 16), wb);
          }
          tmem_ld_wait();
          float2 bs[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          if (p.bias) {
            bs[0] = __ldg(reinterpret_cast<const float2*>(p.bias + c + 2 * (lane & 3)));
            bs[1] = __ldg(reinterpret_cast<const float2*>(p.bias + c + 8 + 2 * (lane & 3)));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!rv[j]) continue;
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              const uint32_t* src = j < 2 ? va : vb;
              float f0, f1;
              if (DUAL) {   // 256 * (hi . W) + (lo . W): exact in 64 bits, ONE rounding to fp32 (lbt_gemm_i8_dual's arithmetic)
                const uint32_t* slo = j < 2 ? wa : wb;
                f0 = __ll2float_rn((long long)(int)src[4 * n + 2 * (j & 1)] * 256ll + (long long)(int)slo[4 * n + 2 * (j & 1)]) * scale;
                f1 = __ll2float_rn((long long)(int)src[4 * n + 2 * (j & 1) + 1] * 256ll + (long long)(int)slo[4 * n + 2 * (j & 1) + 1]) * scale;
              } else {
                f0 = __int2float_rn((int)src[4 * n + 2 * (j & 1)]) * scale;
                f1 = __int2float_rn((int)src[4 * n + 2 * (j & 1) + 1]) * scale;
              }
              if (p.bias) {
                f0 = __fadd_rn(f0, n ? bs[1].x : bs[0].x);
                f1 = __fadd_rn(f1, n ? bs[1].y : bs[0].y);
              }
              if (p.addend) {
                f0 = __fadd_rn(f0, ad[2 * j + n].x);
                f1 = __fadd_rn(f1, ad[2 * j + n].y);
              }
              *reinterpret_cast<float2*>(p.out + ob2 + j * ystep + c + 8 * n) = make_float2(f0, f1);
            }
          }
        }
      } else
#pragma unroll 1
      for (int c = 16 * (int)sub; c < BN; c += 16 * kSub) {
        uint32_t v[16];
        const uint32_t ncol = (uint32_t)c < p.N ? min(16u, p.N - (uint32_t)c) : 0u;
        // everything the chunk needs from memory is requested BEFORE the accumulator wait: the noise (fused) or the addend
        float4 u4[4];
        const bool addv = add_vec && ncol == 16;
        if (fused) {
          if (ncol) bnq_load_noise(p.bnq, bst, pix, rvalid, (uint32_t)c, ncol, p.N, u4);
        } else if (addv) {
#pragma unroll
          for (int j = 0; j < 4; ++j) u4[j] = __ldcs(reinterpret_cast<const float4*>(p.addend + obase + c) + j);
        }
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        if (!ncol) continue;   // warp-uniform
        if (fused) {
          bnq_chunk(p.bnq, bst, v, u4, scale, p.bias ? p.bias + c : nullptr, row, rvalid, (uint32_t)c, ncol, p.N, my_stat, BN,
                    (uint32_t)c, lane);
          continue;
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          f[j] = __int2float_rn((int)v[j]) * scale;
          if (p.bias && j < (int)ncol) f[j] = __fadd_rn(f[j], __ldg(p.bias + c + j));
        }
        if (rvalid) {
          float* o = p.out + obase + c;
          if (p.addend) {   // + an fp32 tensor of the output's shape (the other branch of a gradient sum)
            const float* ad = p.addend + obase + c;
            if (addv) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                f[4 * j + 0] = __fadd_rn(f[4 * j + 0], u4[j].x);
                f[4 * j + 1] = __fadd_rn(f[4 * j + 1], u4[j].y);
                f[4 * j + 2] = __fadd_rn(f[4 * j + 2], u4[j].z);
                f[4 * j + 3] = __fadd_rn(f[4 * j + 3], u4[j].w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < (int)ncol) f[j] = __fadd_rn(f[j], __ldg(ad + j));
            }
          }
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
    if (fused && ok) {
      bnq_flush_cta(p.bnq, my_stat, s_tot, 0, BN, p.N, lane, threadIdx.x - 64, 32 * EPI, 1);
      bnq_finish(p.bnq, bst, (unsigned long long)p.M * p.N, warp == 2, lane);
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

void* driver_fn(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return f;
}

template <int BN, int EPI, bool FUSED, bool DUAL = false>
int launch_halo_impl(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const HaloParams& p, unsigned grid, size_t smem,
                     cudaStream_t st) {
  static size_t attr_done[16] = {};
  const int dev = device_info().device;
  if (attr_done[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, EPI, FUSED, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_cuda_error(e, "cudaFuncSetAttribute(conv_halo_kernel)");
      return LBT_ECUDA;
    }
    attr_done[dev] = smem;
  }
  launch_pdl(conv_halo_kernel<BN, EPI, FUSED, DUAL>, grid, 32 * (2 + EPI), smem, st, ta, ta2, tb, p);
  g_halo_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch(DUAL ? "lbt_conv_i8_fprop_dual" : "lbt_conv_i8_fprop");
}

template <int BN, int EPI>
int launch_halo(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const HaloParams& p, unsigned grid, size_t smem,
                cudaStream_t st, bool dual) {
  if (dual) return launch_halo_impl<BN, EPI, false, true>(ta, ta2, tb, p, grid, smem, st);
  return p.bnq.q.bits != 0 ? launch_halo_impl<BN, EPI, true>(ta, ta2, tb, p, grid, smem, st)
                           : launch_halo_impl<BN, EPI, false>(ta, ta2, tb, p, grid, smem, st);
}

std::atomic<int> g_use_tma_halo{1};   // bit 0: on; bit 1 (tests): take ragged images whatever the patch fill ratio

}  // namespace

void conv_halo_enable(int mode) { g_use_tma_halo.store(mode, std::memory_order_relaxed); }

// Shapes the kernel takes: stride 1, 64- or 128-byte pixels, <= 128 output channels, a filter bank that fits beside a ring
// of at least three patches, and images that fill their 8 x 16 patches to at least 74 %.
bool conv_halo_applies(int N, int OH, int OW, int C, int Cout, int kh, int kw, int sh, int sw) {
  const int mode = g_use_tma_halo.load(std::memory_order_relaxed);
  if (!(mode & 1)) return false;
  if (sh != 1 || sw != 1 || (C != 64 && C != 128) || Cout > 128 || kh * kw <= 1 || kh > 5 || kw > 5) return false;
  const uint64_t tx = (uint64_t)(OW + kPatchW - 1) / kPatchW, ty = (uint64_t)(OH + kPatchH - 1) / kPatchH;
  if (!(mode & 2) && tx * kPatchW * ty * kPatchH * 100 > (uint64_t)OH * OW * 135) return false;
  if ((uint64_t)N * tx * ty >= (1ull << 31)) return false;
  const int bn = Cout <= 64 ? 64 : 128;
  const size_t b_bytes = (size_t)kh * kw * bn * C;
  const size_t stage = (((size_t)(kPatchH + kh - 1) * (kPatchW + kw - 1) * C) + 1023) & ~(size_t)1023;
  return b_bytes + 2 * stage + 2048 <= 204 * 1024;   // 227 KB minus the kernel's static shared memory (statistics partials)
}

int conv_halo_run(const void* src, int src_kind, int N, int H, int W, int C, const void* wp, int w_kind, size_t ldw, int Cout, int kh,
                  int kw, int pt, int pl, int OH, int OW, const int32_t* ibA, const int32_t* ibB, int exp_const, const float* bias,
                  float* out, size_t ldc, const lbt_qsite* q_out, int8_t* k_out, int64_t* sums, const float* addend, void* stream,
                  const OutRemap* remap, const void* src_lo) {
  const bool dual = src_lo != nullptr;   // src = high byte plane (s8), src_lo = low byte plane (u8): k = 256 * hi + lo
  if (dual && (q_out || !out)) return LBT_EINVAL;
  const DeviceInfo& di = device_info();
  static EncodeTiledFn enc_tiled = reinterpret_cast<EncodeTiledFn>(driver_fn("cuTensorMapEncodeTiled"));
  if (!enc_tiled) return LBT_ECUDA;
  const int bn = Cout <= 64 ? 64 : 128;
  uint32_t mode = 0;
  while ((16u << mode) < (uint32_t)C) ++mode;
  const uint32_t tx = (uint32_t)(OW + kPatchW - 1) / kPatchW, ty = (uint32_t)(OH + kPatchH - 1) / kPatchH;

  HaloParams p{};
  p.M = (uint32_t)((size_t)N * OH * OW);
  p.N = (uint32_t)Cout;
  p.OH = (uint32_t)OH;
  p.OW = (uint32_t)OW;
  p.OHW = (uint32_t)(OH * OW);
  p.m_tiles = (uint32_t)N * tx * ty;
  p.d_tiles_img = make_fastdiv(tx * ty);
  p.d_tiles_x = make_fastdiv(tx);
  p.pt = pt;
  p.pl = pl;
  p.kh = (uint32_t)kh;
  p.kw = (uint32_t)kw;
  p.cb = (uint32_t)C;
  p.mode = mode;
  p.halo_w = (uint32_t)(kPatchW + kw - 1);
  const uint32_t halo_h = (uint32_t)(kPatchH + kh - 1);
  p.patch_bytes = p.halo_w * halo_h * p.cb;
  p.stage_bytes = (p.patch_bytes + 1023u) & ~1023u;
  p.b_block = (uint32_t)bn * p.cb;
  p.ibA = ibA;
  p.ibB = ibB;
  p.exp_const = exp_const;
  p.bias = bias;
  p.addend = q_out ? nullptr : addend;
  p.out = out;
  p.ldc = ldc;
  p.idesc = tc::make_idesc_i8(src_kind == LBT_MANT_S8, w_kind == LBT_MANT_S8, false, false, bn, kBlockM);
  p.idesc2 = tc::make_idesc_i8(false, w_kind == LBT_MANT_S8, false, false, bn, kBlockM);
  p.bnq.q = site_from_abi(q_out);
  p.bnq.k = k_out;
  p.bnq.sums = reinterpret_cast<long long*>(sums);
  p.bnq.rows_per_image = (uint32_t)(OH * OW);
  if (remap) {
    if (q_out) return LBT_EINVAL;
    p.remap = 1;
    p.rs_n = remap->sn;
    p.rs_y = remap->sy;
    p.rs_x = remap->sx;
  }

  {
    static const int sector_on = [] {   // A/B switch: LBT_HALO_SECTOR=0 keeps the row-per-lane fp32 epilogue
      const char* e = std::getenv("LBT_HALO_SECTOR");
      return e ? std::atoi(e) : 1;
    }();
    const auto even = [](long long v) { return (v & 1) == 0; };
    const bool al = (reinterpret_cast<uintptr_t>(out) & 7u) == 0 && (!p.addend || (reinterpret_cast<uintptr_t>(p.addend) & 7u) == 0) &&
                    (!bias || (reinterpret_cast<uintptr_t>(bias) & 7u) == 0);
    p.sector_f32 = !q_out && out && Cout % 16 == 0 && al &&
                   (p.remap ? even(p.rs_n) && even(p.rs_y) && even(p.rs_x) : even((long long)ldc)) &&
                   sector_on;
  }
  if (dual) {
    if (!p.sector_f32) return LBT_EUNSUPPORTED;   // the dual mode lives in the sector epilogue only
    p.stage_bytes *= 2;                           // both planes' patches per ring slot
  }
  const size_t b_bytes = (size_t)kh * kw * p.b_block;
  // two CTAs per SM (8 epilogue warps each) when two filter banks + rings fit; else one CTA with 16 epilogue warps
  // (227 KB per SM; each CTA also holds ~1 KB of system shared memory and this kernel's static arrays: 7 - 20 KB)
  const bool two = b_bytes + 3 * (size_t)p.stage_bytes + 2048 <= 102 * 1024;
  const size_t budget = (two ? 102 * 1024 : 204 * 1024) - b_bytes - 2048;
  uint32_t nst = (uint32_t)(budget / p.stage_bytes);
  if (nst > (uint32_t)kMaxStages) nst = kMaxStages;
  if (nst < 2) return LBT_EUNSUPPORTED;
  p.nstages = nst;
  size_t smem = b_bytes + (size_t)nst * p.stage_bytes + 1024;
  if (!two && smem < 120 * 1024) smem = 120 * 1024;   // the 16-warp variant owns all 512 TMEM columns: never two per SM

  CUtensorMap ta, tb;
  const CUtensorMapSwizzle sw = mode == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C, (cuuint64_t)W * C, (cuuint64_t)H * W * C};
    cuuint32_t box[4] = {(cuuint32_t)C, p.halo_w, halo_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(src), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(halo patch)");
      return LBT_ECUDA;
    }
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)kh * kw * C, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {(cuuint64_t)ldw};
    cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(wp), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(weights)");
      return LBT_ECUDA;
    }
  }
  CUtensorMap ta2 = ta;
  if (dual) {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C, (cuuint64_t)W * C, (cuuint64_t)H * W * C};
    cuuint32_t box[4] = {(cuuint32_t)C, p.halo_w, halo_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc_tiled(&ta2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(src_lo), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(halo patch, low plane)");
      return LBT_ECUDA;
    }
  }
  const uint64_t cap = (uint64_t)di.sm_count * (two ? 2 : 1);
  const unsigned grid = (unsigned)(p.m_tiles < cap ? p.m_tiles : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (two) return bn == 64 ? launch_halo<64, 8>(ta, ta2, tb, p, grid, smem, st, dual) : launch_halo<128, 8>(ta, ta2, tb, p, grid, smem, st, dual);
  return bn == 64 ? launch_halo<64, 16>(ta, ta2, tb, p, grid, smem, st, dual) : launch_halo<128, 16>(ta, ta2, tb, p, grid, smem, st, dual);
}

int conv_halo_debug_error() {
  int v = 0, zero = 0;
  if (cudaMemcpyFromSymbol(&v, g_halo_error, sizeof(int)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_halo_error, &zero, sizeof(int));
  return v;
}

}  // namespace lbt

// Test probe (not in lbt.h): launches of the halo kernel so far — lets a test assert which kernel a shape was routed to.
extern "C" long long lbt_conv_halo_launches(void) { return lbt::g_halo_launches.load(std::memory_order_relaxed); }

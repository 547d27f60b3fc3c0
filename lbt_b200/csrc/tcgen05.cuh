// Inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: mbarrier, TMA (tiled and im2col),
// tcgen05 alloc / mma.kind::i8 / commit / ld, and shared-memory matrix descriptors.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "common.cuh"

namespace lbt {
namespace tc {

constexpr long long kWatchdogCycles = 4000000000ll;  // ~2 s at boost clock
// Optional suspend-time hint of mbarrier.try_wait (ns).  Measured on B200 (A/B builds, -DLBT_SUSPEND_HINT_NS=4000 vs 0): no
// effect on the convolution kernels or the training step, -2 % on the 8192^3 GEMM: the default (0 = no hint) stays.
#ifndef LBT_SUSPEND_HINT_NS
#define LBT_SUSPEND_HINT_NS 0
#endif
constexpr uint32_t kSuspendHintNs = LBT_SUSPEND_HINT_NS;   // 0: plain try_wait (system-default suspend time)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  if (kSuspendHintNs == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error flag, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* error_flag) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3f) == 0) {
      if (*abort_flag) return false;
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > kWatchdogCycles) {
        *abort_flag = 1;
        atomicExch(error_flag, 1);
        return false;
      }
    }
  }
  return true;
}

// ---- TMA ----
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// im2col mode over a (C, W, H, N) tensor: loads `pixelsPerColumn` pixels starting at base pixel (w, h, n),
// displaced by the filter tap (off_w, off_h), `channelsPerPixel` channels from c.
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
      "{%7, %8};" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::i8, s32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All MMAs previously issued by this thread arrive on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 16 TMEM lanes x 16 columns in the accumulator-fragment distribution (shape .16x256b, two 8-column groups): thread t holds,
// for column group n (0, 1), v[4n + 0..1] = row t/4, columns 8n + 2(t%4) + {0, 1} and v[4n + 2..3] = row t/4 + 8, same columns.
// Four neighbouring lanes hold 32 contiguous bytes of one row: fp32 rows are then stored as full 32-byte sectors (the
// 32x32b shape gives every lane its own row, i.e. 16-byte pieces 32 rows apart).  taddr's lane = first of the 16 lanes.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- shared-memory matrix descriptors (PTX ISA "tcgen05 matrix descriptor") ----
// bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
// [46,48) version = 1 | [61,64) layout: 0 none (interleaved 8x16B core matrices), 6 = 32B, 4 = 64B, 2 = 128B swizzle.
//
// SwizzleMode m = log2(row bytes / 16): 0 -> 16-byte rows, no swizzle; 1 -> 32B; 2 -> 64B; 3 -> 128B.
__device__ __forceinline__ uint64_t desc_layout_bits(int m) {
  return m == 0 ? 0ull : (m == 1 ? 6ull : (m == 2 ? 4ull : 2ull));
}
// K-major operand block: rows (M or N index) are `16 << m` bytes of K apart; 8-row atoms are contiguous.
//   m == 0: one 16-byte K chunk per block row; the second K chunk of a K=32 instruction lives `lbo_bytes` away.
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr, int m, uint32_t lbo_bytes) {
  const uint32_t sbo = 128u << m;  // 8 rows * row bytes
  const uint32_t lbo = m == 0 ? lbo_bytes : 16u;
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         (desc_layout_bits(m) << 61);
}
// MN-major operand: shared memory holds [K rows][16 << m bytes of M (or N)] blocks — exactly what a TMA load of
// "pixels x channels" produces — and the reduction runs down the rows.  An 8-row K atom is (128 << m) bytes;
// `group_stride` is the byte distance between consecutive (16 << m)-wide groups along M/N (the next block).
//   m == 0: SBO = MN-group stride, LBO = K-atom stride;  swizzled: LBO = MN-group stride, SBO = K-atom stride.
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t smem_addr, int m, uint32_t group_stride) {
  const uint32_t katom = 128u << m;
  const uint32_t lbo = m == 0 ? katom : group_stride;
  const uint32_t sbo = m == 0 ? group_stride : katom;
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | (desc_layout_bits(m) << 61);
}
// 128-byte-swizzled K-major tile (rows of 128 B): the plain GEMM operand.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) { return make_desc_kmajor(smem_addr, 3, 16); }

// instruction descriptor for kind::i8: c_format s32 (2) @4, a_format @7, b_format @10 (0 = u8, 1 = s8),
// a_major @15, b_major @16 (0 = K-major, 1 = MN-major), N>>3 @17, M>>4 @24
__host__ __device__ inline uint32_t make_idesc_i8(bool a_signed, bool b_signed, bool a_mn_major, bool b_mn_major, int n, int m) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace tc
}  // namespace lbt

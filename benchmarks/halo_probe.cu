// Halo-patch probe (B200): can tcgen05.mma read a SWIZZLED K-major operand whose 8-row groups are 8 neighbouring pixels of a
// TMA-loaded [H][W][cb] patch, starting at an arbitrary pixel (dy, dx) — i.e. at a start address that is NOT aligned to the
// swizzle's repeating pattern, with a stride-byte-offset (halo row pitch) that is not a multiple of it either?
//   D[m][n] = sum_k A[m][k] * B[n][k],  B = one-hot (n == k)  =>  D[m][n] = A[m][n]: the result shows which bytes the tensor
//   core fetched for row m.  Variants: descriptor "base offset" field 0 / (start >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I lbt_b200/csrc -o benchmarks/bin/halo_probe benchmarks/halo_probe.cu -lcuda
// Not part of the library; informs the halo loader of the wide-channel convolution kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
namespace lbt {
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_pdl{0};
std::atomic<int> g_carveout{-1};
}  // namespace lbt
#include "tcgen05.cuh"

using namespace lbt::tc;

constexpr int BW = 10, BH = 18;   // halo patch of an 8 x 16 output patch under a 3 x 3 filter
constexpr int kTests = 3 * 3 * 4 * 2;

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c, int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, int mode, int cb, int w0, int h0, int* out,
                                                 uint8_t* raw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = sA + 32 * 1024;
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort, s_err;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    s_abort = 0;
    s_err = 0;
    fence_barrier_init();
  }
  // B[n][k] one-hot, N = 32, K = 32, un-swizzled K-major: offset = (k / 16) * 512 + n * 16 + k % 16
  for (int i = threadIdx.x; i < 1024; i += 128) sB[i] = 0;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int n = threadIdx.x, k = n;
    sB[(k / 16) * 512 + n * 16 + (k % 16)] = 1;
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, (uint32_t)(BW * BH * cb));
    tma_load_4d(&tm, &full_bar, sA, 0, w0, h0, 0);
  }
  mbar_wait(&full_bar, 0, &s_abort, &s_err);
  __syncthreads();
  for (int i = threadIdx.x; i < BW * BH * cb; i += 128) raw[i] = sA[i];   // the patch as the TMA laid it out
  const uint32_t idesc = make_idesc_i8(false, true, false, false, 32, 128);
  uint32_t phase = 0;
  for (int t = 0; t < kTests; ++t) {
    const int variant = t & 1, kk = (t >> 1) & 3, dx = (t >> 3) % 3, dy = (t >> 3) / 3;
    if (kk * 32 >= cb) {
      continue;
    }
    if (threadIdx.x == 0) {
      const uint32_t start = smem_u32(sA) + (uint32_t)((dy * BW + dx) * cb + kk * 32);
      const uint32_t sbo = (uint32_t)(BW * cb);
      uint64_t da = (uint64_t)((start & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
                    (desc_layout_bits(mode) << 61);
      if (variant) da |= (uint64_t)((start >> 7) & 7u) << 49;
      umma_i8(tmem_base, da, make_desc_kmajor(smem_u32(sB), 0, 512), idesc, 0u);
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, phase, &s_abort, &s_err);
    phase ^= 1;
    fence_after();
    uint32_t v[16];
    for (int c = 0; c < 32; c += 16) {
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[((size_t)t * 128 + warp * 32 + lane) * 32 + c + j] = (int)v[j];
    }
    fence_before();
    __syncthreads();
    fence_after();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 32);
  if (threadIdx.x == 0 && s_err) out[0] = -12345;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaFree(0);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
  const int H = 24, W = 20;
  int fails_total = 0;
  for (int mode = 2; mode <= 3; ++mode) {
    const int cb = 16 << mode;
    const size_t n = (size_t)H * W * cb;
    uint8_t* h = (uint8_t*)malloc(n);
    uint32_t s = 12345u + mode;
    for (size_t i = 0; i < n; ++i) {
      s = s * 1664525u + 1013904223u;
      h[i] = (uint8_t)(s >> 24);
    }
    uint8_t *d, *raw;
    int* out;
    cudaMalloc(&d, n);
    cudaMalloc(&raw, BW * BH * cb);
    cudaMalloc(&out, sizeof(int) * kTests * 128 * 32);
    cudaMemcpy(d, h, n, cudaMemcpyHostToDevice);
    cudaMemset(out, 0xff, sizeof(int) * kTests * 128 * 32);
    CUtensorMap tm;
    cuuint64_t gdim[4] = {(cuuint64_t)cb, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t gstr[3] = {(cuuint64_t)cb, (cuuint64_t)W * cb, (cuuint64_t)H * W * cb};
    cuuint32_t box[4] = {(cuuint32_t)cb, BW, BH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mode == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d\n", (int)r);
      return 1;
    }
    for (int origin = 0; origin < 2; ++origin) {
      const int w0 = origin ? 3 : -1, h0 = origin ? 2 : -1;
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      probe<<<1, 128, 64 * 1024>>>(tm, mode, cb, w0, h0, out, raw);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
      int* ho = (int*)malloc(sizeof(int) * kTests * 128 * 32);
      cudaMemcpy(ho, out, sizeof(int) * kTests * 128 * 32, cudaMemcpyDeviceToHost);
      for (int t = 0; t < kTests; ++t) {
        const int variant = t & 1, kk = (t >> 1) & 3, dx = (t >> 3) % 3, dy = (t >> 3) / 3;
        if (kk * 32 >= cb) continue;
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int k = 0; k < 32; ++k) {
            const int py = m / 8 + dy + h0, px = m % 8 + dx + w0;
            const int want = (py >= 0 && py < H && px >= 0 && px < W) ? h[((size_t)py * W + px) * cb + kk * 32 + k] : 0;
            if (ho[((size_t)t * 128 + m) * 32 + k] != want) ++bad;
          }
        printf("mode %d (cb %3d) origin (%2d,%2d) dy %d dx %d kk %d base_offset %s: %s (%d wrong of 4096)\n", mode, cb, w0, h0, dy, dx,
               kk, variant ? "start>>7&7" : "0", bad ? "MISMATCH" : "ok", bad);
        if (bad && !variant) ++fails_total;
      }
      free(ho);
    }
    free(h);
    cudaFree(d);
    cudaFree(raw);
    cudaFree(out);
  }
  printf("variant-0 failures: %d\n", fails_total);
  return 0;
}

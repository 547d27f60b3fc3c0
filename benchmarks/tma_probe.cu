// TMA throughput probe (B200): how fast does one SM pull [128 pixels x cb bytes] boxes through
//   (a) im2col-mode TMA over an NHWC tensor, (b) tiled 2-D TMA over the same bytes viewed as [pixels, C],
//   (c) one tall tiled box (halo tile: 256 pixels x cb).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tma_probe benchmarks/tma_probe.cu
// Not part of the library; informs the conv kernel's loader design.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_im2col(const CUtensorMap* map, uint64_t* bar, void* dst, int c, int w, int h, int n,
                                           uint16_t ow, uint16_t oh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
      : "memory");
}

constexpr int kStages = 8;

// mode 0: im2col, 9 taps per 128-pixel tile; mode 1: tiled [128 x cb] x 9 (same bytes, contiguous rows);
// mode 2: tiled halo [rows x cb] once per tile
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int mode, int cb, int rows, int tiles,
                                                int W, int H, int OHW, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kStages];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  const uint32_t box_bytes = (uint32_t)rows * cb;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    uint32_t issued = 0, waited = 0;
    const int loads_per_tile = mode == 2 ? 1 : 9;
    const long total = (long)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * loads_per_tile;
    long li = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int m0 = tile * 128;
      const int img = m0 / OHW, rem = m0 % OHW, oh = rem / W, ow = rem % W;
      for (int t = 0; t < loads_per_tile; ++t, ++li) {
        if (issued - waited == kStages) {  // ring full: wait for the oldest
          const uint32_t s = waited % kStages, ph = (waited / kStages) & 1;
          while (!mbar_try_wait(&full[s], ph)) {}
          ++waited;
        }
        const uint32_t s = issued % kStages;
        mbar_expect_tx(&full[s], box_bytes);
        uint8_t* dst = base + (size_t)s * ((box_bytes + 1023) & ~1023u);
        if (mode == 0) tma_im2col(&tm, &full[s], dst, 0, ow - 1, oh - 1, img, (uint16_t)(t % 3), (uint16_t)(t / 3));
        else if (mode == 1) tma_2d(&tm, &full[s], dst, 0, m0 + (t / 3) * W + (t % 3));
        else tma_2d(&tm, &full[s], dst, 0, m0);
        ++issued;
      }
    }
    while (waited < issued) {
      const uint32_t s = waited % kStages, ph = (waited / kStages) & 1;
      while (!mbar_try_wait(&full[s], ph)) {}
      ++waited;
    }
    cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    (void)total;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void* drv(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return f;
}
static CUtensorMapSwizzle swz(int cb) {
  return cb == 16 ? CU_TENSOR_MAP_SWIZZLE_NONE : cb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : cb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                               : CU_TENSOR_MAP_SWIZZLE_128B;
}

int main() {
  EncodeTiledFn enc_tiled = (EncodeTiledFn)drv("cuTensorMapEncodeTiled");
  EncodeIm2colFn enc_im2col = (EncodeIm2colFn)drv("cuTensorMapEncodeIm2col");
  if (!enc_tiled || !enc_im2col) return 1;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  struct Case { int N, H, W, C; } cases[] = {{256, 32, 32, 16}, {256, 16, 16, 32}, {256, 56, 56, 64}, {256, 28, 28, 128}};
  unsigned long long* d_cycles;
  cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (auto cs : cases) {
    const size_t npix = (size_t)cs.N * cs.H * cs.W;
    uint8_t* x;
    cudaMalloc(&x, npix * cs.C + 65536);
    cudaMemset(x, 1, npix * cs.C + 65536);
    const int cb = cs.C > 128 ? 128 : cs.C;
    const int tiles = (int)(npix / 128);
    for (int mode = 0; mode < 3; ++mode) {
      CUtensorMap tm;
      int rows = 128;
      if (mode == 0) {
        cuuint64_t gdim[4] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)cs.N};
        cuuint64_t gstr[3] = {(cuuint64_t)cs.C, (cuuint64_t)cs.W * cs.C, (cuuint64_t)cs.H * cs.W * cs.C};
        int lower[2] = {-1, -1}, upper[2] = {-1, -1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc_im2col(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, x, gdim, gstr, lower, upper, cb, 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swz(cb), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
          printf("im2col encode failed\n");
          return 1;
        }
      } else {
        rows = mode == 1 ? 128 : 128 + 2 * (cs.W + 2) + 2;
        if (rows > 256) rows = 256;
        cuuint64_t gdim[2] = {(cuuint64_t)cs.C, (cuuint64_t)npix};
        cuuint64_t gstr[1] = {(cuuint64_t)cs.C};
        cuuint32_t box[2] = {(cuuint32_t)cb, (cuuint32_t)rows};
        cuuint32_t estr[2] = {1, 1};
        if (enc_tiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(cb),
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
          printf("tiled encode failed\n");
          return 1;
        }
      }
      const size_t smem = (size_t)kStages * (((size_t)rows * cb + 1023) & ~(size_t)1023) + 1024;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float best = 1e9f;
      for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        probe<<<sms, 64, smem>>>(tm, mode, cb, rows, tiles, cs.W, cs.H, cs.H * cs.W, d_cycles);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("kernel failed: %s\n", cudaGetErrorString(e));
          return 1;
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double loads = (double)tiles * (mode == 2 ? 1 : 9);
      const double bytes = loads * rows * cb;
      const double cyc_per_sm = best * 1e-3 * clk_khz * 1e3;
      printf("N%d H%d W%d C%d cb%d mode %d (%s): %8.1f us, %.1f MB into smem, %.2f TB/s, %.1f B/clk/SM, %.2f clk per pixel-row, %.0f clk per load\n",
             cs.N, cs.H, cs.W, cs.C, cb, mode, mode == 0 ? "im2col 9 taps" : (mode == 1 ? "tiled 128-row x9" : "tiled halo x1"), best * 1e3,
             bytes / 1e6, bytes / (best * 1e-3) / 1e12, bytes / sms / cyc_per_sm, cyc_per_sm / (loads * rows / sms),
             cyc_per_sm / (loads / sms));
    }
    cudaFree(x);
  }
  return 0;
}

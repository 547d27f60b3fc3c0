// Multi-tensor kernels that turn per-layer launch storms into ONE launch per step:
//   lbt_finalize_multi : every integer gradient sum -> fp32 gradient (+ 2*wd*W)   [was 1 launch per parameter]
//   lbt_param_prep     : every parameter quantiser of the step (conv/dense weights, biases, BN gamma/beta):
//                        stochastic DFXP quantisation + overflow counters + the packed tensor-core operand
//                        layouts, straight from the fp32 masters                 [was ~6 launches per layer]
// The reference does these as dozens of TF ops per variable (dynamic_fixed_point.py:194-200, 289-295, 386-392,
// 679-682 for the quantisers; :207, :302, :457, :689-690 for the gradients).
#include "conv_classes.h"
#include "common.cuh"

namespace lbt {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) finalize_multi_kernel(const lbt_finalize_job* __restrict__ jobs, int njobs,
                                                                  unsigned long long total) {
  pdl_trigger();
  pdl_wait();
  for (unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * kThreads) {
    int lo = 0, hi = njobs - 1;  // last job with start <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].start <= i) lo = mid; else hi = mid - 1;
    }
    const lbt_finalize_job j = jobs[lo];
    const unsigned long long k = i - j.start;
    int e = j.exp_const;
    if (j.ibA) e += *j.ibA;
    if (j.ibB) e += *j.ibB;
    float v = __ll2float_rn((long long)j.acc64[k]) * exp2i(e);
    if (j.add) v = __fadd_rn(v, __fmul_rn(j.add_scale, j.add[k]));
    j.out[k] = v;
  }
}

// ---- noise for the fused tensor-core epilogues ---------------------------------------------------------------
// The epilogue threads of the convolution kernels would otherwise run Philox4x32-10 themselves (100 instructions per
// four outputs on four warps); the noise of a site is shared over the batch, so ONE launch per step materialises
// every site's [n_inner] vector (<= 1/batch of the activation, L2-resident) from the very same stream.
__global__ void __launch_bounds__(kThreads) noise_fill_multi_kernel(const lbt_noise_job* __restrict__ jobs, int njobs,
                                                                    unsigned long long total_groups, uint64_t seed,
                                                                    const uint64_t* dev_step) {
  pdl_trigger();
  pdl_wait();
  const uint64_t step_off = dev_step ? ((*dev_step) << 32) : 0ull;
  for (unsigned long long i = (unsigned long long)blockIdx.x * kThreads + threadIdx.x; i < total_groups;
       i += (unsigned long long)gridDim.x * kThreads) {
    int lo = 0, hi = njobs - 1;  // last job with start <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].start <= i) lo = mid; else hi = mid - 1;
    }
    const lbt_noise_job j = jobs[lo];
    const unsigned long long g = i - j.start;
    const float4 v = philox_noise4(g, seed, j.offset + step_off);
    const unsigned long long e = 4ull * g;
    if (e + 3 < j.n) {
      *reinterpret_cast<float4*>(j.u + e) = v;
    } else {
      if (e + 0 < j.n) j.u[e + 0] = v.x;
      if (e + 1 < j.n) j.u[e + 1] = v.y;
      if (e + 2 < j.n) j.u[e + 2] = v.z;
    }
  }
}

// ---- parameter preparation ---------------------------------------------------------------------
struct QC {
  float m, inv_m, L, hi, half;
};
__device__ __forceinline__ QC make_qc(int bits, int ib) {
  QC c;
  int f = bits - ib - 1;
  f = max(-126, min(126, f));
  c.m = exp2i(f);
  c.inv_m = exp2i(-f);
  c.L = exp2i(bits - 1);
  c.hi = c.L - 1.0f;
  c.half = c.L * 0.5f;
  return c;
}

__global__ void __launch_bounds__(kThreads) param_prep_kernel(const lbt_prep_job* __restrict__ jobs,
                                                              const uint32_t* __restrict__ block_job,
                                                              const uint32_t* __restrict__ block_chunk, uint64_t seed,
                                                              const uint64_t* dev_step, uint32_t chunk_elems) {
  pdl_trigger();
  pdl_wait();
  const lbt_prep_job j = jobs[block_job[blockIdx.x]];
  const QC c = make_qc(j.bits, *j.ib);
  uint64_t off = j.offset;
  if (dev_step) off += (*dev_step) << 32;
  const uint64_t n = (uint64_t)j.n_outer * j.n_inner;
  const uint64_t i0 = (uint64_t)block_chunk[blockIdx.x] * chunk_elems;
  const uint64_t i1 = min(i0 + (uint64_t)chunk_elems, n);
  uint32_t n1 = 0, n2 = 0;
  // one element: quantise (stochastic_identity, dfxp:34-37), statistics, fp32 vector and / or packed operand layouts
  auto emit = [&](const uint64_t i, const float x, const float u) {
    const float y = __fmul_rn(x, c.m);
    n1 += (uint32_t)(y >= c.L) + (uint32_t)(y < -c.L);
    n2 += (uint32_t)(y >= c.half) + (uint32_t)(y < -c.half);
    const float k = floorf(fminf(fmaxf(__fadd_rn(y, u), -c.L), c.hi));   // stochastic_identity, dfxp:34-37
    if (j.out_f32) j.out_f32[i] = k * c.inv_m;
    const int8_t kb = (int8_t)__float2int_rn(k);
    if (j.layout == LBT_PREP_CONV) {
      // i = ((r*kw + s)*Cin + ci)*Cout + co   (HWIO)
      const uint32_t co = (uint32_t)(i % j.Cout);
      const uint64_t t = i / j.Cout;
      const uint32_t ci = (uint32_t)(t % j.Cin), tap = (uint32_t)(t / j.Cin);
      if (j.out_a) {
        if (j.c3pad) {  // first layer: 16 pseudo-channels {W, W, W, 0 x 7} per tap
          int8_t* d = j.out_a + (size_t)co * j.ld_a + (size_t)tap * 16 + ci;
          d[0] = kb; d[3] = kb; d[6] = kb;
        } else {
          j.out_a[(size_t)co * j.ld_a + (size_t)tap * j.Cin + ci] = kb;          // fprop B: [Cout, (r,s,ci)]
        }
      }
      if (j.out_b) {
        const uint32_t r = tap / j.kw, s = tap % j.kw;
        // rot180 = 1: stride-1 input gradient = correlation with the rotated filter; 2: taps grouped by parity class for
        // the strided input gradient (conv_classes.h; strides in the job's sh / sw); 0: filter order
        const uint32_t tap2 = j.rot180 == 1 ? ((j.kh - 1 - r) * j.kw + (j.kw - 1 - s))
                                            : (j.rot180 == 2 ? class_tap_index((int)r, (int)s, (int)j.kh, (int)j.kw, j.sh, j.sw) : tap);
        j.out_b[(size_t)ci * j.ld_b + (size_t)tap2 * j.Cout + co] = kb;          // dgrad B: [Cin, (r',s',co)]
      }
    } else if (j.layout == LBT_PREP_DENSE) {
      // i = in*Cout + out   ([in, out])
      const uint32_t o = (uint32_t)(i % j.Cout), in = (uint32_t)(i / j.Cout);
      if (j.out_a) j.out_a[(size_t)o * j.ld_a + in] = kb;                        // fprop B: [out, in]
      if (j.out_b) j.out_b[(size_t)in * j.ld_b + o] = kb;                        // dgrad B: [in, out] (padded pitch)
    }
  };
  if ((j.n_inner & 3) == 0 && (reinterpret_cast<uintptr_t>(j.x) & 15) == 0) {
    // four consecutive elements per thread share ONE Philox block (the noise index is the inner index / 4; chunks start on
    // multiples of 4): the per-element version ran the ten Philox rounds four times over (270 instructions per parameter, ncu)
    for (uint64_t i = i0 + 4ull * threadIdx.x; i < i1; i += 4ull * kThreads) {
      const float4 u4 = philox_noise4((i % j.n_inner) >> 2, seed, off);
      const float4 xv = __ldg(reinterpret_cast<const float4*>(j.x + i));
      emit(i, xv.x, u4.x);
      emit(i + 1, xv.y, u4.y);
      emit(i + 2, xv.z, u4.z);
      emit(i + 3, xv.w, u4.w);
    }
  } else {
    for (uint64_t i = i0 + threadIdx.x; i < i1; i += kThreads) {
      const uint64_t col = i % j.n_inner;
      const float4 u4 = philox_noise4(col >> 2, seed, off);
      const int l = (int)(col & 3);
      emit(i, j.x[i], l == 0 ? u4.x : (l == 1 ? u4.y : (l == 2 ? u4.z : u4.w)));
    }
  }
  // block-reduce the counters, publish to the quantiser's statistics block
  n1 = warp_sum(n1);
  n2 = warp_sum(n2);
  __shared__ uint32_t s1[kThreads / 32], s2[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    s1[w] = n1;
    s2[w] = n2;
  }
  __syncthreads();
  if (threadIdx.x == 0 && j.counters) {
    uint32_t b1 = 0, b2 = 0;
    for (int q = 0; q < kThreads / 32; ++q) {
      b1 += s1[q];
      b2 += s2[q];
    }
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(j.counters);
    if (b1) atomicAdd(cnt + LBT_CNT_OVER, (unsigned long long)b1);
    if (b2) atomicAdd(cnt + LBT_CNT_OVER_HALF, (unsigned long long)b2);
    atomicAdd(cnt + LBT_CNT_NUMEL, (unsigned long long)(i1 - i0));
  }
}

}  // namespace
}  // namespace lbt

using namespace lbt;

extern "C" int lbt_finalize_multi(const lbt_finalize_job* jobs_dev, size_t njobs, uint64_t total, void* stream) {
  if (!jobs_dev) return LBT_EINVAL;
  if (njobs == 0 || total == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  const uint64_t blocks = (total + kThreads - 1) / kThreads;
  const uint64_t cap = (uint64_t)di.sm_count * 8;
  launch_pdl(finalize_multi_kernel, (unsigned)(blocks < cap ? blocks : cap), kThreads, 0, reinterpret_cast<cudaStream_t>(stream), 
      jobs_dev, (int)njobs, (unsigned long long)total);
  return check_launch("lbt_finalize_multi");
}

extern "C" int lbt_param_prep(const lbt_prep_job* jobs_dev, const uint32_t* block_job_dev, const uint32_t* block_chunk_dev,
                              size_t nblocks, uint32_t chunk_elems, uint64_t seed, const uint64_t* dev_step, void* stream) {
  if (!jobs_dev || !block_job_dev || !block_chunk_dev || chunk_elems == 0) return LBT_EINVAL;
  if (nblocks == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  launch_pdl(param_prep_kernel, (unsigned)nblocks, kThreads, 0, reinterpret_cast<cudaStream_t>(stream), 
      jobs_dev, block_job_dev, block_chunk_dev, seed, dev_step, chunk_elems);
  return check_launch("lbt_param_prep");
}

extern "C" int lbt_noise_fill_multi(const lbt_noise_job* jobs_dev, size_t njobs, uint64_t total_groups, uint64_t seed,
                                    const uint64_t* dev_step, void* stream) {
  if (!jobs_dev) return LBT_EINVAL;
  if (njobs == 0 || total_groups == 0) return LBT_OK;
  if (njobs > 0x7fffffffull) return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  const unsigned long long blocks = (total_groups + kThreads - 1) / kThreads;
  const unsigned grid = (unsigned)(blocks < 148ull * 16 ? blocks : 148ull * 16);
  launch_pdl(noise_fill_multi_kernel, grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream), jobs_dev, (int)njobs, total_groups,
                                                                                        seed, dev_step);
  return check_launch("lbt_noise_fill_multi");
}

// Momentum SGD over the flattened parameter buffer: one launch per step instead of TF's per-variable
// ApplyMomentum kernels (tf.train.MomentumOptimizer.apply_gradients, /root/reference/trainer.py:81-82):
//     accum <- momentum * accum + grad_scale * grad ;  var <- var - lr * accum        (non-Nesterov)
// HBM-bound: 12 B read + 8 B written per parameter, 128-bit accesses.
#include "common.cuh"

namespace lbt {
namespace {

__global__ void __launch_bounds__(256) sgd_momentum_kernel(float* __restrict__ w, float* __restrict__ a,
                                                           const float* __restrict__ g, size_t n, float lr,
                                                           const float* dev_lr, float momentum, float grad_scale) {
  pdl_trigger();   // programmatic dependent launch: the next kernel may be scheduled now ...
  pdl_wait();      // ... and this one touches global memory only after its predecessor has completed
  if (dev_lr) lr = *dev_lr;
  const size_t nv = n / 4;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < nv; i += step) {
    float4 wv = reinterpret_cast<float4*>(w)[i], av = reinterpret_cast<float4*>(a)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    av.x = __fadd_rn(__fmul_rn(momentum, av.x), grad_scale * gv.x);
    av.y = __fadd_rn(__fmul_rn(momentum, av.y), grad_scale * gv.y);
    av.z = __fadd_rn(__fmul_rn(momentum, av.z), grad_scale * gv.z);
    av.w = __fadd_rn(__fmul_rn(momentum, av.w), grad_scale * gv.w);
    wv.x = __fsub_rn(wv.x, __fmul_rn(lr, av.x));
    wv.y = __fsub_rn(wv.y, __fmul_rn(lr, av.y));
    wv.z = __fsub_rn(wv.z, __fmul_rn(lr, av.z));
    wv.w = __fsub_rn(wv.w, __fmul_rn(lr, av.w));
    reinterpret_cast<float4*>(a)[i] = av;
    reinterpret_cast<float4*>(w)[i] = wv;
  }
  for (size_t i = nv * 4 + tid; i < n; i += step) {
    const float av = __fadd_rn(__fmul_rn(momentum, a[i]), grad_scale * g[i]);
    a[i] = av;
    w[i] = __fsub_rn(w[i], __fmul_rn(lr, av));
  }
}

}  // namespace
}  // namespace lbt

using namespace lbt;

extern "C" int lbt_sgd_momentum(float* w, float* accum, const float* grad, size_t n, float lr, const float* dev_lr,
                                float momentum, float grad_scale, void* stream) {
  if (!w || !accum || !grad) return LBT_EINVAL;
  if (n == 0) return LBT_OK;
  if ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(accum) | reinterpret_cast<uintptr_t>(grad)) & 15)
    return LBT_EUNSUPPORTED;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  const size_t blocks = (n / 4 + 255) / 256 + 1;
  const size_t cap = (size_t)di.sm_count * 8;
  launch_pdl(sgd_momentum_kernel, (unsigned)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      w, accum, grad, n, lr, dev_lr, momentum, grad_scale);
  return check_launch("lbt_sgd_momentum");
}

// The reference's input pipeline for one training batch, on the device:
//   main.py:71-75     X = (X - mean(X_train, axis=0)) / 128       numpy float64, cast to fp32 at the feed_dict
//   trainer.py:24-28  random_flip_left_right, pad_to_bounding_box(4, 4, H+8, W+8), random_crop(H, W)
//   trainer.py:92-96  shuffle + batch (the gather through `index`)
// The data set stays in HBM as the raw uint8 images (1 B/element read, 4 B/element written); a thread produces four
// consecutive floats of one output row.
#include "common.cuh"

namespace lbt {
namespace {

__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ src, const double* __restrict__ mean,
                                                      const long long* __restrict__ index, int B, int H, int W, int C, int pad,
                                                      int do_flip, const int32_t* __restrict__ params, uint64_t seed,
                                                      uint64_t offset, const long long* __restrict__ labels,
                                                      long long* __restrict__ labels_out, float* __restrict__ out) {
  const size_t row_elems = (size_t)W * C, img_elems = (size_t)H * row_elems;
  const size_t total = (size_t)B * img_elems;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / img_elems);
    const size_t r = e - (size_t)b * img_elems;
    const int y = (int)(r / row_elems);
    const int xc = (int)(r - (size_t)y * row_elems);
    const int x = xc / C, c = xc - x * C;
    int flip, oy, ox;
    if (params) {
      flip = params[3 * b];
      oy = params[3 * b + 1];
      ox = params[3 * b + 2];
    } else {   // one Philox draw per sample: counter = sample index in the batch, offset = (stream id, step)
      const uint4 q = philox4x32_10(make_uint4((uint32_t)b, 0u, (uint32_t)offset, (uint32_t)(offset >> 32)),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      flip = (int)(q.x & 1u);
      oy = (int)(q.y % (uint32_t)(2 * pad + 1));
      ox = (int)(q.z % (uint32_t)(2 * pad + 1));
    }
    if (!do_flip) flip = 0;
    // output pixel (y, x) = padded image pixel (y + oy, x + ox) = flipped image pixel (y + oy - pad, x + ox - pad)
    const int sy = y + oy - pad;
    int sx = x + ox - pad;
    float v = 0.f;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
      if (flip) sx = W - 1 - sx;
      const size_t off = (size_t)sy * row_elems + (size_t)sx * C + c;
      const long long i = index ? index[b] : (long long)b;
      const double d = (double)src[(size_t)i * img_elems + off] - (mean ? mean[off] : 0.0);   // float64, as numpy
      v = (float)(d * 0.0078125);                                                           // / 128 exact; one fp32 rounding
    }
    out[e] = v;
    if (labels && labels_out && r == 0) labels_out[b] = labels[index ? index[b] : (long long)b];
  }
}

}  // namespace
}  // namespace lbt

using namespace lbt;

extern "C" int lbt_augment_batch(const uint8_t* src, const double* mean, const int64_t* index, int B, int H, int W, int C, int pad,
                                 int do_flip, const int32_t* params, uint64_t seed, uint64_t offset, const int64_t* labels,
                                 int64_t* labels_out, float* out, void* stream) {
  if (!src || !out) return LBT_EINVAL;
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || pad < 0) return LBT_EINVAL;
  if (B == 0) return LBT_OK;
  LBT_REQUIRE_ARCH();
  const DeviceInfo& di = device_info();
  const size_t total = (size_t)B * H * W * C;
  const size_t blocks = (total + 255) / 256, cap = (size_t)di.sm_count * 16;
  augment_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, mean, reinterpret_cast<const long long*>(index), B, H, W, C, pad, do_flip, params, seed, offset,
      reinterpret_cast<const long long*>(labels), reinterpret_cast<long long*>(labels_out), out);
  return check_launch("lbt_augment_batch");
}

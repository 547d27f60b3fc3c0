// Tap order of a STRIDED convolution's input-gradient operand (lbt_conv_i8_dgrad_strided, lbt_param_prep).
//
// dX[sh*i + a, sw*j + b] = sum over the taps (r, s) with r = (a + pad_top) mod sh, s = (b + pad_left) mod sw (mod the stride)
// of g[i + qa - r / sh, j + qb - s / sw] * W[r, s]: every parity class (a, b) of input pixels is a STRIDE-1 convolution
// of g with the sub-filter of the taps in one residue group (r mod sh, s mod sw).  The operand keeps the taps grouped by
// residue, groups in row-major (r mod sh, s mod sw) order, and inside a group in REVERSED (r / sh, s / sw) order — the
// order a stride-1 correlation visits them — so each class's sub-filter is one contiguous column range of
// W2[Cin, taps * Cout].  No transposed im2col matrix, no multiply by the zeros of a dilated gradient.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define LBT_HD __host__ __device__
#else
#define LBT_HD
#endif

namespace lbt {

// taps of one axis whose index is congruent to r0 modulo the stride
LBT_HD inline int class_count(int r0, int k, int stride) { return r0 < k ? (k - 1 - r0) / stride + 1 : 0; }

// first tap (in units of taps) of residue group (r0, s0)
LBT_HD inline uint32_t class_group_offset(int r0, int s0, int kh, int kw, int sh, int sw) {
  // groups (a, b) before (r0, s0) in row-major order: whole rows a < r0 hold every column (kw taps per filter row)
  int rows_before = 0, cols_before = 0;
  for (int a = 0; a < r0; ++a) rows_before += class_count(a, kh, sh);
  for (int b = 0; b < s0; ++b) cols_before += class_count(b, kw, sw);
  return (uint32_t)(rows_before * kw + class_count(r0, kh, sh) * cols_before);
}

// position of filter tap (r, s) in the class-ordered operand
LBT_HD inline uint32_t class_tap_index(int r, int s, int kh, int kw, int sh, int sw) {
  const int r0 = r % sh, s0 = s % sw;
  const int nr = class_count(r0, kh, sh), nc = class_count(s0, kw, sw);
  const int rp = nr - 1 - r / sh, sp = nc - 1 - s / sw;
  return class_group_offset(r0, s0, kh, kw, sh, sw) + (uint32_t)(rp * nc + sp);
}

}  // namespace lbt

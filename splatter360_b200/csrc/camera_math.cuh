// camera_math.cuh -- inverse of one row-major 4x4 matrix, host + device (the kernel in camera.cu runs one per thread; the
// host harness of tests/test_host_math.py runs the same code on the CPU).  Gauss-Jordan with partial pivoting in double
// precision, rounded once to float: within half an ulp of the exact inverse for camera matrices.  Fully unrolled, the
// augmented matrix lives in registers (conditional row swaps instead of indexed ones).  Singular input gives inf / nan.
#pragma once
#include "common.cuh"

namespace s360 {

S360_HD void invert4x4(const float* in, float* out) {
  double a[4][8];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      a[r][c] = (double)in[r * 4 + c];
      a[r][4 + c] = r == c ? 1.0 : 0.0;
    }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    // partial pivoting: bring the largest |a[r][k]|, r >= k, to row k
#pragma unroll
    for (int r = k + 1; r < 4; r++) {
      if (fabs(a[r][k]) > fabs(a[k][k])) {
#pragma unroll
        for (int c = 0; c < 8; c++) { const double t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t; }
      }
    }
    const double inv = 1.0 / a[k][k];
#pragma unroll
    for (int c = 0; c < 8; c++) a[k][c] *= inv;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (r == k) continue;
      const double f = a[r][k];
#pragma unroll
      for (int c = 0; c < 8; c++) a[r][c] -= f * a[k][c];
    }
  }
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) out[r * 4 + c] = (float)a[r][4 + c];
}

}  // namespace s360

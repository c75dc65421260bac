// render_cull.cuh -- per-instance staging and the exact "can the alpha >= 1/255 ellipse reach this 8x8 pixel block" test of
// the render kernels.  Host + device: render.cu inlines it; tests/host_harness builds it for the CPU, where
// tests/test_host_math.py checks by brute force over the pixels that the test never culls a contributing instance.
#pragma once
#include "common.cuh"

namespace s360 {

constexpr int RT = 128;                 // threads per tile: 4 warps, each owns an 8x8 pixel block
constexpr int WARP_W = 8, WARP_H = 8;   // pixel block of one warp; a lane owns (x, y) and (x, y + 4)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float HALF_W = 0.5f * (WARP_W - 1), HALF_H = 0.5f * (WARP_H - 1);


#ifndef S360_EXACT_CULL
#define S360_EXACT_CULL 1
#endif
constexpr float LOG2_ALPHA_MIN = -7.99435343685886f;   // log2(1/255)
constexpr float CULL_MARGIN = 0.004f;                  // log2 units; covers fp32 rounding of the quadratic form

// Per-instance staging shared by both kernels.  s_cull = {x, y, kx, ky} with kx = -B'/(2A'), ky = -B'/(2C')
// (the maximiser of the quadratic form along a horizontal / vertical line), s_ev = {A', B', C', log2 o},
// s_col = {r, g, b, thr}: the instance can only pass alpha >= 1/255 where the form is >= thr.
S360_HD void stage_instance(const float4& r0, const float4& r1, const float4& r2, float4& cull,
                                               float4& ev, float4& col) {
  const float A = -0.5f * LOG2E * r0.z, B = -LOG2E * r0.w, C = -0.5f * LOG2E * r1.x;
#ifdef __CUDA_ARCH__
  const float lop = __log2f(r1.y);
  const float iA = __frcp_rn(A), iC = __frcp_rn(C);
#else
  const float lop = log2f(r1.y);
  const float iA = 1.f / A, iC = 1.f / C;
#endif
  ev = make_float4(A, B, C, lop);
  cull = make_float4(r0.x, r0.y, -0.5f * B * iA, -0.5f * B * iC);
  // r1.z = hx is +inf when the caller disabled tight culling: then never cull
  const float thr = (r1.z < 3.0e38f) ? (LOG2_ALPHA_MIN - CULL_MARGIN) - lop : -s360_inf();
  col = make_float4(r2.x, r2.y, r2.z, thr);
}

// max over the warp's pixel rectangle of the (concave, log2-scaled) quadratic form; d = centre - pixel,
// (ddx, ddy) = centre - rectangle centre.  Exact for the continuous rectangle hull of the pixel centres.
S360_HD bool rect_can_contribute(const float4& q, const float4& e, float thr, float ddx, float ddy) {
  const float xlo = ddx - HALF_W, xhi = ddx + HALF_W, ylo = ddy - HALF_H, yhi = ddy + HALF_H;
  const bool inx = (xlo <= 0.f) && (xhi >= 0.f), iny = (ylo <= 0.f) && (yhi >= 0.f);
  float best = (inx && iny) ? 0.f : -s360_inf();
  if (!inx) {
    const float xe = xlo > 0.f ? xlo : xhi;
    const float ys = fminf(yhi, fmaxf(ylo, q.w * xe));
    best = fmaf(fmaf(e.z, ys, e.y * xe), ys, e.x * xe * xe);
  }
  if (!iny) {
    const float ye = ylo > 0.f ? ylo : yhi;
    const float xs = fminf(xhi, fmaxf(xlo, q.z * ye));
    best = fmaxf(best, fmaf(fmaf(e.x, xs, e.y * ye), xs, e.z * ye * ye));
  }
  return !(best < thr);
}


}  // namespace s360

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

// ---- the 48-B geometry record -------------------------------------------------------------------------------
// S360_REC_STAGED = 1 (default): K1 stores what the render kernels would otherwise derive per (warp, chunk, instance)
// from the raw conic -- the log2-scaled conic A', B', C', log2(opacity) and the line maximisers kx, ky -- so that staging
// a chunk is three 128-bit loads and a handful of moves instead of a log2, two IEEE reciprocals and ~25 multiplies
// (measured: the per-chunk prologue was 21 % of the forward kernel's instructions):
//   r0 = {x, y, A', B'}   r1 = {C', log2 o, kx, ky}   r2 = {r, g, b, depth}
//   A' and C' are strictly negative (cov2D + 0.3 I is positive definite), so their sign bits carry two flags:
//   A' stored POSITIVE: tight culling disabled for this Gaussian;  C' stored POSITIVE: wider than half the panorama (erp).
//   (Colours cannot carry flags: precomputed colours -- the depth-as-colour pass -- may be negative.)
// S360_REC_STAGED = 0: r0 = {x, y, conicA, conicB}  r1 = {conicC, opacity, hx, hy}  r2 = {r, g, b, depth} (round 1).
#ifndef S360_REC_STAGED
#define S360_REC_STAGED 1
#endif

S360_HD bool s360_signbit(float x) { return (s360_float_bits(x) >> 31) != 0u; }

// hx, hy: half extents of the alpha >= 1/255 box (+inf: tight culling off); image_width only matters in erp mode
S360_HD void pack_record(float px, float py, float cA, float cB, float cC, float op, float hx, float hy, const float* col,
                         float depth, int image_width, float4& r0, float4& r1, float4& r2) {
#if S360_REC_STAGED
  (void)hy;
  const float A = -0.5f * LOG2E * cA, B = -LOG2E * cB, C = -0.5f * LOG2E * cC;
  const bool nocull = !(hx < 3.0e38f);
  const bool huge = !(hx < 0.5f * (float)image_width - (float)WARP_W);
  r0 = make_float4(px, py, nocull ? fabsf(A) : -fabsf(A), B);
  r1 = make_float4(huge ? fabsf(C) : -fabsf(C), log2f(op), -0.5f * B / A, -0.5f * B / C);
  r2 = make_float4(col[0], col[1], col[2], depth);
#else
  (void)image_width;
  r0 = make_float4(px, py, cA, cB);
  r1 = make_float4(cC, op, hx, hy);
  r2 = make_float4(col[0], col[1], col[2], depth);
#endif
}

// inverse of pack_record for the debug unpackers: conic, opacity, colour
S360_HD void unpack_record(const float4& r0, const float4& r1, const float4& r2, float* conic_op, float* rgb) {
#if S360_REC_STAGED
  conic_op[0] = fabsf(r0.z) / (0.5f * LOG2E); conic_op[1] = r0.w / -LOG2E; conic_op[2] = fabsf(r1.x) / (0.5f * LOG2E);
  conic_op[3] = exp2f(r1.y);
  rgb[0] = r2.x; rgb[1] = r2.y; rgb[2] = r2.z;
#else
  conic_op[0] = r0.z; conic_op[1] = r0.w; conic_op[2] = r1.x; conic_op[3] = r1.y;
  rgb[0] = r2.x; rgb[1] = r2.y; rgb[2] = r2.z;
#endif
}

// erp: does this Gaussian need the per-pixel seam wrap (wider than half the panorama minus a warp block)?
S360_HD bool record_is_wide(const float4& r1, const float4& r2, float halfW) {
#if S360_REC_STAGED
  (void)r2; (void)halfW;
  return !s360_signbit(r1.x);
#else
  (void)r2;
  return !(r1.z < halfW - (float)WARP_W);
#endif
}

// Per-instance staging shared by both kernels.  s_cull = {x, y, kx, ky} with kx = -B'/(2A'), ky = -B'/(2C')
// (the maximiser of the quadratic form along a horizontal / vertical line), s_ev = {A', B', C', log2 o},
// s_col = {r, g, b, thr}: the instance can only pass alpha >= 1/255 where the form is >= thr.
S360_HD void stage_instance(const float4& r0, const float4& r1, const float4& r2, float4& cull,
                                               float4& ev, float4& col) {
#if S360_REC_STAGED
  ev = make_float4(-fabsf(r0.z), r0.w, -fabsf(r1.x), r1.y);
  cull = make_float4(r0.x, r0.y, r1.z, r1.w);
  const float thr = s360_signbit(r0.z) ? (LOG2_ALPHA_MIN - CULL_MARGIN) - r1.y : -s360_inf();
  col = make_float4(r2.x, r2.y, r2.z, thr);
#else
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
#endif
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

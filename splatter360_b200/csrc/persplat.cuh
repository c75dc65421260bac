// persplat.cuh -- the per-Gaussian, per-view math of K1 and K8: covariance load/store, projection into a view
// (pinhole / equirectangular, EWA cov2D, conic, extents, tile rectangle), SH -> RGB, and the geometry backward.
// Host + device: preprocess.cu inlines these into the single-view and the batched kernels; tests/host_harness builds
// the same code for the CPU and checks it against the oracle (no GPU needed).
#pragma once
#include "common.cuh"

namespace s360 {

// covariance in: [P,6] (xx,xy,xz,yy,yz,zz) or the reference's [P,3,3] (upper triangle is read,
// cuda_splatting.py:115,123), multiplied by scene_scale^2
S360_HD void load_cov6(const S360View& v, const float* __restrict__ cov, int idx, float* cv) {
  const float s2 = v.scene_scale * v.scene_scale;
  if (v.cov_layout == 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) cv[k] = cov[6 * (size_t)idx + k] * s2;
  } else {
    const float* c = cov + 9 * (size_t)idx;
    cv[0] = c[0] * s2; cv[1] = c[1] * s2; cv[2] = c[2] * s2; cv[3] = c[4] * s2; cv[4] = c[5] * s2; cv[5] = c[8] * s2;
  }
}
// gradient out in the same layout; for [P,3,3] only the upper triangle carries gradient, exactly like autograd
// through the reference's triu gather
S360_HD void store_dcov(const S360View& v, float* __restrict__ d_cov, int idx, const float* g, float s2) {
  if (v.cov_layout == 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) d_cov[6 * (size_t)idx + k] = g[k] * s2;
  } else {
    float* d = d_cov + 9 * (size_t)idx;
    d[0] = g[0] * s2; d[1] = g[1] * s2; d[2] = g[2] * s2;
    d[3] = 0.f;       d[4] = g[3] * s2; d[5] = g[4] * s2;
    d[6] = 0.f;       d[7] = 0.f;       d[8] = g[5] * s2;
  }
}

// atan2 for the equirectangular pixel coordinates.  CUDA's atan2f is good to ~3 ulp; on a 1024-wide panorama that is
// 1e-4 px, and for sub-pixel splats the screen-space gradient sums (q dx over ~9 pixels, almost cancelling) amplify a centre
// error by 10-100x.  The double-precision atan2 rounded to float is correctly rounded (what the oracle's libm delivers) and
// costs nothing measurable in the HBM-bound K1 (S360_ERP_ATAN2_F64=0 restores atan2f).
#ifndef S360_ERP_ATAN2_F64
#define S360_ERP_ATAN2_F64 1
#endif
S360_HD float s360_atan2(float y, float x) {
#if defined(__CUDA_ARCH__) && S360_ERP_ATAN2_F64
  return (float)atan2((double)y, (double)x);
#else
  return atan2f(y, x);
#endif
}

// Result of projecting one Gaussian into one view (K1 geometry; shared by the single-view and the batched kernel).
struct Proj {
  bool upstream_visible;   // upstream's radius > 0 (tile rectangle non-empty before the tight box)
  uint32_t tiles;          // tiles kept after the alpha >= 1/255 box intersection
  uint2 rect;              // packed tile rectangle (x0 | nx << 16, y0 | ny << 16), zero when tiles == 0
  uint32_t key;            // depth sort key (0xFFFFFFFF when tiles == 0)
  int radius;
  float px, py, cA, cB, cC, op, hx, hy, sortkey;
  uint8_t cl;              // bit 3 = jacobian clamp x, bit 4 = clamp y
};

template <int MODE>
S360_HD void project_view(const S360View& v, const float* V, const float* PM, float mx, float my,
                                             float mz, const float* cv, const float* __restrict__ opac, int idx,
                                             float wf, Proj& o) {
  const int W = v.image_width, H = v.image_height;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  o.upstream_visible = false;
  o.tiles = 0; o.rect = make_uint2(0u, 0u); o.key = 0xFFFFFFFFu; o.radius = 0; o.cl = 0;
  o.px = o.py = o.cA = o.cB = o.cC = o.op = o.hx = o.hy = 0.f;
  // near cull first (upstream in_frustum): the same expressions geo_compute evaluates for the view-space centre, so
  // the sort key is bit-identical; a culled Gaussian skips the covariance projection altogether
  {
    const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
    if (MODE == S360_MODE_PINHOLE) { o.sortkey = tz; }
    else {
      const float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
      const float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
      o.sortkey = sqrtf(tx * tx + ty * ty + tz * tz);
    }
  }
  if (!(o.sortkey > v.near_cull)) return;
  if (MODE == S360_MODE_PINHOLE) {
    // Cheap conservative frustum reject (most Gaussians miss most cube faces): upstream's radius is
    // ceil(3 sqrt(lambda1)) with lambda1 <= tr(cov2D) + sqrt(0.1) and tr(cov2D) <= |J|_F^2 |W|_F^2 tr(Sigma) + 2 lowpass,
    // |J|_F^2 <= (fx^2 (1 + limx^2) + fy^2 (1 + limy^2)) / z^2 because J uses the clamped centre.  A centre farther
    // than that bound from the image has an empty tile rectangle, which is all the full path would find out.
    const float fx = (float)W / (2.f * v.tanfovx), fy = (float)H / (2.f * v.tanfovy);
    const float limx = v.fov_clamp * v.tanfovx, limy = v.fov_clamp * v.tanfovy;
    // approximate reciprocal / square root are fine here: the bound carries 1 % + 2 px of slack
    float iz, pw;
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(o.sortkey));
#else
    iz = 1.f / o.sortkey;
#endif
    const float jb = (fx * fx * (1.f + limx * limx) + fy * fy * (1.f + limy * limy)) * iz * iz;
    float rb;
#ifdef __CUDA_ARCH__
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(jb * wf * fmaxf(cv[0] + cv[3] + cv[5], 0.f) + 2.f * fabsf(v.lowpass) + 0.32f));
#else
    rb = sqrtf(jb * wf * fmaxf(cv[0] + cv[3] + cv[5], 0.f) + 2.f * fabsf(v.lowpass) + 0.32f);
#endif
    rb = 3.03f * rb + 2.f;
    const float qx = PM[0] * mx + PM[4] * my + PM[8] * mz + PM[12];
    const float qy = PM[1] * mx + PM[5] * my + PM[9] * mz + PM[13];
    const float qw = PM[3] * mx + PM[7] * my + PM[11] * mz + PM[15];
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(qw + 0.0000001f));
#else
    pw = 1.f / (qw + 0.0000001f);
#endif
    const float cx = ((qx * pw + 1.f) * W - 1.f) * 0.5f, cy = ((qy * pw + 1.f) * H - 1.f) * 0.5f;
    if (cx + rb < 0.f || cx - rb > (float)(gx * TILE) || cy + rb < 0.f || cy - rb > (float)(gy * TILE)) return;
  }
  Geo g;
  geo_compute<MODE>(v, V, mx, my, mz, cv, g);
  const float det = g.a * g.c - g.b * g.b;
  const bool alive = det != 0.f;
  if (alive) {
    const float det_inv = 1.f / det;
    o.cA = g.c * det_inv; o.cB = -g.b * det_inv; o.cC = g.a * det_inv;
    int ex, ey;
    if (MODE == S360_MODE_PINHOLE) {
      const float mid = 0.5f * (g.a + g.c);
      const float root = sqrtf(fmaxf(0.1f, mid * mid - det));
      const float lam1 = mid + root, lam2 = mid - root;
      ex = ey = (int)ceilf(3.f * sqrtf(fmaxf(lam1, lam2)));
      const float qx = PM[0] * mx + PM[4] * my + PM[8] * mz + PM[12];
      const float qy = PM[1] * mx + PM[5] * my + PM[9] * mz + PM[13];
      const float qw = PM[3] * mx + PM[7] * my + PM[11] * mz + PM[15];
      const float pw = 1.f / (qw + 0.0000001f);
      o.px = ((qx * pw + 1.f) * W - 1.f) * 0.5f;
      o.py = ((qy * pw + 1.f) * H - 1.f) * 0.5f;
    } else {
      ex = (int)ceilf(3.f * sqrtf(g.a));
      ey = (int)ceilf(3.f * sqrtf(g.c));
      if (ex > W / 2) ex = W / 2;
      const float su = -(float)W / (2.f * PI_F), sv = -(float)H / PI_F;
      o.px = su * s360_atan2(g.t[0], g.t[2]) + 0.5f * W - 0.5f;
      o.py = sv * s360_atan2(g.t[1], sqrtf(g.t[0] * g.t[0] + g.t[2] * g.t[2])) + 0.5f * H - 0.5f;
    }
    const float px = o.px, py = o.py;
    // upstream tile rectangle
    int ymin = (int)((py - ey) / TILE), ymax = (int)((py + ey + TILE - 1) / TILE);
    ymin = min(gy, max(0, ymin)); ymax = min(gy, max(0, ymax));
    int xmin, xmax;
    if (MODE == S360_MODE_PINHOLE) {
      xmin = (int)((px - ex) / TILE); xmax = (int)((px + ex + TILE - 1) / TILE);
      xmin = min(gx, max(0, xmin)); xmax = min(gx, max(0, xmax));
    } else {
      // unwrapped column range; capped to one full row only after the tight-box intersection
      xmin = (int)floorf((px - ex) / TILE); xmax = (int)floorf((px + ex + TILE - 1) / TILE);
    }
    o.upstream_visible = (xmax - xmin) * (ymax - ymin) > 0;
    if (o.upstream_visible) {
      o.radius = max(ex, ey);
      o.op = opac[idx];
      // box outside which alpha = op * exp(power) < 1/255 for certain
      o.hx = s360_inf(); o.hy = o.hx;
      if (v.tight_bbox) {
        const float tau = logf(255.f * o.op);
        if (tau > 0.f) {
          o.hx = sqrtf(2.f * tau * g.a) * 1.0005f + 1e-3f;
          o.hy = sqrtf(2.f * tau * g.c) * 1.0005f + 1e-3f;
          const int ty0 = (int)floorf((py - o.hy) / TILE), ty1 = (int)floorf((py + o.hy) / TILE) + 1;
          const int tx0 = (int)floorf((px - o.hx) / TILE), tx1 = (int)floorf((px + o.hx) / TILE) + 1;
          ymin = max(ymin, ty0); ymax = min(ymax, ty1);
          xmin = max(xmin, tx0); xmax = min(xmax, tx1);
        } else if (tau <= 0.f) {   // opacity < 1/255: can never pass the alpha test (NaN falls through)
          xmax = xmin; ymax = ymin;
        }
      }
      int nx = max(0, xmax - xmin);
      const int ny = max(0, ymax - ymin);
      if (MODE == S360_MODE_ERP) nx = min(nx, gx);
      if (nx * ny > 0) {
        o.rect = make_uint2(((uint32_t)xmin & 0xffffu) | ((uint32_t)nx << 16), (uint32_t)ymin | ((uint32_t)ny << 16));
        o.key = s360_float_bits(o.sortkey);
        o.tiles = (uint32_t)(nx * ny);
      }
      o.cl = (g.clampx ? 8 : 0) | (g.clampy ? 16 : 0);
    }
  }
}

// squared Frobenius norm of the rotation block of a view matrix (3 for a rigid camera); scales the reject bound of
// project_view so that it stays conservative for any matrix a caller hands in
S360_HD float view_frobenius2(const float* V) {
  return V[0] * V[0] + V[1] * V[1] + V[2] * V[2] + V[4] * V[4] + V[5] * V[5] + V[6] * V[6] + V[8] * V[8] + V[9] * V[9] +
         V[10] * V[10];
}

// SH -> RGB for one Gaussian (row `sh` of the staged block) seen from `campos`, together with J[c][a] = d rgb_c / d dir_a
// (w.r.t. the unnormalised view direction, before the +0.5 and the clamp); sets the clamp bits 0..2 of cl.  The ONE colour
// routine of every kernel (single-view K1, batched K1c): identical expressions with identical uses, so a Gaussian gets
// bit-identical colours whichever path renders it.  The 36-B Jacobian spares the backward pass the 300-B SH row.
S360_HD void sh_to_rgb_jac(const S360View& v, const float* sh, float mx, float my, float mz, const float* campos,
                           float* col, uint8_t& cl, float* J) {
  float dx = mx - campos[0], dy = my - campos[1], dz = mz - campos[2];
  const float inv = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
  dx *= inv; dy *= inv; dz *= inv;
  const int ks = v.sh_layout ? 1 : 3, cs = v.sh_layout ? v.M : 1;   // [P,M,3] or the reference's [P,3,M]
  sh_colour_and_jacobian(min(v.sh_degree, v.max_sh_degree), dx, dy, dz, [&](int k, int c) { return sh[ks * k + cs * c]; }, col, J);
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float r = col[ch] + 0.5f;
    if (r < 0.f) cl |= (1 << ch);
    col[ch] = fmaxf(r, 0.f);
  }
}
S360_HD void sh_to_rgb(const S360View& v, const float* sh, float mx, float my, float mz, const float* campos, float* col,
                       uint8_t& cl) {
  float J[9];
  sh_to_rgb_jac(v, sh, mx, my, mz, campos, col, cl, J);
}

// ------------------------------------------------------------------------------------------------
// Per-view geometry backward (K8): the moments the render pass accumulated for one (view, Gaussian) pair ->
// dL/d(mean3D) through the projection and the Jacobian, dL/d(cov3D), screen-space gradient dm2 (NDC units).
// a0 = {dL/dr, dL/dg, dL/db, sum q dx}, a1 = {sum q dy, sum q dx^2, sum q dxdy, sum q dy^2}, q = G dL/dalpha.
// Shared by the single-view and the batched kernel; the SH part (direction gradient) is added by the caller.
// ACCUM: add to dm / dcov instead of overwriting them (the batched kernel folds several views into one gradient set and
// has no registers to spare for per-view temporaries).
template <int MODE, bool DEPTH = false, bool ACCUM = false>
S360_HD void view_backward(const S360View& v, const float* V, const float* PM, float mx, float my,
                                              float mz, const float* cv, float op, const float4& a0, const float4& a1,
                                              float* dm, float* dm2, float* dcov, const DepthSpec& ds, float dl_dd) {
  const int W = v.image_width, H = v.image_height;
  Geo g;
  geo_compute<MODE>(v, V, mx, my, mz, cv, g);
  const float denom = g.a * g.c - g.b * g.b;
  // the render pass accumulated moments of q = G dL/dalpha:  a0.w = sum q dx, a1 = sum q {dy, dx^2, dxdy, dy^2};
  // dL/dG = o dL/dalpha turns them into the screen-space gradients
  const float det_inv = 1.f / denom;
  const float cA = g.c * det_inv, cB = -g.b * det_inv, cC = g.a * det_inv;   // conic, as in the forward pass
#if S360_BWD_QPRIME
  const float mo = 1.f;   // the moments already carry the opacity factor
#else
  const float mo = op;
#endif
  const float S1 = mo * a0.w, S2 = mo * a1.x;
  const float gu = -cA * S1 - cB * S2, gv = -cC * S2 - cB * S1;
  const float gA = -0.5f * mo * a1.y, gB = -mo * a1.z, gC = -0.5f * mo * a1.w;
  dm2[0] = gu * 0.5f * W; dm2[1] = gv * 0.5f * H;
  const float inv2 = 1.f / (denom * denom + 0.0000001f);
  const float da = inv2 * (-g.c * g.c * gA + g.b * g.c * gB + (denom - g.a * g.c) * gC);
  const float dc = inv2 * (-g.a * g.a * gC + g.a * g.b * gB + (denom - g.a * g.c) * gA);
  const float db = inv2 * (2.f * g.b * g.c * gA - (denom + 2.f * g.b * g.b) * gB + 2.f * g.a * g.b * gC);
  const float(*Mm)[3] = g.Mm;
  dcov[0] = (ACCUM ? dcov[0] : 0.f) + (Mm[0][0] * Mm[0][0] * da + Mm[0][0] * Mm[1][0] * db + Mm[1][0] * Mm[1][0] * dc);
  dcov[3] = (ACCUM ? dcov[3] : 0.f) + (Mm[0][1] * Mm[0][1] * da + Mm[0][1] * Mm[1][1] * db + Mm[1][1] * Mm[1][1] * dc);
  dcov[5] = (ACCUM ? dcov[5] : 0.f) + (Mm[0][2] * Mm[0][2] * da + Mm[0][2] * Mm[1][2] * db + Mm[1][2] * Mm[1][2] * dc);
  dcov[1] = (ACCUM ? dcov[1] : 0.f) + (2.f * Mm[0][0] * Mm[0][1] * da + (Mm[0][0] * Mm[1][1] + Mm[0][1] * Mm[1][0]) * db + 2.f * Mm[1][0] * Mm[1][1] * dc);
  dcov[2] = (ACCUM ? dcov[2] : 0.f) + (2.f * Mm[0][0] * Mm[0][2] * da + (Mm[0][0] * Mm[1][2] + Mm[0][2] * Mm[1][0]) * db + 2.f * Mm[1][0] * Mm[1][2] * dc);
  dcov[4] = (ACCUM ? dcov[4] : 0.f) + (2.f * Mm[0][2] * Mm[0][1] * da + (Mm[0][1] * Mm[1][2] + Mm[0][2] * Mm[1][1]) * db + 2.f * Mm[1][1] * Mm[1][2] * dc);
  const float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
  float dM[2][3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float Sm0 = S[k][0] * Mm[0][0] + S[k][1] * Mm[0][1] + S[k][2] * Mm[0][2];
    const float Sm1 = S[k][0] * Mm[1][0] + S[k][1] * Mm[1][1] + S[k][2] * Mm[1][2];
    dM[0][k] = 2.f * da * Sm0 + db * Sm1;
    dM[1][k] = 2.f * dc * Sm1 + db * Sm0;
  }
  // dJ[r][k] = sum_j R[k][j] dM[r][j],  R[k][j] = V[4j + k]
  float dJ[2][3];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) dJ[r][k] = V[k] * dM[r][0] + V[4 + k] * dM[r][1] + V[8 + k] * dM[r][2];
  float dt[3] = {0.f, 0.f, 0.f};
  if (!ACCUM) dm[0] = dm[1] = dm[2] = 0.f;
  if (MODE == S360_MODE_PINHOLE) {
    const float fx = (float)W / (2.f * v.tanfovx), fy = (float)H / (2.f * v.tanfovy);
    const float tz = 1.f / g.tc[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float xm = g.clampx ? 0.f : 1.f, ym = g.clampy ? 0.f : 1.f;
    dt[0] = xm * -fx * tz2 * dJ[0][2];
    dt[1] = ym * -fy * tz2 * dJ[1][2];
    dt[2] = -fx * tz2 * dJ[0][0] - fy * tz2 * dJ[1][1] + (2.f * fx * g.tc[0]) * tz3 * dJ[0][2] +
            (2.f * fy * g.tc[1]) * tz3 * dJ[1][2];
    const float hw = PM[3] * mx + PM[7] * my + PM[11] * mz + PM[15];
    const float mw = 1.f / (hw + 0.0000001f);
    const float mul1 = (PM[0] * mx + PM[4] * my + PM[8] * mz + PM[12]) * mw * mw;
    const float mul2 = (PM[1] * mx + PM[5] * my + PM[9] * mz + PM[13]) * mw * mw;
    dm[0] += (PM[0] * mw - PM[3] * mul1) * dm2[0] + (PM[1] * mw - PM[3] * mul2) * dm2[1];
    dm[1] += (PM[4] * mw - PM[7] * mul1) * dm2[0] + (PM[5] * mw - PM[7] * mul2) * dm2[1];
    dm[2] += (PM[8] * mw - PM[11] * mul1) * dm2[0] + (PM[9] * mw - PM[11] * mul2) * dm2[1];
  } else {
    const float su = -(float)W / (2.f * PI_F), sv = -(float)H / PI_F;
    const float x = g.tc[0], y = g.tc[1], z = g.tc[2];
    if (!g.clampx) {
      const float q = x * x + z * z, rho = sqrtf(q), r2 = q + y * y, q2 = q * q;
      const float f = 1.f / (rho * r2);
      const float dfx = -x * (r2 + 2.f * q) / (rho * q * r2 * r2);
      const float dfz = -z * (r2 + 2.f * q) / (rho * q * r2 * r2);
      const float dfy = -2.f * y / (rho * r2 * r2);
      const float dJ00x = -2.f * su * x * z / q2, dJ00z = su * (x * x - z * z) / q2;
      const float dJ02x = su * (x * x - z * z) / q2, dJ02z = 2.f * su * x * z / q2;
      const float dJ10x = -sv * y * (f + x * dfx), dJ10y = -sv * x * (f + y * dfy), dJ10z = -sv * x * y * dfz;
      const float dJ12x = -sv * z * y * dfx, dJ12y = -sv * z * (f + y * dfy), dJ12z = -sv * y * (f + z * dfz);
      const float dJ11x = sv * x * (r2 - 2.f * q) / (rho * r2 * r2), dJ11z = sv * z * (r2 - 2.f * q) / (rho * r2 * r2);
      const float dJ11y = -2.f * sv * rho * y / (r2 * r2);
      dt[0] = dJ[0][0] * dJ00x + dJ[0][2] * dJ02x + dJ[1][0] * dJ10x + dJ[1][1] * dJ11x + dJ[1][2] * dJ12x;
      dt[1] = dJ[1][0] * dJ10y + dJ[1][1] * dJ11y + dJ[1][2] * dJ12y;
      dt[2] = dJ[0][0] * dJ00z + dJ[0][2] * dJ02z + dJ[1][0] * dJ10z + dJ[1][1] * dJ11z + dJ[1][2] * dJ12z;
    }
    dt[0] += g.J[0][0] * gu + g.J[1][0] * gv;
    dt[1] += g.J[0][1] * gu + g.J[1][1] * gv;
    dt[2] += g.J[0][2] * gu + g.J[1][2] * gv;
  }
  if (DEPTH) {
    // fused depth channel: dl_dd = dL/d(depth value of this Gaussian); the value is a function of the sort depth
    // (camera z, or radial distance in erp mode), whose view-space gradient is e_z resp. t / |t|
    if (MODE == S360_MODE_PINHOLE) {
      dt[2] += dl_dd * depth_value_grad(ds, g.t[2]);
    } else {
      const float r = sqrtf(g.t[0] * g.t[0] + g.t[1] * g.t[1] + g.t[2] * g.t[2]);
      const float coef = dl_dd * depth_value_grad(ds, r) / r;
      dt[0] += coef * g.t[0]; dt[1] += coef * g.t[1]; dt[2] += coef * g.t[2];
    }
  }
  // mean3D <- view-space gradient: dm_k += sum_i R[i][k] dt_i,  R[i][k] = V[4k + i]
#pragma unroll
  for (int k = 0; k < 3; k++) dm[k] += V[4 * k] * dt[0] + V[4 * k + 1] * dt[1] + V[4 * k + 2] * dt[2];
}


}  // namespace s360

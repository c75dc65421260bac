// render.cu -- per-tile compositing kernels.
//   K6 render_forward_kernel : front-to-back alpha compositing            
//   K7 render_backward_kernel: back-to-front gradient pass                 
// One CTA = one 16x16 tile = 8 warps; a warp owns an 8x4 pixel block.  A batch of sorted instances
// is gathered into shared memory once per tile; every warp first tests, one instance per lane,
// whether the instance's alpha>=1/255 box can reach its 8x4 block (ballot) and then evaluates only
// the survivors: centre broadcast by SHFL from the testing lane, conic broadcast from shared memory.
// The conic is staged pre-multiplied by -0.5*log2(e) and the opacity as log2(opacity), so that
// alpha = ex2(A'dx^2 + C'dy^2 + B'dxdy + log2 o) costs 5 FP32 ops + one MUFU.EX2.
//
// Backward: per (pixel, instance) only the moments q, q*dx, q*dy, q*dx^2, q*dxdy, q*dy^2
// (q = G * dL/dalpha) and the three colour terms are formed; they are summed over the warp with a
// transposed butterfly (14 shuffles), written to a per-warp private slot in shared memory (no
// atomics), and converted to dL/d{mean2D, conic, opacity} once per instance and tile before a
// single global atomic per value.  Replaces upstream renderCUDA (forward.cu / backward.cu) behind
// /root/reference/src/model/decoder/cuda_splatting.py:113-124.
#include "common.cuh"

namespace s360 {

constexpr int RT = 128;                 // threads per tile: 4 warps, each owns an 8x8 pixel block
constexpr int WARP_W = 8, WARP_H = 8;   // pixel block of one warp; a lane owns (x, y) and (x, y + 4)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float HALF_W = 0.5f * (WARP_W - 1), HALF_H = 0.5f * (WARP_H - 1);

template <int MODE>
__device__ __forceinline__ float wrap_dx(float dx, float W, float halfW) {
  if (MODE == S360_MODE_ERP) {
    if (dx > halfW) dx -= W;
    else if (dx < -halfW) dx += W;
  }
  return dx;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifndef S360_EXACT_CULL
#define S360_EXACT_CULL 1
#endif
constexpr float LOG2_ALPHA_MIN = -7.99435343685886f;   // log2(1/255)
constexpr float CULL_MARGIN = 0.004f;                  // log2 units; covers fp32 rounding of the quadratic form

// Per-instance staging shared by both kernels.  s_cull = {x, y, kx, ky} with kx = -B'/(2A'), ky = -B'/(2C')
// (the maximiser of the quadratic form along a horizontal / vertical line), s_ev = {A', B', C', log2 o},
// s_col = {r, g, b, thr}: the instance can only pass alpha >= 1/255 where the form is >= thr.
__device__ __forceinline__ void stage_instance(const float4& r0, const float4& r1, const float4& r2, float4& cull,
                                               float4& ev, float4& col) {
  const float A = -0.5f * LOG2E * r0.z, B = -LOG2E * r0.w, C = -0.5f * LOG2E * r1.x;
  const float lop = __log2f(r1.y);
  ev = make_float4(A, B, C, lop);
  cull = make_float4(r0.x, r0.y, -0.5f * B * __frcp_rn(A), -0.5f * B * __frcp_rn(C));
  // r1.z = hx is +inf when the caller disabled tight culling: then never cull
  const float thr = (r1.z < 3.0e38f) ? (LOG2_ALPHA_MIN - CULL_MARGIN) - lop : -__int_as_float(0x7f800000);
  col = make_float4(r2.x, r2.y, r2.z, thr);
}

// max over the warp's pixel rectangle of the (concave, log2-scaled) quadratic form; d = centre - pixel,
// (ddx, ddy) = centre - rectangle centre.  Exact for the continuous rectangle hull of the pixel centres.
__device__ __forceinline__ bool rect_can_contribute(const float4& q, const float4& e, float thr, float ddx, float ddy) {
  const float xlo = ddx - HALF_W, xhi = ddx + HALF_W, ylo = ddy - HALF_H, yhi = ddy + HALF_H;
  const bool inx = (xlo <= 0.f) && (xhi >= 0.f), iny = (ylo <= 0.f) && (yhi >= 0.f);
  float best = (inx && iny) ? 0.f : -__int_as_float(0x7f800000);
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

#ifndef S360_FWD_MINB
#define S360_FWD_MINB 1
#endif
#ifndef S360_BWD_MINB
#define S360_BWD_MINB 1
#endif

constexpr int NWARPS = RT / 32;

// Warp-autonomous streaming of a tile's instance list: lane j of every warp gathers instance
// (chunk*32 + j) itself (the four warps of a tile hit the same lines in L1), two chunks of ids and one
// chunk of records are always in flight, and no CTA-wide barrier exists in the kernel.
struct ChunkRegs {
  float4 r0, r1, r2;
  uint32_t gid;
};

__device__ __forceinline__ void load_records(ChunkRegs& c, const float4* __restrict__ rec, bool valid) {
  if (valid) {
    c.r0 = __ldg(rec + 3 * (size_t)c.gid);
    c.r1 = __ldg(rec + 3 * (size_t)c.gid + 1);
    c.r2 = __ldg(rec + 3 * (size_t)c.gid + 2);
  }
}

template <int MODE>
__global__ void __launch_bounds__(RT, S360_FWD_MINB)
render_forward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                      const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color) {
  __shared__ float4 s_ev[NWARPS][32];    // A', B', C' (log2-scaled conic), log2(opacity)
  __shared__ float4 s_col[NWARPS][32];   // r, g, b, -
  const int gx = (W + TILE - 1) / TILE;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py0 = wy0 + (lane >> 3), py1 = py0 + 4;
  const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
  const float pxf = (float)px, pyf = (float)py0;
  const float wcx = (float)wx0 + HALF_W, wcy = (float)wy0 + HALF_H;
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];

  float T0 = 1.f, T1 = 1.f, Ca0 = 0.f, Ca1 = 0.f, Ca2 = 0.f, Cb0 = 0.f, Cb1 = 0.f, Cb2 = 0.f;
  uint32_t last0 = 0, last1 = 0;
  bool done0 = !in0, done1 = !in1;

  ChunkRegs nx;
  nx.r0 = nx.r1 = nx.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
  nx.gid = 0;
  if (range.x + lane < range.y) nx.gid = point_list[range.x + lane];
  load_records(nx, rec, range.x + lane < range.y);
  uint32_t gid2 = (range.x + 32 + lane < range.y) ? point_list[range.x + 32 + lane] : 0u;

  for (uint32_t base = range.x; base < range.y; base += 32) {
    if (__all_sync(0xffffffffu, done0 && done1)) break;
    const bool valid = base + lane < range.y;
    float4 cull, ev, col;
    stage_instance(nx.r0, nx.r1, nx.r2, cull, ev, col);
    const float thr = col.w;
    const bool huge = MODE == S360_MODE_ERP && !(nx.r1.z < halfW - (float)WARP_W);
    s_ev[warp][lane] = ev;
    s_col[warp][lane] = col;
    __syncwarp();
    // keep the pipeline full: records of the next chunk, ids of the one after
    nx.gid = gid2;
    load_records(nx, rec, base + 32 + lane < range.y);
    gid2 = (base + 64 + lane < range.y) ? point_list[base + 64 + lane] : 0u;

    const float ddx = wrap_dx<MODE>(cull.x - wcx, Wf, halfW), ddy = cull.y - wcy;
    const bool hit = valid && (huge || rect_can_contribute(cull, ev, thr, ddx, ddy));
    const float cx = MODE == S360_MODE_ERP ? wcx + ddx : cull.x;   // nearest periodic copy w.r.t. this warp
    const float cy = cull.y;
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    const unsigned hmask = MODE == S360_MODE_ERP ? __ballot_sync(0xffffffffu, hit && huge) : 0u;
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      const float xs = __shfl_sync(0xffffffffu, cx, k), ys = __shfl_sync(0xffffffffu, cy, k);
      const float4 e = s_ev[warp][k];
      float dx = xs - pxf;
      if (MODE == S360_MODE_ERP && ((hmask >> k) & 1u)) dx = wrap_dx<MODE>(dx, Wf, halfW);
      const float dy0 = ys - pyf, dy1 = dy0 - 4.f;
      const float u = e.y * dx;
      const float pb = e.x * dx * dx;                            // A'dx^2
      const float p0 = fmaf(fmaf(e.z, dy0, u), dy0, pb);         // + C'dy^2 + B'dxdy  (= power * log2 e)
      const float p1 = fmaf(fmaf(e.z, dy1, u), dy1, pb);
      const float al0 = fminf(ALPHA_MAX, ex2_approx(p0 + e.w)), al1 = fminf(ALPHA_MAX, ex2_approx(p1 + e.w));
      bool ok0 = !done0 && (p0 <= 0.f) && (al0 >= ALPHA_MIN);
      bool ok1 = !done1 && (p1 <= 0.f) && (al1 >= ALPHA_MIN);
      const float tt0 = T0 * (1.f - al0), tt1 = T1 * (1.f - al1);
      if (ok0 && tt0 < T_EPS) { done0 = true; ok0 = false; }
      if (ok1 && tt1 < T_EPS) { done1 = true; ok1 = false; }
      if (ok0 || ok1) {
        const float4 c = s_col[warp][k];
        const uint32_t pos = base - range.x + (uint32_t)k + 1u;
        if (ok0) { const float w = al0 * T0; Ca0 += c.x * w; Ca1 += c.y * w; Ca2 += c.z * w; T0 = tt0; last0 = pos; }
        if (ok1) { const float w = al1 * T1; Cb0 += c.x * w; Cb1 += c.y * w; Cb2 += c.z * w; T1 = tt1; last1 = pos; }
      }
    }
    __syncwarp();
  }
  const size_t plane = (size_t)H * W;
  const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
  if (in0) {
    const size_t pid = (size_t)py0 * W + px;
    final_T[pid] = T0; n_contrib[pid] = last0;
    out_color[pid] = Ca0 + T0 * b0; out_color[plane + pid] = Ca1 + T0 * b1; out_color[2 * plane + pid] = Ca2 + T0 * b2;
  }
  if (in1) {
    const size_t pid = (size_t)py1 * W + px;
    final_T[pid] = T1; n_contrib[pid] = last1;
    out_color[pid] = Cb0 + T1 * b0; out_color[plane + pid] = Cb1 + T1 * b1; out_color[2 * plane + pid] = Cb2 + T1 * b2;
  }
}

int launch_render_forward(const S360View& v, GeomState g, const uint32_t* point_list, ImageState img,
                          float* out_color, cudaStream_t st) {
  const int W = v.image_width, H = v.image_height;
  const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  if (tiles == 0) return 0;
  if (v.mode == S360_MODE_PINHOLE)
    render_forward_kernel<S360_MODE_PINHOLE><<<tiles, RT, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, out_color);
  else
    render_forward_kernel<S360_MODE_ERP><<<tiles, RT, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, out_color);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Sum eight per-lane values across the warp with a transposed butterfly: after the call every lane
// holds the complete sum of value index ((lane >> 2) & 7).  9 shuffles instead of 40.
__device__ __forceinline__ float warp_reduce8(float v0, float v1, float v2, float v3, float v4, float v5,
                                              float v6, float v7, int lane) {
  const unsigned F = 0xffffffffu;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  float w0 = (h16 ? v4 : v0) + __shfl_xor_sync(F, h16 ? v0 : v4, 16);
  float w1 = (h16 ? v5 : v1) + __shfl_xor_sync(F, h16 ? v1 : v5, 16);
  float w2 = (h16 ? v6 : v2) + __shfl_xor_sync(F, h16 ? v2 : v6, 16);
  float w3 = (h16 ? v7 : v3) + __shfl_xor_sync(F, h16 ? v3 : v7, 16);
  float u0 = (h8 ? w2 : w0) + __shfl_xor_sync(F, h8 ? w0 : w2, 8);
  float u1 = (h8 ? w3 : w1) + __shfl_xor_sync(F, h8 ? w1 : w3, 8);
  float s = (h4 ? u1 : u0) + __shfl_xor_sync(F, h4 ? u0 : u1, 4);
  s += __shfl_xor_sync(F, s, 2);
  s += __shfl_xor_sync(F, s, 1);
  return s;  // value index = 4*bit4 + 2*bit3 + bit2 of the lane id = (lane >> 2) & 7
}

constexpr int NACC = 9;         // colour x3, q*dx, q*dy, q*dx^2, q*dxdy, q*dy^2, q

// per-pixel running state of the back-to-front pass
struct PixB {
  float T, T_final, bgT;
  float ar0, ar1, ar2, lc0, lc1, lc2, last_alpha;
  float dp0, dp1, dp2;
  uint32_t last_contributor;
};

__device__ __forceinline__ void pix_init(PixB& p, bool inside, size_t pid, size_t plane, const float* final_T,
                                         const uint32_t* n_contrib, const float* dL, const float* bg) {
  p.T_final = inside ? final_T[pid] : 0.f;
  p.last_contributor = inside ? n_contrib[pid] : 0u;
  p.dp0 = p.dp1 = p.dp2 = 0.f;
  if (inside) { p.dp0 = dL[pid]; p.dp1 = dL[plane + pid]; p.dp2 = dL[2 * plane + pid]; }
  p.bgT = -p.T_final * (bg[0] * p.dp0 + bg[1] * p.dp1 + bg[2] * p.dp2);
  p.T = p.T_final;
  p.ar0 = p.ar1 = p.ar2 = p.lc0 = p.lc1 = p.lc2 = p.last_alpha = 0.f;
}

// one (pixel, instance) term; accumulates the nine partials into v[]
__device__ __forceinline__ void pix_term(PixB& p, float alpha, float G, float dx, float dy, const float4& c,
                                         float* v) {
  const float inv1ma = __frcp_rn(1.f - alpha);
  p.T = p.T * inv1ma;
  const float w = alpha * p.T;
  p.ar0 = p.last_alpha * p.lc0 + (1.f - p.last_alpha) * p.ar0;
  p.ar1 = p.last_alpha * p.lc1 + (1.f - p.last_alpha) * p.ar1;
  p.ar2 = p.last_alpha * p.lc2 + (1.f - p.last_alpha) * p.ar2;
  p.lc0 = c.x; p.lc1 = c.y; p.lc2 = c.z;
  float dL_dalpha = (c.x - p.ar0) * p.dp0 + (c.y - p.ar1) * p.dp1 + (c.z - p.ar2) * p.dp2;
  dL_dalpha = dL_dalpha * p.T + p.bgT * inv1ma;
  p.last_alpha = alpha;
  const float q = G * dL_dalpha;
  const float qx = q * dx, qy = q * dy;
  v[0] += w * p.dp0; v[1] += w * p.dp1; v[2] += w * p.dp2;
  v[3] += qx; v[4] += qy; v[5] += qx * dx; v[6] += qx * dy; v[7] += qy * dy;
  v[8] += q;
}

template <int MODE>
__global__ void __launch_bounds__(RT, S360_BWD_MINB)
render_backward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                       const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dcolor, float* __restrict__ acc) {
  __shared__ float4 s_ev[NWARPS][32];    // A', B', C', log2(opacity)
  __shared__ float4 s_col[NWARPS][32];   // r, g, b, bits of the Gaussian id
  const int gx = (W + TILE - 1) / TILE;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py0 = wy0 + (lane >> 3), py1 = py0 + 4;
  const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
  const float pxf = (float)px, pyf = (float)py0;
  const float wcx = (float)wx0 + HALF_W, wcy = (float)wy0 + HALF_H;
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];
  const size_t plane = (size_t)H * W;

  PixB P0, P1;
  pix_init(P0, in0, (size_t)py0 * W + px, plane, final_T, n_contrib, dL_dcolor, bg);
  pix_init(P1, in1, (size_t)py1 * W + px, plane, final_T, n_contrib, dL_dcolor, bg);
  // instances [0, todo) of this tile's list can matter to this warp's 64 pixels
  const uint32_t todo = __reduce_max_sync(0xffffffffu, max(P0.last_contributor, P1.last_contributor));
  const int nchunks = (int)((todo + 31u) >> 5);

  ChunkRegs nx;
  nx.r0 = nx.r1 = nx.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
  nx.gid = 0;
  uint32_t gid2 = 0;
  if (nchunks > 0) {
    const uint32_t p = (uint32_t)(nchunks - 1) * 32u + lane;
    if (p < todo) nx.gid = point_list[range.x + p];
    load_records(nx, rec, p < todo);
    if (nchunks > 1) gid2 = point_list[range.x + p - 32u];
  }
  for (int ci = nchunks - 1; ci >= 0; --ci) {
    const uint32_t pos0 = (uint32_t)ci * 32u;
    const bool valid = pos0 + lane < todo;
    float4 cull, ev, col;
    stage_instance(nx.r0, nx.r1, nx.r2, cull, ev, col);
    const float thr = col.w;
    const bool huge = MODE == S360_MODE_ERP && !(nx.r1.z < halfW - (float)WARP_W);
    col.w = __uint_as_float(nx.gid);
    s_ev[warp][lane] = ev;
    s_col[warp][lane] = col;
    __syncwarp();
    nx.gid = gid2;                       // chunks below the last one are always full
    load_records(nx, rec, ci > 0);
    gid2 = (ci > 1) ? point_list[range.x + pos0 - 64u + lane] : 0u;

    const float ddx = wrap_dx<MODE>(cull.x - wcx, Wf, halfW), ddy = cull.y - wcy;
    const bool hit = valid && (huge || rect_can_contribute(cull, ev, thr, ddx, ddy));
    const float cx = MODE == S360_MODE_ERP ? wcx + ddx : cull.x;
    const float cy = cull.y;
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    const unsigned hmask = MODE == S360_MODE_ERP ? __ballot_sync(0xffffffffu, hit && huge) : 0u;
    while (mask) {
      const int k = 31 - __clz(mask);
      mask &= ~(1u << k);
      const uint32_t pos = pos0 + (uint32_t)k;
      const float xs = __shfl_sync(0xffffffffu, cx, k), ys = __shfl_sync(0xffffffffu, cy, k);
      const float4 e = s_ev[warp][k];
      float dx = xs - pxf;
      if (MODE == S360_MODE_ERP && ((hmask >> k) & 1u)) dx = wrap_dx<MODE>(dx, Wf, halfW);
      const float dy0 = ys - pyf, dy1 = dy0 - 4.f;
      const float u = e.y * dx;
      const float pb = e.x * dx * dx;
      const float p0 = fmaf(fmaf(e.z, dy0, u), dy0, pb);
      const float p1 = fmaf(fmaf(e.z, dy1, u), dy1, pb);
      // same expression as the forward pass, so that T / (1 - alpha) undoes exactly what it applied
      const float al0 = fminf(ALPHA_MAX, ex2_approx(p0 + e.w)), al1 = fminf(ALPHA_MAX, ex2_approx(p1 + e.w));
      const bool ok0 = (pos < P0.last_contributor) && (p0 <= 0.f) && (al0 >= ALPHA_MIN);
      const bool ok1 = (pos < P1.last_contributor) && (p1 <= 0.f) && (al1 >= ALPHA_MIN);
      if (!__any_sync(0xffffffffu, ok0 || ok1)) continue;
      const float4 c = s_col[warp][k];
      float v[NACC];
#pragma unroll
      for (int i = 0; i < NACC; i++) v[i] = 0.f;
      if (ok1) pix_term(P1, al1, ex2_approx(p1), dx, dy1, c, v);
      if (ok0) pix_term(P0, al0, ex2_approx(p0), dx, dy0, c, v);
      const float s8 = warp_reduce8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], lane);
      float v8 = v[8];
      v8 += __shfl_xor_sync(0xffffffffu, v8, 16);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 8);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 4);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 2);
      v8 += __shfl_xor_sync(0xffffffffu, v8, 1);
      // nine moment sums go straight to the per-Gaussian accumulator (fire-and-forget RED);
      // lanes 0,4,..,28 hold sums 0..7, lane 1 holds sum 8.  Conversion to gradients happens once per
      // Gaussian in preprocess_backward_kernel.
      float* dst = acc + (size_t)__float_as_uint(c.w) * ACC_STRIDE;
      if ((lane & 3) == 0) atomicAdd(dst + ((lane >> 2) & 7), s8);
      if (lane == 1) atomicAdd(dst + 8, v8);
    }
    __syncwarp();
  }
}

int launch_render_backward(const S360View& v, GeomState g, const uint32_t* point_list, ImageState img,
                           const float* dL_dcolor, float* acc, cudaStream_t st) {
  const int W = v.image_width, H = v.image_height;
  const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  if (tiles == 0) return 0;
  if (v.mode == S360_MODE_PINHOLE)
    render_backward_kernel<S360_MODE_PINHOLE><<<tiles, RT, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, dL_dcolor, acc);
  else
    render_backward_kernel<S360_MODE_ERP><<<tiles, RT, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, dL_dcolor, acc);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace s360

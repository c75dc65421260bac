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

constexpr int WARP_W = 8, WARP_H = 4;  // pixel block of one warp
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

constexpr int FWD_BATCH = TILE_PIX;

template <int MODE>
__global__ void __launch_bounds__(TILE_PIX)
render_forward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                      const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color) {
  __shared__ float4 s_cull[FWD_BATCH];  // x, y, hx, hy
  __shared__ float4 s_ev[FWD_BATCH];    // A', B', C' (log2-scaled conic), log2(opacity)
  __shared__ float4 s_col[FWD_BATCH];   // r, g, b, -
  const int gx = (W + TILE - 1) / TILE;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wcx = (float)wx0 + HALF_W, wcy = (float)wy0 + HALF_H;
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];

  float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last = 0;
  bool done = !inside;

  for (uint32_t base = range.x; base < range.y; base += FWD_BATCH) {
    if (__syncthreads_and(done)) break;
    const uint32_t i = base + tid;
    if (i < range.y) {
      const size_t gid = point_list[i];
      const float4 r0 = __ldg(rec + 3 * gid);
      const float4 r1 = __ldg(rec + 3 * gid + 1);
      const float4 r2 = __ldg(rec + 3 * gid + 2);
      s_cull[tid] = make_float4(r0.x, r0.y, r1.z, r1.w);
      s_ev[tid] = make_float4(-0.5f * LOG2E * r0.z, -LOG2E * r0.w, -0.5f * LOG2E * r1.x, __log2f(r1.y));
      s_col[tid] = r2;
    }
    __syncthreads();
    const int cnt = min((uint32_t)FWD_BATCH, range.y - base);
    if (__all_sync(0xffffffffu, done)) continue;
    for (int c0 = 0; c0 < cnt; c0 += 32) {
      const int j = c0 + lane;
      bool hit = false, huge = false;
      float cx = 0.f, cy = 0.f;
      if (j < cnt) {
        const float4 q = s_cull[j];
        const float ddx = wrap_dx<MODE>(q.x - wcx, Wf, halfW), ddy = q.y - wcy;
        hit = !(fabsf(ddx) - HALF_W > q.z) && !(fabsf(ddy) - HALF_H > q.w);
        cx = MODE == S360_MODE_ERP ? wcx + ddx : q.x;   // nearest periodic copy w.r.t. this warp
        cy = q.y;
        huge = MODE == S360_MODE_ERP && !(q.z < halfW - (float)WARP_W);
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      const unsigned hmask = MODE == S360_MODE_ERP ? __ballot_sync(0xffffffffu, hit && huge) : 0u;
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const float xs = __shfl_sync(0xffffffffu, cx, k), ys = __shfl_sync(0xffffffffu, cy, k);
        const float4 e = s_ev[c0 + k];
        float dx = xs - pxf;
        const float dy = ys - pyf;
        if (MODE == S360_MODE_ERP && ((hmask >> k) & 1u)) dx = wrap_dx<MODE>(dx, Wf, halfW);
        const float p2 = e.x * dx * dx + e.z * dy * dy + e.y * dx * dy;
        const float alpha = fminf(ALPHA_MAX, ex2_approx(p2 + e.w));
        bool ok = !done && (p2 <= 0.f) && (alpha >= ALPHA_MIN);
        const float test_T = T * (1.f - alpha);
        if (ok && test_T < T_EPS) { done = true; ok = false; }
        if (ok) {
          const float4 c = s_col[c0 + k];
          const float w = alpha * T;
          C0 += c.x * w; C1 += c.y * w; C2 += c.z * w;
          T = test_T;
          last = base - range.x + (uint32_t)(c0 + k) + 1u;
        }
      }
      if (__all_sync(0xffffffffu, done)) break;
    }
  }
  if (inside) {
    const size_t pid = (size_t)py * W + px, plane = (size_t)H * W;
    final_T[pid] = T;
    n_contrib[pid] = last;
    out_color[pid] = C0 + T * bg[0];
    out_color[plane + pid] = C1 + T * bg[1];
    out_color[2 * plane + pid] = C2 + T * bg[2];
  }
}

int launch_render_forward(const S360View& v, GeomState g, const uint32_t* point_list, ImageState img,
                          float* out_color, cudaStream_t st) {
  const int W = v.image_width, H = v.image_height;
  const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  if (tiles == 0) return 0;
  if (v.mode == S360_MODE_PINHOLE)
    render_forward_kernel<S360_MODE_PINHOLE><<<tiles, TILE_PIX, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, out_color);
  else
    render_forward_kernel<S360_MODE_ERP><<<tiles, TILE_PIX, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, out_color);
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
constexpr int BWD_BATCH = 128;  // instances staged per round
constexpr int NWARPS = TILE_PIX / 32;

template <int MODE>
__global__ void __launch_bounds__(TILE_PIX)
render_backward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                       const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dcolor, float* __restrict__ acc) {
  __shared__ float4 s_cull[BWD_BATCH];  // x, y, hx, hy
  __shared__ float4 s_ev[BWD_BATCH];    // A', B', C', log2(opacity)
  __shared__ float4 s_col[BWD_BATCH];   // r, g, b, -
  __shared__ float4 s_raw[BWD_BATCH];   // A, B, C, opacity (unscaled, for the per-instance conversion)
  __shared__ uint32_t s_gid[BWD_BATCH];
  __shared__ float s_acc[NWARPS][BWD_BATCH][NACC];  // per-warp private partial sums
  __shared__ __align__(16) uint8_t s_vis[NWARPS][BWD_BATCH];      // 1 if that warp produced a partial for that instance
  __shared__ uint32_t s_max;
  const int gx = (W + TILE - 1) / TILE;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wcx = (float)wx0 + HALF_W, wcy = (float)wy0 + HALF_H;
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];
  const size_t pid = (size_t)py * W + px, plane = (size_t)H * W;

  const float T_final = inside ? final_T[pid] : 0.f;
  const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;
  float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
  if (inside) { dp0 = dL_dcolor[pid]; dp1 = dL_dcolor[plane + pid]; dp2 = dL_dcolor[2 * plane + pid]; }
  const float bgT = -T_final * (bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2);

  if (tid == 0) s_max = 0;
  __syncthreads();
  const uint32_t warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
  if (lane == 0) atomicMax(&s_max, warp_max);
  __syncthreads();
  const uint32_t todo = s_max;  // instances [0, todo) of this tile's list can matter

  float T = T_final;
  float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;     // accum_rec
  float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;     // last_color
  float last_alpha = 0.f;

  const int rounds = (int)((todo + BWD_BATCH - 1) / BWD_BATCH);
  for (int bi = rounds - 1; bi >= 0; --bi) {
    __syncthreads();
    const uint32_t pos0 = (uint32_t)bi * BWD_BATCH;
    const int cnt = (int)min((uint32_t)BWD_BATCH, todo - pos0);
    {
      const int j = tid & (BWD_BATCH - 1);
      if (j < cnt) {
        const size_t gid = point_list[range.x + pos0 + j];
        if (tid < BWD_BATCH) {
          const float4 r0 = __ldg(rec + 3 * gid);
          const float4 r1 = __ldg(rec + 3 * gid + 1);
          s_cull[j] = make_float4(r0.x, r0.y, r1.z, r1.w);
          s_ev[j] = make_float4(-0.5f * LOG2E * r0.z, -LOG2E * r0.w, -0.5f * LOG2E * r1.x, __log2f(r1.y));
          s_raw[j] = make_float4(r0.z, r0.w, r1.x, r1.y);
        } else {
          s_col[j] = __ldg(rec + 3 * gid + 2);
          s_gid[j] = (uint32_t)gid;
        }
      }
      // clear the visited flags: NWARPS * BWD_BATCH bytes = 256 x uint32
      reinterpret_cast<uint32_t*>(&s_vis[0][0])[tid] = 0u;
    }
    __syncthreads();
    for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
      if (pos0 + (uint32_t)c0 >= warp_max) continue;
      const int j = c0 + lane;
      bool hit = false, huge = false;
      float cx = 0.f, cy = 0.f;
      if (j < cnt && pos0 + (uint32_t)j < warp_max) {
        const float4 q = s_cull[j];
        const float ddx = wrap_dx<MODE>(q.x - wcx, Wf, halfW), ddy = q.y - wcy;
        hit = !(fabsf(ddx) - HALF_W > q.z) && !(fabsf(ddy) - HALF_H > q.w);
        cx = MODE == S360_MODE_ERP ? wcx + ddx : q.x;
        cy = q.y;
        huge = MODE == S360_MODE_ERP && !(q.z < halfW - (float)WARP_W);
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      const unsigned hmask = MODE == S360_MODE_ERP ? __ballot_sync(0xffffffffu, hit && huge) : 0u;
      while (mask) {
        const int k = 31 - __clz(mask);
        mask &= ~(1u << k);
        const int jj = c0 + k;
        const float xs = __shfl_sync(0xffffffffu, cx, k), ys = __shfl_sync(0xffffffffu, cy, k);
        const float4 e = s_ev[jj];
        float dx = xs - pxf;
        const float dy = ys - pyf;
        if (MODE == S360_MODE_ERP && ((hmask >> k) & 1u)) dx = wrap_dx<MODE>(dx, Wf, halfW);
        const float p2 = e.x * dx * dx + e.z * dy * dy + e.y * dx * dy;
        const float G = ex2_approx(p2);
        const float alpha = fminf(ALPHA_MAX, ex2_approx(p2 + e.w));
        const bool ok = (pos0 + (uint32_t)jj < last_contributor) && (p2 <= 0.f) && (alpha >= ALPHA_MIN);
        if (!__any_sync(0xffffffffu, ok)) continue;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
        if (ok) {
          const float4 c = s_col[jj];
          const float inv1ma = __frcp_rn(1.f - alpha);
          T = T * inv1ma;
          const float w = alpha * T;
          ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0;
          ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1;
          ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2;
          lc0 = c.x; lc1 = c.y; lc2 = c.z;
          float dL_dalpha = (c.x - ar0) * dp0 + (c.y - ar1) * dp1 + (c.z - ar2) * dp2;
          dL_dalpha = dL_dalpha * T + bgT * inv1ma;
          last_alpha = alpha;
          const float q = G * dL_dalpha;
          const float qx = q * dx, qy = q * dy;
          v0 = w * dp0; v1 = w * dp1; v2 = w * dp2;
          v3 = qx; v4 = qy; v5 = qx * dx; v6 = qx * dy; v7 = qy * dy;
          v8 = q;
        }
        const float s8 = warp_reduce8(v0, v1, v2, v3, v4, v5, v6, v7, lane);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 16);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 8);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 4);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 2);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 1);
        // each (warp, instance) pair is visited at most once per round: plain stores, no atomics
        float* slot = &s_acc[warp][jj][0];
        if ((lane & 3) == 0) slot[(lane >> 2) & 7] = s8;
        if (lane == 1) slot[8] = v8;
        if (lane == 2) s_vis[warp][jj] = 1;
      }
    }
    __syncthreads();
    if (tid < cnt) {
      float sum[NACC];
#pragma unroll
      for (int k = 0; k < NACC; k++) sum[k] = 0.f;
      bool any = false;
#pragma unroll
      for (int w = 0; w < NWARPS; w++) {
        if (s_vis[w][tid]) {
          any = true;
#pragma unroll
          for (int k = 0; k < NACC; k++) sum[k] += s_acc[w][tid][k];
        }
      }
      if (any) {
        // moments -> gradients: dL/dG = o * dL/dalpha, q = G dL/dalpha
        const float4 r = s_raw[tid];  // A, B, C, opacity
        const float o = r.w;
        const float S1 = o * sum[3], S2 = o * sum[4];
        float* dst = acc + (size_t)s_gid[tid] * ACC_STRIDE;
        atomicAdd(dst + 0, sum[0]);
        atomicAdd(dst + 1, sum[1]);
        atomicAdd(dst + 2, sum[2]);
        atomicAdd(dst + 3, -r.x * S1 - r.y * S2);       // dL/du   (pixel units)
        atomicAdd(dst + 4, -r.z * S2 - r.y * S1);       // dL/dv
        atomicAdd(dst + 5, -0.5f * o * sum[5]);         // dL/dconicA
        atomicAdd(dst + 6, -o * sum[6]);                // dL/dconicB (true off-diagonal gradient)
        atomicAdd(dst + 7, -0.5f * o * sum[7]);         // dL/dconicC
        atomicAdd(dst + 8, sum[8]);                     // dL/dopacity
      }
    }
  }
}

int launch_render_backward(const S360View& v, GeomState g, const uint32_t* point_list, ImageState img,
                           const float* dL_dcolor, float* acc, cudaStream_t st) {
  const int W = v.image_width, H = v.image_height;
  const int tiles = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
  if (tiles == 0) return 0;
  if (v.mode == S360_MODE_PINHOLE)
    render_backward_kernel<S360_MODE_PINHOLE><<<tiles, TILE_PIX, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, dL_dcolor, acc);
  else
    render_backward_kernel<S360_MODE_ERP><<<tiles, TILE_PIX, 0, st>>>(W, H, v.bg, g.rec, point_list, img.ranges, img.final_T, img.n_contrib, dL_dcolor, acc);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace s360

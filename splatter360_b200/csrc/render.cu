// render.cu -- per-tile compositing kernels.
//   K6 render_forward_kernel : front-to-back alpha compositing            
//   K7 render_backward_kernel: back-to-front gradient pass                 
// One CTA = one 16x16 tile = 8 warps; a warp owns an 8x4 pixel block.  Each batch of 256 sorted
// instances is staged into shared memory (3 x float4 per instance); every warp first tests, one
// instance per lane, whether the instance's alpha>=1/255 box can reach its 8x4 block (ballot) and
// then evaluates only the survivors, broadcasting their records from shared memory.  The backward
// pass reduces the nine per-instance partials across the warp with a transposed butterfly
// (14 shuffles), accumulates them per tile in shared memory and issues one global atomic per value
// and instance.  Replaces upstream renderCUDA (forward.cu / backward.cu) behind
// /root/reference/src/model/decoder/cuda_splatting.py:113-124.
#include "common.cuh"

namespace s360 {

constexpr int WARP_W = 8, WARP_H = 4;  // pixel block of one warp

template <int MODE>
__device__ __forceinline__ float wrap_dx(float dx, float W, float halfW) {
  if (MODE == S360_MODE_ERP) {
    if (dx > halfW) dx -= W;
    else if (dx < -halfW) dx += W;
  }
  return dx;
}

template <int MODE>
__global__ void __launch_bounds__(TILE_PIX)
render_forward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                      const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                      float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color) {
  __shared__ float4 s_r0[TILE_PIX], s_r1[TILE_PIX], s_r2[TILE_PIX];
  const int gx = (W + TILE - 1) / TILE;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wcx = (float)wx0 + 0.5f * (WARP_W - 1), wcy = (float)wy0 + 0.5f * (WARP_H - 1);
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];

  float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last = 0;
  bool done = !inside;

  for (uint32_t base = range.x; base < range.y; base += TILE_PIX) {
    if (__syncthreads_and(done)) break;
    const uint32_t i = base + tid;
    if (i < range.y) {
      const size_t gid = point_list[i];
      s_r0[tid] = __ldg(rec + 3 * gid);
      s_r1[tid] = __ldg(rec + 3 * gid + 1);
      s_r2[tid] = __ldg(rec + 3 * gid + 2);
    }
    __syncthreads();
    const int cnt = min((uint32_t)TILE_PIX, range.y - base);
    if (__all_sync(0xffffffffu, done)) continue;
    for (int c0 = 0; c0 < cnt; c0 += 32) {
      const int j = c0 + lane;
      bool hit = false;
      if (j < cnt) {
        const float4 a = s_r0[j];
        const float4 b = s_r1[j];
        const float ddx = wrap_dx<MODE>(a.x - wcx, Wf, halfW), ddy = a.y - wcy;
        hit = !(fabsf(ddx) - 0.5f * (WARP_W - 1) > b.z) && !(fabsf(ddy) - 0.5f * (WARP_H - 1) > b.w);
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const int jj = c0 + k;
        const float4 a = s_r0[jj];
        const float4 b = s_r1[jj];
        const float dx = wrap_dx<MODE>(a.x - pxf, Wf, halfW), dy = a.y - pyf;
        const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
        const float alpha = fminf(ALPHA_MAX, b.y * __expf(power));
        bool ok = !done && (power <= 0.f) && (alpha >= ALPHA_MIN);
        const float test_T = T * (1.f - alpha);
        if (ok && test_T < T_EPS) { done = true; ok = false; }
        if (ok) {
          const float4 c = s_r2[jj];
          const float w = alpha * T;
          C0 += c.x * w; C1 += c.y * w; C2 += c.z * w;
          T = test_T;
          last = base - range.x + (uint32_t)jj + 1u;
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
  // xor 16: lower half keeps 0..3, upper half keeps 4..7
  float w0 = (h16 ? v4 : v0) + __shfl_xor_sync(F, h16 ? v0 : v4, 16);
  float w1 = (h16 ? v5 : v1) + __shfl_xor_sync(F, h16 ? v1 : v5, 16);
  float w2 = (h16 ? v6 : v2) + __shfl_xor_sync(F, h16 ? v2 : v6, 16);
  float w3 = (h16 ? v7 : v3) + __shfl_xor_sync(F, h16 ? v3 : v7, 16);
  // xor 8: keeps {0,1} or {2,3}
  float u0 = (h8 ? w2 : w0) + __shfl_xor_sync(F, h8 ? w0 : w2, 8);
  float u1 = (h8 ? w3 : w1) + __shfl_xor_sync(F, h8 ? w1 : w3, 8);
  // xor 4: keeps 0 or 1
  float s = (h4 ? u1 : u0) + __shfl_xor_sync(F, h4 ? u0 : u1, 4);
  s += __shfl_xor_sync(F, s, 2);
  s += __shfl_xor_sync(F, s, 1);
  return s;  // value index = 4*bit4 + 2*bit3 + bit2 of the lane id = (lane >> 2) & 7
}

constexpr int NACC = 9;  // dL/d{r,g,b}, dL/d{u,v}, dL/d{conicA,conicB,conicC}, dL/dopacity

template <int MODE>
__global__ void __launch_bounds__(TILE_PIX)
render_backward_kernel(const int W, const int H, const float* __restrict__ bg, const float4* __restrict__ rec,
                       const uint32_t* __restrict__ point_list, const uint2* __restrict__ ranges,
                       const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dcolor, float* __restrict__ acc) {
  __shared__ float4 s_r0[TILE_PIX], s_r1[TILE_PIX], s_r2[TILE_PIX];
  __shared__ uint32_t s_gid[TILE_PIX];
  __shared__ float s_acc[TILE_PIX * NACC];
  __shared__ uint32_t s_max;
  const int gx = (W + TILE - 1) / TILE;
  const int tile = blockIdx.x;
  const int tx = tile % gx, ty = tile / gx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx0 = tx * TILE + (warp & 1) * WARP_W, wy0 = ty * TILE + (warp >> 1) * WARP_H;
  const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
  const bool inside = px < W && py < H;
  const float pxf = (float)px, pyf = (float)py;
  const float wcx = (float)wx0 + 0.5f * (WARP_W - 1), wcy = (float)wy0 + 0.5f * (WARP_H - 1);
  const float Wf = (float)W, halfW = 0.5f * (float)W;
  const uint2 range = ranges[tile];
  const size_t pid = (size_t)py * W + px, plane = (size_t)H * W;

  const float T_final = inside ? final_T[pid] : 0.f;
  const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;
  float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
  if (inside) { dp0 = dL_dcolor[pid]; dp1 = dL_dcolor[plane + pid]; dp2 = dL_dcolor[2 * plane + pid]; }
  const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;

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

  const int rounds = (int)((todo + TILE_PIX - 1) / TILE_PIX);
  for (int bi = rounds - 1; bi >= 0; --bi) {
    __syncthreads();
    const uint32_t pos0 = (uint32_t)bi * TILE_PIX;
    const int cnt = (int)min((uint32_t)TILE_PIX, todo - pos0);
    if (tid < cnt) {
      const uint32_t gid = point_list[range.x + pos0 + tid];
      s_gid[tid] = gid;
      s_r0[tid] = __ldg(rec + 3 * (size_t)gid);
      s_r1[tid] = __ldg(rec + 3 * (size_t)gid + 1);
      s_r2[tid] = __ldg(rec + 3 * (size_t)gid + 2);
    }
#pragma unroll
    for (int k = 0; k < NACC; k++) s_acc[k * TILE_PIX + tid] = 0.f;  // linear zero fill
    __syncthreads();
    for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
      if (pos0 + (uint32_t)c0 >= warp_max) continue;
      const int j = c0 + lane;
      bool hit = false;
      if (j < cnt && pos0 + (uint32_t)j < warp_max) {
        const float4 a = s_r0[j];
        const float4 b = s_r1[j];
        const float ddx = wrap_dx<MODE>(a.x - wcx, Wf, halfW), ddy = a.y - wcy;
        hit = !(fabsf(ddx) - 0.5f * (WARP_W - 1) > b.z) && !(fabsf(ddy) - 0.5f * (WARP_H - 1) > b.w);
      }
      unsigned mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int k = 31 - __clz(mask);
        mask &= ~(1u << k);
        const int jj = c0 + k;
        const float4 a = s_r0[jj];
        const float4 b = s_r1[jj];
        const float dx = wrap_dx<MODE>(a.x - pxf, Wf, halfW), dy = a.y - pyf;
        const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
        const float G = __expf(power);
        const float alpha = fminf(ALPHA_MAX, b.y * G);
        const bool ok = (pos0 + (uint32_t)jj < last_contributor) && (power <= 0.f) && (alpha >= ALPHA_MIN);
        if (!__any_sync(0xffffffffu, ok)) continue;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
        if (ok) {
          const float4 c = s_r2[jj];
          T = T / (1.f - alpha);
          const float w = alpha * T;
          ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0;
          ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1;
          ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2;
          lc0 = c.x; lc1 = c.y; lc2 = c.z;
          float dL_dalpha = (c.x - ar0) * dp0 + (c.y - ar1) * dp1 + (c.z - ar2) * dp2;
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
          const float dL_dG = b.y * dL_dalpha;
          const float gdx = G * dx, gdy = G * dy;
          v0 = w * dp0; v1 = w * dp1; v2 = w * dp2;
          v3 = dL_dG * (-gdx * a.z - gdy * a.w);
          v4 = dL_dG * (-gdy * b.x - gdx * a.w);
          v5 = -0.5f * gdx * dx * dL_dG;
          v6 = -gdx * dy * dL_dG;
          v7 = -0.5f * gdy * dy * dL_dG;
          v8 = G * dL_dalpha;
        }
        const float s8 = warp_reduce8(v0, v1, v2, v3, v4, v5, v6, v7, lane);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 16);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 8);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 4);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 2);
        v8 += __shfl_xor_sync(0xffffffffu, v8, 1);
        if ((lane & 3) == 0) atomicAdd(&s_acc[jj * NACC + ((lane >> 2) & 7)], s8);
        if (lane == 1) atomicAdd(&s_acc[jj * NACC + 8], v8);
      }
    }
    __syncthreads();
    if (tid < cnt) {
      float* dst = acc + (size_t)s_gid[tid] * ACC_STRIDE;
      bool any = false;
      float vals[NACC];
#pragma unroll
      for (int k = 0; k < NACC; k++) { vals[k] = s_acc[tid * NACC + k]; any = any || (vals[k] != 0.f); }
      if (any) {
#pragma unroll
        for (int k = 0; k < NACC; k++) atomicAdd(dst + k, vals[k]);
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

// preprocess.cu -- per-Gaussian kernels.
//   K1  preprocess_kernel          : world->view, pinhole / equirectangular projection, EWA cov2D,
//                                    conic, extent, tile rect, SH -> RGB        
//   K8+K9 preprocess_backward_kernel: dL/d(conic, mean2D, rgb, opacity) -> dL/d(cov3D, mean3D, SH)
//   K10 mark_visible_kernel
// Replaces upstream preprocessCUDA / computeCov2DCUDA / preprocessCUDA(bwd) / checkFrustum behind
// /root/reference/src/model/decoder/cuda_splatting.py:113-124.
#include "common.cuh"

namespace s360 {

constexpr int PRE_THREADS = 128;

// covariance in: [P,6] (xx,xy,xz,yy,yz,zz) or the reference's [P,3,3] (upper triangle is read,
// cuda_splatting.py:115,123), multiplied by scene_scale^2
__device__ __forceinline__ void load_cov6(const S360View& v, const float* __restrict__ cov, int idx, float* cv) {
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
__device__ __forceinline__ void store_dcov(const S360View& v, float* __restrict__ d_cov, int idx, const float* g, float s2) {
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
__device__ __forceinline__ void project_view(const S360View& v, const float* V, const float* PM, float mx, float my,
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
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(o.sortkey));
    const float jb = (fx * fx * (1.f + limx * limx) + fy * fy * (1.f + limy * limy)) * iz * iz;
    float rb;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(jb * wf * fmaxf(cv[0] + cv[3] + cv[5], 0.f) + 2.f * fabsf(v.lowpass) + 0.32f));
    rb = 3.03f * rb + 2.f;
    const float qx = PM[0] * mx + PM[4] * my + PM[8] * mz + PM[12];
    const float qy = PM[1] * mx + PM[5] * my + PM[9] * mz + PM[13];
    const float qw = PM[3] * mx + PM[7] * my + PM[11] * mz + PM[15];
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(qw + 0.0000001f));
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
      o.px = su * atan2f(g.t[0], g.t[2]) + 0.5f * W - 0.5f;
      o.py = sv * atan2f(g.t[1], sqrtf(g.t[0] * g.t[0] + g.t[2] * g.t[2])) + 0.5f * H - 0.5f;
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
      o.hx = __int_as_float(0x7f800000); o.hy = o.hx;
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
        o.key = __float_as_uint(o.sortkey);
        o.tiles = (uint32_t)(nx * ny);
      }
      o.cl = (g.clampx ? 8 : 0) | (g.clampy ? 16 : 0);
    }
  }
}

// squared Frobenius norm of the rotation block of a view matrix (3 for a rigid camera); scales the reject bound of
// project_view so that it stays conservative for any matrix a caller hands in
__device__ __forceinline__ float view_frobenius2(const float* V) {
  return V[0] * V[0] + V[1] * V[1] + V[2] * V[2] + V[4] * V[4] + V[5] * V[5] + V[6] * V[6] + V[8] * V[8] + V[9] * V[9] +
         V[10] * V[10];
}

// SH -> RGB for one Gaussian (row `sh` of the staged block) seen from `campos`; sets the clamp bits 0..2 of cl
__device__ __forceinline__ void sh_to_rgb(const S360View& v, const float* sh, float mx, float my, float mz,
                                          const float* campos, float* col, uint8_t& cl) {
  float dx = mx - campos[0], dy = my - campos[1], dz = mz - campos[2];
  const float inv = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
  dx *= inv; dy *= inv; dz *= inv;
  float b[25];
  const int deg = min(v.sh_degree, v.max_sh_degree);
  const int n = sh_basis(deg, dx, dy, dz, b);
  const int ks = v.sh_layout ? 1 : 3, cs = v.sh_layout ? v.M : 1;   // [P,M,3] or the reference's [P,3,M]
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 25; k++) {
    if (k < n) {
      acc[0] += b[k] * sh[ks * k];
      acc[1] += b[k] * sh[ks * k + cs];
      acc[2] += b[k] * sh[ks * k + 2 * cs];
    }
  }
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const float r = acc[ch] + 0.5f;
    if (r < 0.f) cl |= (1 << ch);
    col[ch] = fmaxf(r, 0.f);
  }
}

template <int MODE>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_kernel(const S360View v, const float* __restrict__ means, const float* __restrict__ cov3D,
                  const float* __restrict__ opac, const float* __restrict__ shs,
                  const float* __restrict__ colors, GeomState gs, int32_t* __restrict__ radii,
                  uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ ids, S360Counters* counters) {
  extern __shared__ __align__(128) float s_sh[];   // [PRE_THREADS][M*3] SH block of this CTA
  __shared__ uint64_t s_bar;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = v.P;
  const int row = v.M * 3;                                   // floats per Gaussian
  const int rows = min(PRE_THREADS, P - blockIdx.x * PRE_THREADS);
  const float* sh_src = shs ? shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  const uint32_t sh_bytes = (uint32_t)rows * row * 4u;
  const bool bulk_ok = shs && (sh_bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(sh_src) & 15u) == 0u);
  if (shs) {
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    // erp: every Gaussian is a candidate, start the copy before the geometry math (overlap)
    if (MODE == S360_MODE_ERP && bulk_ok && threadIdx.x == 0) {
      mbar_expect_tx(&s_bar, sh_bytes);
      bulk_load(s_sh, sh_src, sh_bytes, &s_bar);
    }
  }
  bool upstream_visible = false;
  bool want_color = false;
  uint32_t my_tiles = 0;
  float px = 0.f, py = 0.f, cA = 0.f, cB = 0.f, cC = 0.f, op = 0.f, hx = 0.f, hy = 0.f, sortkey = 0.f;
  float mx = 0.f, my = 0.f, mz = 0.f;
  uint8_t cl = 0;
  Cam cam;
  if (idx < P) {
    load_cam(v, cam, MODE == S360_MODE_PINHOLE);
    const float sc = v.scene_scale;   // reference's 1/near rescale (cuda_splatting.py:64-71), folded into the load
    mx = means[3 * idx] * sc; my = means[3 * idx + 1] * sc; mz = means[3 * idx + 2] * sc;
    float cv[6];
    load_cov6(v, cov3D, idx, cv);
    Proj pr;
    project_view<MODE>(v, cam.V, cam.PM, mx, my, mz, cv, opac, idx, view_frobenius2(cam.V), pr);
    upstream_visible = pr.upstream_visible;
    my_tiles = pr.tiles;
    px = pr.px; py = pr.py; cA = pr.cA; cB = pr.cB; cC = pr.cC; op = pr.op; hx = pr.hx; hy = pr.hy; sortkey = pr.sortkey;
    cl = pr.cl;
    want_color = upstream_visible;
    const int radius = pr.radius;
    const uint2 rect = pr.rect;
    const uint32_t key = pr.key;
    radii[idx] = radius;
    gs.rect[idx] = rect;
    depth_keys[idx] = key;
    ids[idx] = (uint32_t)idx;
  }
  // block totals -> one atomic each.  The instance total is known here already (the scan only orders it), which
  // lets the host size the instance buffers while the depth sort is still running.
  {
    __shared__ uint32_t s_cnt[2];
    if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    __syncthreads();
    const unsigned m = __ballot_sync(0xffffffffu, upstream_visible);
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, my_tiles);
    if ((threadIdx.x & 31) == 0) {
      if (m) atomicAdd(&s_cnt[0], (uint32_t)__popc(m));
      if (wsum) atomicAdd(&s_cnt[1], wsum);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (s_cnt[0]) atomicAdd(&counters->num_visible, s_cnt[0]);
      if (s_cnt[1]) atomicAdd(&counters->num_rendered, s_cnt[1]);
    }
  }

  // ---- colour: SH block staged in shared memory (TMA bulk copy; coalesced fallback for odd tails)
  float col[3] = {0.f, 0.f, 0.f};
  if (shs != nullptr) {
    bool need = true;
    if (MODE == S360_MODE_PINHOLE || !bulk_ok) need = __syncthreads_or(want_color);   // most pinhole blocks are culled
    if (need) {
      if (bulk_ok) {
        if (MODE == S360_MODE_PINHOLE && threadIdx.x == 0) {
          mbar_expect_tx(&s_bar, sh_bytes);
          bulk_load(s_sh, sh_src, sh_bytes, &s_bar);
        }
        mbar_wait(&s_bar, 0);
      } else {
        for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) s_sh[i] = sh_src[i];
        __syncthreads();
      }
      if (want_color) sh_to_rgb(v, s_sh + threadIdx.x * row, mx, my, mz, cam.cam, col, cl);
    }
  } else if (want_color) {
    col[0] = colors[3 * idx]; col[1] = colors[3 * idx + 1]; col[2] = colors[3 * idx + 2];
  }
  if (want_color) {
    gs.rec[3 * (size_t)idx + 0] = make_float4(px, py, cA, cB);
    gs.rec[3 * (size_t)idx + 1] = make_float4(cC, op, hx, hy);
    gs.rec[3 * (size_t)idx + 2] = make_float4(col[0], col[1], col[2], sortkey);
    gs.clamped[idx] = cl;
  }
}

int launch_preprocess(const S360View& v, const float* means, const float* cov, const float* opac,
                      const float* shs, const float* colors, GeomState g, int32_t* radii,
                      uint32_t* depth_keys, uint32_t* ids, S360Counters* counters, cudaStream_t st) {
  if (v.P == 0) return 0;
  const int grid = (v.P + PRE_THREADS - 1) / PRE_THREADS;
  const size_t smem = shs ? (size_t)PRE_THREADS * v.M * 3 * sizeof(float) : 0;
  if (smem > 200 * 1024) return S360_ERR_UNSUPPORTED;
  if (v.mode == S360_MODE_PINHOLE) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(preprocess_kernel<S360_MODE_PINHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    preprocess_kernel<S360_MODE_PINHOLE><<<grid, PRE_THREADS, smem, st>>>(v, means, cov, opac, shs, colors, g, radii, depth_keys, ids, counters);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(preprocess_kernel<S360_MODE_ERP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    preprocess_kernel<S360_MODE_ERP><<<grid, PRE_THREADS, smem, st>>>(v, means, cov, opac, shs, colors, g, radii, depth_keys, ids, counters);
  }
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Per-view geometry backward (K8): the moments the render pass accumulated for one (view, Gaussian) pair ->
// dL/d(mean3D) through the projection and the Jacobian, dL/d(cov3D), screen-space gradient dm2 (NDC units).
// a0 = {dL/dr, dL/dg, dL/db, sum q dx}, a1 = {sum q dy, sum q dx^2, sum q dxdy, sum q dy^2}, q = G dL/dalpha.
// Shared by the single-view and the batched kernel; the SH part (direction gradient) is added by the caller.
template <int MODE, bool DEPTH = false>
__device__ __forceinline__ void view_backward(const S360View& v, const float* V, const float* PM, float mx, float my,
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
  dcov[0] = Mm[0][0] * Mm[0][0] * da + Mm[0][0] * Mm[1][0] * db + Mm[1][0] * Mm[1][0] * dc;
  dcov[3] = Mm[0][1] * Mm[0][1] * da + Mm[0][1] * Mm[1][1] * db + Mm[1][1] * Mm[1][1] * dc;
  dcov[5] = Mm[0][2] * Mm[0][2] * da + Mm[0][2] * Mm[1][2] * db + Mm[1][2] * Mm[1][2] * dc;
  dcov[1] = 2.f * Mm[0][0] * Mm[0][1] * da + (Mm[0][0] * Mm[1][1] + Mm[0][1] * Mm[1][0]) * db + 2.f * Mm[1][0] * Mm[1][1] * dc;
  dcov[2] = 2.f * Mm[0][0] * Mm[0][2] * da + (Mm[0][0] * Mm[1][2] + Mm[0][2] * Mm[1][0]) * db + 2.f * Mm[1][0] * Mm[1][2] * dc;
  dcov[4] = 2.f * Mm[0][2] * Mm[0][1] * da + (Mm[0][1] * Mm[1][2] + Mm[0][2] * Mm[1][1]) * db + 2.f * Mm[1][1] * Mm[1][2] * dc;
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
  dm[0] = dm[1] = dm[2] = 0.f;
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
    dm[0] = (PM[0] * mw - PM[3] * mul1) * dm2[0] + (PM[1] * mw - PM[3] * mul2) * dm2[1];
    dm[1] = (PM[4] * mw - PM[7] * mul1) * dm2[0] + (PM[5] * mw - PM[7] * mul2) * dm2[1];
    dm[2] = (PM[8] * mw - PM[11] * mul1) * dm2[0] + (PM[9] * mw - PM[11] * mul2) * dm2[1];
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

// ------------------------------------------------------------------------------------------------
// K8 + K9 fused.  acc[idx*12 + 0..8] = {dL/dr, dL/dg, dL/db, and the moments sum(q dx), sum(q dy), sum(q dx^2),
// sum(q dx dy), sum(q dy^2), sum(q)} with q = G dL/dalpha, accumulated by render_backward_kernel
template <int MODE, bool DEPTH>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_backward_kernel(const S360View v, const float* __restrict__ means, const float* __restrict__ cov3D,
                           const float* __restrict__ opac, const float* __restrict__ shs, GeomState gs, const int32_t* __restrict__ radii,
                           const float* __restrict__ acc, float* __restrict__ d_means,
                           float* __restrict__ d_means2D, float* __restrict__ d_cov, float* __restrict__ d_opac,
                           float* __restrict__ d_shs, float* __restrict__ d_colors, const DepthSpec dspec) {
  extern __shared__ __align__(128) float s_sh[];   // [PRE_THREADS][M*3]: SH in, dL/dSH out (in place)
  __shared__ uint64_t s_bar;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = v.P;
  const int row = v.M * 3;
  const int rows = min(PRE_THREADS, P - blockIdx.x * PRE_THREADS);
  const float* sh_src = shs ? shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  float* dsh_dst = shs ? d_shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  const uint32_t sh_bytes = (uint32_t)rows * row * 4u;
  const bool bulk_ok = shs && (sh_bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(sh_src) & 15u) == 0u) &&
                       ((reinterpret_cast<uintptr_t>(dsh_dst) & 15u) == 0u);
  const bool vis = idx < P && radii[idx] > 0;
  bool need = false;
  if (shs) {
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    need = __syncthreads_or(vis);
    if (need) {
      if (bulk_ok) {
        if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, sh_bytes); bulk_load(s_sh, sh_src, sh_bytes, &s_bar); }
      } else {
        for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) s_sh[i] = sh_src[i];
        __syncthreads();
      }
    } else {
      // nothing visible in this block: the SH gradient block is all zeros
      for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) dsh_dst[i] = 0.f;
    }
  }
  float dm[3] = {0.f, 0.f, 0.f}, dm2[2] = {0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dop = 0.f, dcol[3] = {0.f, 0.f, 0.f};
  if (vis) {
    Cam cam;
    load_cam(v, cam, MODE == S360_MODE_PINHOLE);
    const float* V = cam.V;
    const float4 a0 = *reinterpret_cast<const float4*>(acc + (size_t)idx * ACC_STRIDE);
    const float4 a1 = *reinterpret_cast<const float4*>(acc + (size_t)idx * ACC_STRIDE + 4);
    const float a2 = acc[(size_t)idx * ACC_STRIDE + 8];
    dcol[0] = a0.x; dcol[1] = a0.y; dcol[2] = a0.z;
#if S360_BWD_QPRIME
    dop = opac[idx] > 0.f ? a2 / opac[idx] : 0.f;
#else
    dop = a2;
#endif
    const float sc = v.scene_scale;
    const float mx = means[3 * idx] * sc, my = means[3 * idx + 1] * sc, mz = means[3 * idx + 2] * sc;
    float cv[6];
    load_cov6(v, cov3D, idx, cv);
    const float op = opac[idx];
    view_backward<MODE, DEPTH>(v, V, cam.PM, mx, my, mz, cv, op, a0, a1, dm, dm2, dcov, dspec,
                               DEPTH ? acc[(size_t)idx * ACC_STRIDE + 9] : 0.f);
    if (shs != nullptr) {
      const float ox = mx - cam.cam[0], oy = my - cam.cam[1], oz = mz - cam.cam[2];
      const float inv = 1.f / sqrtf(ox * ox + oy * oy + oz * oz);
      const float dx = ox * inv, dy = oy * inv, dz = oz * inv;
      const int deg = min(v.sh_degree, v.max_sh_degree);
      float b[25], bx[25], by[25], bz[25];
      const int n = sh_basis(deg, dx, dy, dz, b);
      sh_basis_grad(deg, dx, dy, dz, bx, by, bz);
      const uint8_t cl = gs.clamped[idx];
      const float drgb[3] = {(cl & 1) ? 0.f : dcol[0], (cl & 2) ? 0.f : dcol[1], (cl & 4) ? 0.f : dcol[2]};
      if (bulk_ok) mbar_wait(&s_bar, 0);
      float* sh = s_sh + threadIdx.x * row;      // this thread's row: read SH, overwrite with dL/dSH
      const int ks = v.sh_layout ? 1 : 3, cs = v.sh_layout ? v.M : 1;
      float ddir[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 25; k++) {
        if (k < n) {
          const float s = sh[ks * k] * drgb[0] + sh[ks * k + cs] * drgb[1] + sh[ks * k + 2 * cs] * drgb[2];
          ddir[0] += bx[k] * s; ddir[1] += by[k] * s; ddir[2] += bz[k] * s;
          sh[ks * k] = b[k] * drgb[0]; sh[ks * k + cs] = b[k] * drgb[1]; sh[ks * k + 2 * cs] = b[k] * drgb[2];
        }
      }
      for (int k = n; k < v.M; k++) { sh[ks * k] = 0.f; sh[ks * k + cs] = 0.f; sh[ks * k + 2 * cs] = 0.f; }
      const float dot = dx * ddir[0] + dy * ddir[1] + dz * ddir[2];
      dm[0] += (ddir[0] - dx * dot) * inv;
      dm[1] += (ddir[1] - dy * dot) * inv;
      dm[2] += (ddir[2] - dz * dot) * inv;
    }
  } else if (shs != nullptr && need && idx < P) {
    if (bulk_ok) mbar_wait(&s_bar, 0);
    float* sh = s_sh + threadIdx.x * row;
    for (int k = 0; k < row; k++) sh[k] = 0.f;
  }
  if (shs != nullptr && need) {
    if (bulk_ok) {
      if (idx >= P) mbar_wait(&s_bar, 0);
      fence_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) { bulk_store(dsh_dst, s_sh, sh_bytes); bulk_store_wait_read(); }
    } else {
      __syncthreads();
      for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) dsh_dst[i] = s_sh[i];
    }
  }
  if (idx >= P) return;
  // gradients w.r.t. the caller's (unscaled) means / covariances
  const float gsc = v.scene_scale;
#pragma unroll
  for (int k = 0; k < 3; k++) d_means[3 * idx + k] = dm[k] * gsc;
  d_means2D[3 * idx] = dm2[0]; d_means2D[3 * idx + 1] = dm2[1]; d_means2D[3 * idx + 2] = 0.f;
  store_dcov(v, d_cov, idx, dcov, gsc * gsc);
  d_opac[idx] = dop;
  if (d_colors != nullptr) {
    const bool pre = shs == nullptr;
#pragma unroll
    for (int k = 0; k < 3; k++) d_colors[3 * idx + k] = pre ? dcol[k] : 0.f;
  }
}

int launch_preprocess_backward(const S360View& v, const float* means, const float* cov, const float* opac, const float* shs,
                               GeomState g, const int32_t* radii, const float* acc, float* d_means,
                               float* d_means2D, float* d_cov, float* d_opac, float* d_shs, float* d_colors,
                               int has_depth, int depth_mode, float depth_near, float depth_far, cudaStream_t st) {
  if (v.P == 0) return 0;
  const int grid = (v.P + PRE_THREADS - 1) / PRE_THREADS;
  const size_t smem = shs ? (size_t)PRE_THREADS * v.M * 3 * sizeof(float) : 0;
  if (smem > 200 * 1024) return S360_ERR_UNSUPPORTED;
  DepthSpec ds;
  ds.mode = depth_mode; ds.inv_scale = 1.f / v.scene_scale; ds.near = depth_near; ds.far = depth_far;
#define S360_LAUNCH_K8(MODE_, DEPTH_) do { \
    if (smem > 48 * 1024) cudaFuncSetAttribute(preprocess_backward_kernel<MODE_, DEPTH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    preprocess_backward_kernel<MODE_, DEPTH_><<<grid, PRE_THREADS, smem, st>>>(v, means, cov, opac, shs, g, radii, acc, d_means, d_means2D, d_cov, d_opac, d_shs, d_colors, ds); } while (0)
  if (v.mode == S360_MODE_PINHOLE) { if (has_depth) S360_LAUNCH_K8(S360_MODE_PINHOLE, true); else S360_LAUNCH_K8(S360_MODE_PINHOLE, false); }
  else { if (has_depth) S360_LAUNCH_K8(S360_MODE_ERP, true); else S360_LAUNCH_K8(S360_MODE_ERP, false); }
#undef S360_LAUNCH_K8
  count_launch();
  return (int)cudaGetLastError();
}

// ================================================================================================
// Batched multi-view path: V views of the same Gaussians in ONE pass.  The reference
// renders a panorama as six pinhole cube faces, i.e. six full rasterizer calls that each re-read all P Gaussians
// (/root/reference/src/model/decoder/decoder_splatting_cuda.py:44-59, model_wrapper_erp.py:336-345).  Here every
// Gaussian is read once, projected into all V views, and each (view, Gaussian) PAIR that touches at least one tile
// gets a slot in compacted pair buffers -- in (Gaussian, view) order (count kernel, scan of the per-CTA counts, write
// kernel: an in-kernel look-back was measured 45 % slower, its CTAs convoy behind the slowest predecessor), so that
// equal depths keep index order exactly like V separate calls.  Everything
// downstream (depth sort, scan, emission, tile sort, compositing, render backward) then runs once over the pairs on
// a virtual image of V stacked views; per-pair moments are folded back per Gaussian in the batched K8+K9.
constexpr int CAM_F = 36;   // floats per staged camera: V[16], PM[16], campos[3], pad

__device__ __forceinline__ void stage_cameras(const S360View& v, int NV, bool need_proj, float (*s_cam)[CAM_F]) {
  for (int i = threadIdx.x; i < NV * CAM_F; i += blockDim.x) {
    const int view = i / CAM_F, j = i - view * CAM_F;
    float x = 0.f;
    if (j < 16) x = __ldg(v.viewmatrix + 16 * view + j);
    else if (j < 32) x = need_proj ? __ldg(v.projmatrix + 16 * view + (j - 16)) : 0.f;
    else if (j < 35) x = __ldg(v.campos + 3 * view + (j - 32));
    if (j < 35) s_cam[view][j] = x;
  }
  // slot 35: squared Frobenius norm of the rotation block (bound of project_view's frustum reject)
  for (int view = threadIdx.x; view < NV; view += blockDim.x) {
    const float* V = v.viewmatrix + 16 * view;
    float f = 0.f;
    for (int c = 0; c < 3; c++)
      for (int r = 0; r < 3; r++) { const float x = __ldg(V + 4 * c + r); f += x * x; }
    s_cam[view][35] = f;
  }
}

#ifndef S360_MV_FWD_MINB
#define S360_MV_FWD_MINB 5
#endif
#ifndef S360_MV_BWD_MINB
#define S360_MV_BWD_MINB 4
#endif

// K1a: which views does each Gaussian reach?  One thread per Gaussian projects it into all NV views (near cull and a
// conservative frustum test reject most of them before any covariance math), stores the view mask, and the CTA
// stores its pair count; the instance total is accumulated here (final before the sort, as in the single-view K1).
template <int MODE>
__global__ void __launch_bounds__(PRE_THREADS)
multi_count_kernel(const S360View v, const int NV, const float* __restrict__ means, const float* __restrict__ cov3D,
                   const float* __restrict__ opac, PairState ps, int32_t* __restrict__ radii,
                   uint32_t* __restrict__ block_count, S360Counters* counters) {
  __shared__ float s_cam[S360_MAX_VIEWS][CAM_F];
  __shared__ uint32_t s_cnt[2];
  if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
  stage_cameras(v, NV, MODE == S360_MODE_PINHOLE, s_cam);
  __syncthreads();
  const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
  const int P = v.P;
  uint32_t mask = 0, tiles = 0;
  if (idx < P) {
    const float sc = v.scene_scale;
    const float mx = means[3 * idx] * sc, my = means[3 * idx + 1] * sc, mz = means[3 * idx + 2] * sc;
    float cv[6];
    load_cov6(v, cov3D, idx, cv);
    for (int view = 0; view < NV; view++) {
      Proj pr;
      project_view<MODE>(v, s_cam[view], s_cam[view] + 16, mx, my, mz, cv, opac, idx, s_cam[view][35], pr);
      if (pr.tiles) { mask |= 1u << view; tiles += pr.tiles; }
      if (radii) radii[(size_t)view * P + idx] = pr.radius;
    }
    ps.mask[idx] = mask;
  }
  const uint32_t wc = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mask));
  const uint32_t wt = __reduce_add_sync(0xffffffffu, tiles);
  if ((threadIdx.x & 31) == 0) {
    if (wc) atomicAdd(&s_cnt[0], wc);
    if (wt) atomicAdd(&s_cnt[1], wt);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    block_count[blockIdx.x] = s_cnt[0];
    if (s_cnt[1]) atomicAdd(&counters->num_rendered, s_cnt[1]);
  }
}

// K1b: exclusive scan of the per-CTA pair counts (one CTA; in place) -> first pair slot of every K1c CTA, the pair
// total, and the overflow flag.  Pairs are therefore numbered in (Gaussian, view) order: equal depths keep index
// order through the stable sorts exactly like separate per-view calls.
__global__ void __launch_bounds__(1024)
multi_scan_kernel(int nblocks, uint32_t* __restrict__ block_count, uint32_t cap, PairState ps, S360Counters* counters) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t x = i < nblocks ? block_count[i] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += s_w[w];
    const uint32_t carry = s_carry;
    if (i < nblocks) block_count[i] = carry + woff + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t total = s_carry;
    counters->num_visible = total;              // pairs this batch needs
    *ps.count = min(total, cap);                // pairs stored
    if (total > cap) atomicOr(&counters->overflow, 2u);
  }
}

// K1c: write the pair records.  Same CTA partition as K1a; a thread's first slot is its CTA's base plus the
// in-CTA prefix of the pair counts; only the views in the mask are projected again.
template <int MODE>
__global__ void __launch_bounds__(PRE_THREADS, S360_MV_FWD_MINB)
multi_write_kernel(const S360View v, const int NV, const uint32_t cap, const float* __restrict__ means,
                   const float* __restrict__ cov3D, const float* __restrict__ opac, const float* __restrict__ shs,
                   const float* __restrict__ colors, GeomState gs, PairState ps,
                   const uint32_t* __restrict__ block_base, uint32_t* __restrict__ depth_keys,
                   uint32_t* __restrict__ ids) {
  extern __shared__ __align__(128) float s_sh[];   // [PRE_THREADS][M*3] SH block of this CTA
  __shared__ uint64_t s_bar;
  __shared__ float s_cam[S360_MAX_VIEWS][CAM_F];
  __shared__ uint32_t s_w[PRE_THREADS / 32];
  const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
  const int P = v.P;
  const int gy = (v.image_height + TILE - 1) / TILE;
  const int row = v.M * 3;
  const int rows = min(PRE_THREADS, P - blockIdx.x * PRE_THREADS);
  const float* sh_src = shs ? shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  const uint32_t sh_bytes = (uint32_t)rows * row * 4u;
  const bool bulk_ok = shs && (sh_bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(sh_src) & 15u) == 0u);
  const uint32_t mask = idx < P ? ps.mask[idx] : 0u;
  stage_cameras(v, NV, MODE == S360_MODE_PINHOLE, s_cam);
  if (shs && threadIdx.x == 0) mbar_init(&s_bar, 1);
  const bool need = __syncthreads_or(mask != 0u) != 0;   // also publishes the cameras and the barrier
  if (!need) {                                            // no pair in this CTA: nothing to write, SH never read
    if (idx < P) ps.base[idx] = 0u;                       // (the backward pass loads base next to mask, unconditionally)
    return;
  }
  if (shs) {
    if (bulk_ok) {
      if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, sh_bytes); bulk_load(s_sh, sh_src, sh_bytes, &s_bar); }
    } else {
      for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) s_sh[i] = sh_src[i];   // visible after the scan barrier
    }
  }
  // in-CTA exclusive prefix of the pair counts
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t cnt = (uint32_t)__popc(mask);
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_w[warp] = incl;
  float mx = 0.f, my = 0.f, mz = 0.f;
  float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (mask) {
    const float sc = v.scene_scale;
    mx = means[3 * idx] * sc; my = means[3 * idx + 1] * sc; mz = means[3 * idx + 2] * sc;
    load_cov6(v, cov3D, idx, cv);
  }
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < PRE_THREADS / 32; w++) woff += (w < warp) ? s_w[w] : 0u;
  uint32_t slot = block_base[blockIdx.x] + woff + incl - cnt;
  // every thread waits for the bulk copy: no thread may leave while the TMA still writes this CTA's shared memory
  if (shs && bulk_ok) mbar_wait(&s_bar, 0);
  if (idx >= P) return;
  ps.base[idx] = slot;
  uint32_t kept = mask;
  float col[3] = {0.f, 0.f, 0.f};
  uint8_t clc = 0;
  int col_view = -1;
  for (uint32_t m = mask; m; m &= m - 1, slot++) {
    const int view = __ffs(m) - 1;
    if (slot >= cap) { kept &= ~m; break; }   // pair buffers full: drop this and the remaining views (flagged by K1b)
    const float* cam = s_cam[view];
    Proj pr;
    project_view<MODE>(v, cam, cam + 16, mx, my, mz, cv, opac, idx, cam[35], pr);
    if (shs != nullptr) {
      // same camera centre as the last evaluated view (cube faces): same direction, same colour
      const bool same = col_view >= 0 && s_cam[col_view][32] == cam[32] && s_cam[col_view][33] == cam[33] &&
                        s_cam[col_view][34] == cam[34];
      if (!same) { clc = 0; sh_to_rgb(v, s_sh + threadIdx.x * row, mx, my, mz, cam + 32, col, clc); col_view = view; }
    } else if (col_view < 0) {
      col[0] = colors[3 * idx]; col[1] = colors[3 * idx + 1]; col[2] = colors[3 * idx + 2];
      col_view = view;
    }
    gs.rec[3 * (size_t)slot + 0] = make_float4(pr.px, pr.py, pr.cA, pr.cB);
    gs.rec[3 * (size_t)slot + 1] = make_float4(pr.cC, pr.op, pr.hx, pr.hy);
    gs.rec[3 * (size_t)slot + 2] = make_float4(col[0], col[1], col[2], pr.sortkey);
    gs.rect[slot] = make_uint2(pr.rect.x, pr.rect.y + (uint32_t)(view * gy));   // tile row on the stacked image
    gs.clamped[slot] = pr.cl | clc;
    depth_keys[slot] = pr.key;
    ids[slot] = slot;
  }
  if (kept != mask) ps.mask[idx] = kept;
}

int launch_preprocess_multi(const S360View& v, int NV, int64_t pair_capacity, const float* means, const float* cov,
                            const float* opac, const float* shs, const float* colors, GeomState g, PairState ps,
                            int32_t* radii, uint32_t* depth_keys, uint32_t* ids, S360Counters* counters,
                            uint32_t* status, cudaStream_t st) {
  if (v.P == 0) return 0;
  const int grid = (v.P + PRE_THREADS - 1) / PRE_THREADS;
  const size_t smem = shs ? (size_t)PRE_THREADS * v.M * 3 * sizeof(float) : 0;
  if (smem > 190 * 1024) return S360_ERR_UNSUPPORTED;
  const uint32_t cap = (uint32_t)(pair_capacity < 0x3fffffff ? pair_capacity : 0x3fffffff);
  uint32_t* block_count = status;   // [grid]: pair count per CTA, scanned in place to the CTA's first slot
  if (v.mode == S360_MODE_PINHOLE)
    multi_count_kernel<S360_MODE_PINHOLE><<<grid, PRE_THREADS, 0, st>>>(v, NV, means, cov, opac, ps, radii, block_count, counters);
  else
    multi_count_kernel<S360_MODE_ERP><<<grid, PRE_THREADS, 0, st>>>(v, NV, means, cov, opac, ps, radii, block_count, counters);
  multi_scan_kernel<<<1, 1024, 0, st>>>(grid, block_count, cap, ps, counters);
  if (v.mode == S360_MODE_PINHOLE) {
    if (smem > 40 * 1024) cudaFuncSetAttribute(multi_write_kernel<S360_MODE_PINHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    multi_write_kernel<S360_MODE_PINHOLE><<<grid, PRE_THREADS, smem, st>>>(v, NV, cap, means, cov, opac, shs, colors, g, ps, block_count, depth_keys, ids);
  } else {
    if (smem > 40 * 1024) cudaFuncSetAttribute(multi_write_kernel<S360_MODE_ERP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    multi_write_kernel<S360_MODE_ERP><<<grid, PRE_THREADS, smem, st>>>(v, NV, cap, means, cov, opac, shs, colors, g, ps, block_count, depth_keys, ids);
  }
  count_launch(3);
  return (int)cudaGetLastError();
}

// zero the first min(*n_dev, cap) pair accumulators (the pair buffers are sized for the worst case V * P)
__global__ void zero_acc_kernel(float4* __restrict__ acc, const uint32_t* __restrict__ n_dev, int64_t cap) {
  int64_t n = (int64_t)(*n_dev);
  if (n > cap) n = cap;
  n *= ACC_STRIDE / 4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc[i] = z;
}
int launch_zero_acc(float* acc, const uint32_t* n_dev, int64_t cap, cudaStream_t st) {
  zero_acc_kernel<<<148 * 8, 256, 0, st>>>((float4*)acc, n_dev, cap);
  count_launch();
  return (int)cudaGetLastError();
}

// direction part of the SH backward for one view: dL/d(mean) through the normalised view direction
__device__ __forceinline__ void sh_dir_backward(const S360View& v, const float* sh, float mx, float my, float mz,
                                                const float* campos, const float* drgb, float* dm) {
  const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
  const float inv = 1.f / sqrtf(ox * ox + oy * oy + oz * oz);
  const float dx = ox * inv, dy = oy * inv, dz = oz * inv;
  const int deg = min(v.sh_degree, v.max_sh_degree);
  float bx[25], by[25], bz[25];
  const int n = (deg + 1) * (deg + 1);
  sh_basis_grad(deg, dx, dy, dz, bx, by, bz);
  const int ks = v.sh_layout ? 1 : 3, cs = v.sh_layout ? v.M : 1;
  float ddir[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 25; k++) {
    if (k < n) {
      const float s = sh[ks * k] * drgb[0] + sh[ks * k + cs] * drgb[1] + sh[ks * k + 2 * cs] * drgb[2];
      ddir[0] += bx[k] * s; ddir[1] += by[k] * s; ddir[2] += bz[k] * s;
    }
  }
  const float dot = dx * ddir[0] + dy * ddir[1] + dz * ddir[2];
  dm[0] += (ddir[0] - dx * dot) * inv;
  dm[1] += (ddir[1] - dy * dot) * inv;
  dm[2] += (ddir[2] - dz * dot) * inv;
}

// coefficient part: row (+)= basis(dir) x drgb
__device__ __forceinline__ void sh_coeff_backward(const S360View& v, float* sh, float mx, float my, float mz,
                                                  const float* campos, const float* drgb, bool first) {
  const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
  const float inv = 1.f / sqrtf(ox * ox + oy * oy + oz * oz);
  float b[25];
  const int deg = min(v.sh_degree, v.max_sh_degree);
  const int n = sh_basis(deg, ox * inv, oy * inv, oz * inv, b);
  const int ks = v.sh_layout ? 1 : 3, cs = v.sh_layout ? v.M : 1;
  if (first) {
#pragma unroll
    for (int k = 0; k < 25; k++) {
      if (k < n) { sh[ks * k] = b[k] * drgb[0]; sh[ks * k + cs] = b[k] * drgb[1]; sh[ks * k + 2 * cs] = b[k] * drgb[2]; }
    }
    for (int k = n; k < v.M; k++) { sh[ks * k] = 0.f; sh[ks * k + cs] = 0.f; sh[ks * k + 2 * cs] = 0.f; }
  } else {
#pragma unroll
    for (int k = 0; k < 25; k++) {
      if (k < n) { sh[ks * k] += b[k] * drgb[0]; sh[ks * k + cs] += b[k] * drgb[1]; sh[ks * k + 2 * cs] += b[k] * drgb[2]; }
    }
  }
}

// Batched K8 + K9: one thread per Gaussian folds the moments of all its pairs into ONE set of gradients
// (the reference gets the same sum from autograd over V separate rasterizer calls).
template <int MODE, bool DEPTH>
__global__ void __launch_bounds__(PRE_THREADS, S360_MV_BWD_MINB)
preprocess_multi_backward_kernel(const S360View v, const int NV, const float* __restrict__ means,
                                 const float* __restrict__ cov3D, const float* __restrict__ opac,
                                 const float* __restrict__ shs, GeomState gs, PairState ps,
                                 const float* __restrict__ acc, float* __restrict__ d_means,
                                 float* __restrict__ d_cov, float* __restrict__ d_opac, float* __restrict__ d_shs,
                                 float* __restrict__ d_colors, const DepthSpec dspec) {
  extern __shared__ __align__(128) float s_sh[];   // [PRE_THREADS][M*3]: SH in, dL/dSH out (in place)
  __shared__ uint64_t s_bar;
  __shared__ float s_cam[S360_MAX_VIEWS][CAM_F];
  __shared__ int s_same;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = v.P;
  const int row = v.M * 3;
  const int rows = min(PRE_THREADS, P - blockIdx.x * PRE_THREADS);
  const float* sh_src = shs ? shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  float* dsh_dst = shs ? d_shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  const uint32_t sh_bytes = (uint32_t)rows * row * 4u;
  const bool bulk_ok = shs && (sh_bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(sh_src) & 15u) == 0u) &&
                       ((reinterpret_cast<uintptr_t>(dsh_dst) & 15u) == 0u);
  const uint32_t mask = idx < P ? ps.mask[idx] : 0u;
  const uint32_t slot0 = idx < P ? ps.base[idx] : 0u;   // loaded with the mask: one dependent global load less before acc
  const bool vis = mask != 0u;
  stage_cameras(v, NV, MODE == S360_MODE_PINHOLE, s_cam);
  if (threadIdx.x == 0) {
    if (shs) mbar_init(&s_bar, 1);
    int same = 1;   // all views share one camera centre (cube faces): one SH evaluation serves every view
    for (int k = 1; k < NV; k++)
      for (int j = 0; j < 3; j++) same &= (__ldg(v.campos + 3 * k + j) == __ldg(v.campos + j)) ? 1 : 0;
    s_same = same;
  }
  const bool need = __syncthreads_or(vis) != 0;
  if (shs) {
    if (need) {
      if (bulk_ok) {
        if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, sh_bytes); bulk_load(s_sh, sh_src, sh_bytes, &s_bar); }
      } else {
        for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) s_sh[i] = sh_src[i];
        __syncthreads();
      }
    } else {
      for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) dsh_dst[i] = 0.f;
    }
  }
  const bool same_cam = s_same != 0;
  float dm[3] = {0.f, 0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dop = 0.f, dcol[3] = {0.f, 0.f, 0.f};
  if (vis) {
    const float sc = v.scene_scale;
    const float mx = means[3 * idx] * sc, my = means[3 * idx + 1] * sc, mz = means[3 * idx + 2] * sc;
    float cv[6];
    load_cov6(v, cov3D, idx, cv);
    const float op = opac[idx];
    float* sh = s_sh + threadIdx.x * row;
    if (shs != nullptr && bulk_ok) mbar_wait(&s_bar, 0);
    uint32_t slot = slot0;
    int first_view = -1;
    for (uint32_t m = mask; m; m &= m - 1, slot++) {
      const int view = __ffs(m) - 1;
      if (first_view < 0) first_view = view;
      const float* cam = s_cam[view];
      const float4 a0 = *reinterpret_cast<const float4*>(acc + (size_t)slot * ACC_STRIDE);
      const float4 a1 = *reinterpret_cast<const float4*>(acc + (size_t)slot * ACC_STRIDE + 4);
#if S360_BWD_QPRIME
      dop += op > 0.f ? acc[(size_t)slot * ACC_STRIDE + 8] / op : 0.f;
#else
      dop += acc[(size_t)slot * ACC_STRIDE + 8];
#endif
      float dmv[3], dm2v[2], dcv[6];
      view_backward<MODE, DEPTH>(v, cam, cam + 16, mx, my, mz, cv, op, a0, a1, dmv, dm2v, dcv, dspec,
                                 DEPTH ? acc[(size_t)slot * ACC_STRIDE + 9] : 0.f);
#pragma unroll
      for (int k = 0; k < 3; k++) dm[k] += dmv[k];
#pragma unroll
      for (int k = 0; k < 6; k++) dcov[k] += dcv[k];
      if (shs != nullptr) {
        const uint8_t cl = gs.clamped[slot];
        const float drgb[3] = {(cl & 1) ? 0.f : a0.x, (cl & 2) ? 0.f : a0.y, (cl & 4) ? 0.f : a0.z};
        if (same_cam) { dcol[0] += drgb[0]; dcol[1] += drgb[1]; dcol[2] += drgb[2]; }
        else sh_dir_backward(v, sh, mx, my, mz, cam + 32, drgb, dm);
      } else {
        dcol[0] += a0.x; dcol[1] += a0.y; dcol[2] += a0.z;
      }
    }
    if (shs != nullptr) {
      if (same_cam) {
        sh_dir_backward(v, sh, mx, my, mz, s_cam[first_view] + 32, dcol, dm);
        sh_coeff_backward(v, sh, mx, my, mz, s_cam[first_view] + 32, dcol, true);
      } else {
        // second sweep: every SH value has been read, the row can now take the gradient
        slot = slot0;
        bool first = true;
        for (uint32_t m = mask; m; m &= m - 1, slot++) {
          const int view = __ffs(m) - 1;
          const float4 a0 = *reinterpret_cast<const float4*>(acc + (size_t)slot * ACC_STRIDE);
          const uint8_t cl = gs.clamped[slot];
          const float drgb[3] = {(cl & 1) ? 0.f : a0.x, (cl & 2) ? 0.f : a0.y, (cl & 4) ? 0.f : a0.z};
          sh_coeff_backward(v, sh, mx, my, mz, s_cam[view] + 32, drgb, first);
          first = false;
        }
      }
    }
  } else if (shs != nullptr && need && idx < P) {
    if (bulk_ok) mbar_wait(&s_bar, 0);
    float* sh = s_sh + threadIdx.x * row;
    for (int k = 0; k < row; k++) sh[k] = 0.f;
  }
  if (shs != nullptr && need) {
    if (bulk_ok) {
      if (idx >= P) mbar_wait(&s_bar, 0);
      fence_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) { bulk_store(dsh_dst, s_sh, sh_bytes); bulk_store_wait_read(); }
    } else {
      __syncthreads();
      for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) dsh_dst[i] = s_sh[i];
    }
  }
  if (idx >= P) return;
  const float gsc = v.scene_scale;
#pragma unroll
  for (int k = 0; k < 3; k++) d_means[3 * idx + k] = dm[k] * gsc;
  store_dcov(v, d_cov, idx, dcov, gsc * gsc);
  d_opac[idx] = dop;
  if (d_colors != nullptr) {
    const bool pre = shs == nullptr;
#pragma unroll
    for (int k = 0; k < 3; k++) d_colors[3 * idx + k] = pre ? dcol[k] : 0.f;
  }
}

int launch_preprocess_multi_backward(const S360View& v, int NV, const float* means, const float* cov, const float* opac,
                                     const float* shs, GeomState g, PairState ps, const float* acc, float* d_means,
                                     float* d_cov, float* d_opac, float* d_shs, float* d_colors, int has_depth,
                                     int depth_mode, float depth_near, float depth_far, cudaStream_t st) {
  if (v.P == 0) return 0;
  const int grid = (v.P + PRE_THREADS - 1) / PRE_THREADS;
  const size_t smem = shs ? (size_t)PRE_THREADS * v.M * 3 * sizeof(float) : 0;
  if (smem > 190 * 1024) return S360_ERR_UNSUPPORTED;
  DepthSpec ds;
  ds.mode = depth_mode; ds.inv_scale = 1.f / v.scene_scale; ds.near = depth_near; ds.far = depth_far;
#define S360_LAUNCH_MK8(MODE_, DEPTH_) do { \
    if (smem > 40 * 1024) cudaFuncSetAttribute(preprocess_multi_backward_kernel<MODE_, DEPTH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    preprocess_multi_backward_kernel<MODE_, DEPTH_><<<grid, PRE_THREADS, smem, st>>>(v, NV, means, cov, opac, shs, g, ps, acc, d_means, d_cov, d_opac, d_shs, d_colors, ds); } while (0)
  if (v.mode == S360_MODE_PINHOLE) { if (has_depth) S360_LAUNCH_MK8(S360_MODE_PINHOLE, true); else S360_LAUNCH_MK8(S360_MODE_PINHOLE, false); }
  else { if (has_depth) S360_LAUNCH_MK8(S360_MODE_ERP, true); else S360_LAUNCH_MK8(S360_MODE_ERP, false); }
#undef S360_LAUNCH_MK8
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__global__ void mark_visible_kernel(const S360View v, const float* __restrict__ means, uint8_t* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= v.P) return;
  const float* V = v.viewmatrix;
  const float sc = v.scene_scale;
  const float mx = means[3 * idx] * sc, my = means[3 * idx + 1] * sc, mz = means[3 * idx + 2] * sc;
  const float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
  const float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
  const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
  const float d = v.mode == S360_MODE_PINHOLE ? tz : sqrtf(tx * tx + ty * ty + tz * tz);
  present[idx] = d > v.near_cull ? 1 : 0;
}

int launch_mark_visible(const S360View& v, const float* means, uint8_t* present, cudaStream_t st) {
  if (v.P == 0) return 0;
  mark_visible_kernel<<<(v.P + 255) / 256, 256, 0, st>>>(v, means, present);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace s360

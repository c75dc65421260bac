// preprocess.cu -- per-Gaussian kernels.
//   K1  preprocess_kernel          : world->view, pinhole / equirectangular projection, EWA cov2D,
//                                    conic, extent, tile rect, SH -> RGB        
//   K8+K9 preprocess_backward_kernel: dL/d(conic, mean2D, rgb, opacity) -> dL/d(cov3D, mean3D, SH)
//   K10 mark_visible_kernel
// Replaces upstream preprocessCUDA / computeCov2DCUDA / preprocessCUDA(bwd) / checkFrustum behind
// /root/reference/src/model/decoder/cuda_splatting.py:113-124.
#include "common.cuh"
#include "persplat.cuh"
#include "render_cull.cuh"

namespace s360 {

constexpr int PRE_THREADS = 128;

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
      if (want_color) {
        float J[9];
        sh_to_rgb_jac(v, s_sh + threadIdx.x * row, mx, my, mz, cam.cam, col, cl, J);
#pragma unroll
        for (int k = 0; k < 9; k++) gs.sh_jac[9 * (size_t)idx + k] = J[k];
      }
    }
  } else if (want_color) {
    col[0] = colors[3 * idx]; col[1] = colors[3 * idx + 1]; col[2] = colors[3 * idx + 2];
  }
  if (want_color) {
    float4 r0, r1, r2;
    pack_record(px, py, cA, cB, cC, op, hx, hy, col, sortkey, v.image_width, r0, r1, r2);
    gs.rec[3 * (size_t)idx + 0] = r0;
    gs.rec[3 * (size_t)idx + 1] = r1;
    gs.rec[3 * (size_t)idx + 2] = r2;
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
  // opt in to the dynamic size whenever dynamic + static shared memory may pass the 48 KB default
  if (v.mode == S360_MODE_PINHOLE) {
    if (smem > 40 * 1024) cudaFuncSetAttribute(preprocess_kernel<S360_MODE_PINHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    preprocess_kernel<S360_MODE_PINHOLE><<<grid, PRE_THREADS, smem, st>>>(v, means, cov, opac, shs, colors, g, radii, depth_keys, ids, counters);
  } else {
    if (smem > 40 * 1024) cudaFuncSetAttribute(preprocess_kernel<S360_MODE_ERP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    preprocess_kernel<S360_MODE_ERP><<<grid, PRE_THREADS, smem, st>>>(v, means, cov, opac, shs, colors, g, radii, depth_keys, ids, counters);
  }
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K8 + K9 fused.  acc[idx*12 + 0..8] = {dL/dr, dL/dg, dL/db, and the moments sum(q dx), sum(q dy), sum(q dx^2),
// sum(q dx dy), sum(q dy^2), sum(q)} with q = G dL/dalpha, accumulated by render_backward_kernel.
// The SH coefficients are NOT read here: dL/dSH = basis(dir) x dL/drgb needs only the direction, and the direction
// gradient of the colour uses the 3x3 Jacobian K1 stored (gs.sh_jac) -- 36 B instead of the 300-B SH row per Gaussian.
// The CTA's dL/dSH block (128 x 300 B) is assembled in shared memory and leaves by one TMA bulk store.
template <int MODE, bool DEPTH>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_backward_kernel(const S360View v, const float* __restrict__ means, const float* __restrict__ cov3D,
                           const float* __restrict__ opac, const bool has_sh, GeomState gs, const int32_t* __restrict__ radii,
                           const float* __restrict__ acc, float* __restrict__ d_means,
                           float* __restrict__ d_means2D, float* __restrict__ d_cov, float* __restrict__ d_opac,
                           float* __restrict__ d_shs, float* __restrict__ d_colors, const DepthSpec dspec) {
  pdl_enter();
  extern __shared__ __align__(128) float s_sh[];   // [PRE_THREADS][M*3]: dL/dSH of this CTA
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = v.P;
  const int row = v.M * 3;
  const int rows = min(PRE_THREADS, P - blockIdx.x * PRE_THREADS);
  float* dsh_dst = has_sh ? d_shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  const uint32_t sh_bytes = (uint32_t)rows * row * 4u;
  const bool bulk_ok = has_sh && (sh_bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(dsh_dst) & 15u) == 0u);
  const bool vis = idx < P && radii[idx] > 0;
  bool need = false;
  if (has_sh) {
    need = __syncthreads_or(vis);
    if (!need) {
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
    if (has_sh) {
      const float ox = mx - cam.cam[0], oy = my - cam.cam[1], oz = mz - cam.cam[2];
      const float inv = 1.f / sqrtf(ox * ox + oy * oy + oz * oz);
      const float dx = ox * inv, dy = oy * inv, dz = oz * inv;
      const int deg = min(v.sh_degree, v.max_sh_degree);
      const uint8_t cl = gs.clamped[idx];
      const float drgb[3] = {(cl & 1) ? 0.f : dcol[0], (cl & 2) ? 0.f : dcol[1], (cl & 4) ? 0.f : dcol[2]};
      // direction gradient of the colour through K1's Jacobian, then through the normalisation
      const float* J = gs.sh_jac + 9 * (size_t)idx;
      float ddir[3];
#pragma unroll
      for (int a = 0; a < 3; a++) ddir[a] = J[a] * drgb[0] + J[3 + a] * drgb[1] + J[6 + a] * drgb[2];
      const float dot = dx * ddir[0] + dy * ddir[1] + dz * ddir[2];
      dm[0] += (ddir[0] - dx * dot) * inv;
      dm[1] += (ddir[1] - dy * dot) * inv;
      dm[2] += (ddir[2] - dz * dot) * inv;
      // dL/dSH row: basis(dir) x dL/drgb, zeros beyond the active degree
      float* sh = s_sh + threadIdx.x * row;
      const int ks = v.sh_layout ? 1 : 3, cs = v.sh_layout ? v.M : 1;
      const int n = sh_basis_each(deg, dx, dy, dz, [&](int k, float b) {
        sh[ks * k] = b * drgb[0]; sh[ks * k + cs] = b * drgb[1]; sh[ks * k + 2 * cs] = b * drgb[2]; });
      for (int k = n; k < v.M; k++) { sh[ks * k] = 0.f; sh[ks * k + cs] = 0.f; sh[ks * k + 2 * cs] = 0.f; }
    }
  } else if (has_sh && need && idx < P) {
    float* sh = s_sh + threadIdx.x * row;
    for (int k = 0; k < row; k++) sh[k] = 0.f;
  }
  if (has_sh && need) {
    if (bulk_ok) {
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
    const bool pre = !has_sh;
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
    if (smem > 40 * 1024) cudaFuncSetAttribute(preprocess_backward_kernel<MODE_, DEPTH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    launch_pdl(preprocess_backward_kernel<MODE_, DEPTH_>, dim3(grid), dim3(PRE_THREADS), smem, st, v, means, cov, opac, shs != nullptr, g, radii, acc, d_means, d_means2D, d_cov, d_opac, d_shs, d_colors, ds); } while (0)
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
  pdl_enter();
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
  pdl_enter();
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
  pdl_enter();
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
  float J[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
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
      if (!same) { clc = 0; sh_to_rgb_jac(v, s_sh + threadIdx.x * row, mx, my, mz, cam + 32, col, clc, J); col_view = view; }
#pragma unroll
      for (int k = 0; k < 9; k++) gs.sh_jac[9 * (size_t)slot + k] = J[k];   // per pair: the backward reads no SH
    } else if (col_view < 0) {
      col[0] = colors[3 * idx]; col[1] = colors[3 * idx + 1]; col[2] = colors[3 * idx + 2];
      col_view = view;
    }
    float4 r0, r1, r2;
    pack_record(pr.px, pr.py, pr.cA, pr.cB, pr.cC, pr.op, pr.hx, pr.hy, col, pr.sortkey, v.image_width, r0, r1, r2);
    gs.rec[3 * (size_t)slot + 0] = r0;
    gs.rec[3 * (size_t)slot + 1] = r1;
    gs.rec[3 * (size_t)slot + 2] = r2;
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
  launch_pdl(multi_scan_kernel, dim3(1), dim3(1024), 0, st, grid, block_count, cap, ps, counters);
  if (v.mode == S360_MODE_PINHOLE) {
    if (smem > 32 * 1024) cudaFuncSetAttribute(multi_write_kernel<S360_MODE_PINHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(multi_write_kernel<S360_MODE_PINHOLE>, dim3(grid), dim3(PRE_THREADS), smem, st, v, NV, cap, means, cov, opac, shs, colors, g, ps, block_count, depth_keys, ids);
  } else {
    if (smem > 32 * 1024) cudaFuncSetAttribute(multi_write_kernel<S360_MODE_ERP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(multi_write_kernel<S360_MODE_ERP>, dim3(grid), dim3(PRE_THREADS), smem, st, v, NV, cap, means, cov, opac, shs, colors, g, ps, block_count, depth_keys, ids);
  }
  count_launch(3);
  return (int)cudaGetLastError();
}

// zero the first min(*n_dev, cap) pair accumulators (the pair buffers are sized for the worst case V * P)
__global__ void zero_acc_kernel(float4* __restrict__ acc, const uint32_t* __restrict__ n_dev, int64_t cap) {
  pdl_enter();
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

// direction part of the SH backward for one pair: dL/d(mean) through the normalised view direction, from the Jacobian
// J[c][a] = d rgb_c / d dir_a that K1c stored for the pair
__device__ __forceinline__ void sh_dir_backward(const float* __restrict__ J, float mx, float my, float mz,
                                                const float* campos, const float* drgb, float* dm) {
  const float ox = mx - campos[0], oy = my - campos[1], oz = mz - campos[2];
  const float inv = 1.f / sqrtf(ox * ox + oy * oy + oz * oz);
  const float dx = ox * inv, dy = oy * inv, dz = oz * inv;
  float ddir[3];
#pragma unroll
  for (int a = 0; a < 3; a++) ddir[a] = J[a] * drgb[0] + J[3 + a] * drgb[1] + J[6 + a] * drgb[2];
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
  const int deg = min(v.sh_degree, v.max_sh_degree);
  const int ks = v.sh_layout ? 1 : 3, cs = v.sh_layout ? v.M : 1;
  if (first) {
    const int n = sh_basis_each(deg, ox * inv, oy * inv, oz * inv, [&](int k, float b) {
      sh[ks * k] = b * drgb[0]; sh[ks * k + cs] = b * drgb[1]; sh[ks * k + 2 * cs] = b * drgb[2]; });
    for (int k = n; k < v.M; k++) { sh[ks * k] = 0.f; sh[ks * k + cs] = 0.f; sh[ks * k + 2 * cs] = 0.f; }
  } else {
    sh_basis_each(deg, ox * inv, oy * inv, oz * inv, [&](int k, float b) {
      sh[ks * k] += b * drgb[0]; sh[ks * k + cs] += b * drgb[1]; sh[ks * k + 2 * cs] += b * drgb[2]; });
  }
}

// Batched K8 + K9: one thread per Gaussian folds the moments of all its pairs into ONE set of gradients
// (the reference gets the same sum from autograd over V separate rasterizer calls).  No SH coefficient is read: the
// direction gradient uses the per-pair Jacobians of K1c, dL/dSH = sum_v basis(dir_v) x dL/drgb_v is assembled in shared
// memory (one basis evaluation when all views share the camera centre) and leaves by one TMA bulk store.
template <int MODE, bool DEPTH>
__global__ void __launch_bounds__(PRE_THREADS, S360_MV_BWD_MINB)
preprocess_multi_backward_kernel(const S360View v, const int NV, const float* __restrict__ means,
                                 const float* __restrict__ cov3D, const float* __restrict__ opac,
                                 const bool has_sh, GeomState gs, PairState ps,
                                 const float* __restrict__ acc, float* __restrict__ d_means,
                                 float* __restrict__ d_cov, float* __restrict__ d_opac, float* __restrict__ d_shs,
                                 float* __restrict__ d_colors, const DepthSpec dspec) {
  pdl_enter();
  extern __shared__ __align__(128) float s_sh[];   // [PRE_THREADS][M*3]: dL/dSH of this CTA
  __shared__ float s_cam[S360_MAX_VIEWS][CAM_F];
  __shared__ int s_same;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int P = v.P;
  const int row = v.M * 3;
  const int rows = min(PRE_THREADS, P - blockIdx.x * PRE_THREADS);
  float* dsh_dst = has_sh ? d_shs + (size_t)blockIdx.x * PRE_THREADS * row : nullptr;
  const uint32_t sh_bytes = (uint32_t)rows * row * 4u;
  const bool bulk_ok = has_sh && (sh_bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(dsh_dst) & 15u) == 0u);
  const uint32_t mask = idx < P ? ps.mask[idx] : 0u;
  const uint32_t slot0 = idx < P ? ps.base[idx] : 0u;   // loaded with the mask: one dependent global load less before acc
  const bool vis = mask != 0u;
  stage_cameras(v, NV, MODE == S360_MODE_PINHOLE, s_cam);
  if (threadIdx.x == 0) {
    int same = 1;   // all views share one camera centre (cube faces): one basis evaluation serves every view
    for (int k = 1; k < NV; k++)
      for (int j = 0; j < 3; j++) same &= (__ldg(v.campos + 3 * k + j) == __ldg(v.campos + j)) ? 1 : 0;
    s_same = same;
  }
  const bool need = __syncthreads_or(vis) != 0;
  if (has_sh && !need)
    for (int i = threadIdx.x; i < rows * row; i += PRE_THREADS) dsh_dst[i] = 0.f;
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
    uint32_t slot = slot0;
    int first_view = -1;
#pragma unroll 1
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
      float dm2v[2];
      view_backward<MODE, DEPTH, true>(v, cam, cam + 16, mx, my, mz, cv, op, a0, a1, dm, dm2v, dcov, dspec,
                                       DEPTH ? acc[(size_t)slot * ACC_STRIDE + 9] : 0.f);
      if (has_sh) {
        const uint8_t cl = gs.clamped[slot];
        const float drgb[3] = {(cl & 1) ? 0.f : a0.x, (cl & 2) ? 0.f : a0.y, (cl & 4) ? 0.f : a0.z};
        if (same_cam) { dcol[0] += drgb[0]; dcol[1] += drgb[1]; dcol[2] += drgb[2]; }   // same direction, same Jacobian
        else {
          sh_dir_backward(gs.sh_jac + 9 * (size_t)slot, mx, my, mz, cam + 32, drgb, dm);
          sh_coeff_backward(v, sh, mx, my, mz, cam + 32, drgb, view == first_view);
        }
      } else {
        dcol[0] += a0.x; dcol[1] += a0.y; dcol[2] += a0.z;
      }
    }
    if (has_sh && same_cam) {
      sh_dir_backward(gs.sh_jac + 9 * (size_t)slot0, mx, my, mz, s_cam[first_view] + 32, dcol, dm);
      sh_coeff_backward(v, sh, mx, my, mz, s_cam[first_view] + 32, dcol, true);
    }
  } else if (has_sh && need && idx < P) {
    float* sh = s_sh + threadIdx.x * row;
    for (int k = 0; k < row; k++) sh[k] = 0.f;
  }
  if (has_sh && need) {
    if (bulk_ok) {
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
    const bool pre = !has_sh;
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
    launch_pdl(preprocess_multi_backward_kernel<MODE_, DEPTH_>, dim3(grid), dim3(PRE_THREADS), smem, st, v, NV, means, cov, opac, shs != nullptr, g, ps, acc, d_means, d_cov, d_opac, d_shs, d_colors, ds); } while (0)
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

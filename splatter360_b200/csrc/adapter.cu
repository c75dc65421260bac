// adapter.cu -- fused Gaussian adapter: encoder head outputs -> rasterizer inputs in ONE pass.
//   adapter_forward_kernel : raw features [G, 7 + 3 d_sh] + depth [G] + per-view pose / SH rotation ->
//                            means [G,3], covariances [G,3,3], harmonics [G,3,d_sh] (+ scales [G,3], rotations [G,4])
//   adapter_backward_kernel: cotangents of those -> d raw [G, 7 + 3 d_sh], d depth [G]
// Replaces the ~25 elementwise / einsum launches of /root/reference/src/model/encoder/common/gaussian_adapter_erp.py:49-119
// (+ gaussians.py:8-44, sphere_projection.py:6-87, sh_rotation.py:10-30), which stream every Gaussian through HBM a dozen
// times, by one read of 332 B and one write of 360 B per Gaussian.  The CTA's block of raw rows (128 x 328 B) is staged in
// shared memory by one TMA bulk copy; results go back through the same block so that the stores are coalesced.
// G = B * V * H * W Gaussians in (batch, view, row, col) order: one per context pixel (gaussians_per_pixel = num_surfaces = 1).
#include "adapter_math.cuh"

namespace s360 {

constexpr int AD_THREADS = 128;

__global__ void __launch_bounds__(AD_THREADS)
adapter_forward_kernel(const AdapterCfg cfg, const int64_t G, const float* __restrict__ raw, const float* __restrict__ depth,
                       const float* __restrict__ pose, const float* __restrict__ rot, float* __restrict__ means,
                       float* __restrict__ cov, float* __restrict__ harmonics, float* __restrict__ scales,
                       float* __restrict__ rotations) {
  extern __shared__ __align__(128) float s_raw[];   // [AD_THREADS][C]
  __shared__ uint64_t s_bar;
  const int C = 7 + 3 * cfg.d_sh;
  const int64_t g0 = (int64_t)blockIdx.x * AD_THREADS;
  const int rows = (int)min((int64_t)AD_THREADS, G - g0);
  const float* src = raw + g0 * C;
  const uint32_t bytes = (uint32_t)rows * C * 4u;
  const bool bulk_ok = (bytes % 16u == 0u) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0u);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (bulk_ok) {
    if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, bytes); bulk_load(s_raw, src, bytes, &s_bar); }
    mbar_wait(&s_bar, 0);
  } else {
    for (int i = threadIdx.x; i < rows * C; i += AD_THREADS) s_raw[i] = src[i];
    __syncthreads();
  }
  const int64_t g = g0 + threadIdx.x;
  if (g < G) {
    const int64_t hw = (int64_t)cfg.H * cfg.W;
    const int64_t bv = g / hw;
    const int pix = (int)(g - bv * hw);
    const int row = pix / cfg.W, col = pix - row * cfg.W;
    float p[AD_POSE_F];
#pragma unroll
    for (int i = 0; i < AD_POSE_F; i++) p[i] = __ldg(pose + bv * AD_POSE_F + i);
    float* r = s_raw + threadIdx.x * C;
    float mean[3], cv[9];
    AdapterFwd f;
    adapter_forward_one(cfg, r, depth[g], p, row, col, mean, cv, f);
#pragma unroll
    for (int k = 0; k < 3; k++) means[3 * g + k] = mean[k];
#pragma unroll
    for (int k = 0; k < 9; k++) cov[9 * g + k] = cv[k];
    if (scales) {
#pragma unroll
      for (int k = 0; k < 3; k++) scales[3 * g + k] = f.s[k];
    }
    if (rotations) {
#pragma unroll
      for (int k = 0; k < 4; k++) rotations[4 * g + k] = f.qn[k];
    }
    const float* D = rot + bv * AD_ROT_F;
    for (int ch = 0; ch < 3; ch++) adapter_rotate_sh<false>(cfg.sh_degree, D, r + 7 + ch * cfg.d_sh);
  }
  __syncthreads();
  // coalesced copy-out of the rotated harmonics: element e of the CTA's [rows][3 d_sh] block
  const int n_sh = 3 * cfg.d_sh;
  float* dst = harmonics + g0 * n_sh;
  for (int e = threadIdx.x; e < rows * n_sh; e += AD_THREADS) {
    const int rr = e / n_sh, k = e - rr * n_sh;
    dst[e] = s_raw[rr * C + 7 + k];
  }
}

__global__ void __launch_bounds__(AD_THREADS)
adapter_backward_kernel(const AdapterCfg cfg, const int64_t G, const float* __restrict__ raw, const float* __restrict__ depth,
                        const float* __restrict__ pose, const float* __restrict__ rot, const float* __restrict__ g_means,
                        const float* __restrict__ g_cov, const float* __restrict__ g_harmonics, float* __restrict__ d_raw,
                        float* __restrict__ d_depth) {
  extern __shared__ __align__(128) float s_out[];   // [AD_THREADS][C]: harmonics cotangent in, d raw out
  const int C = 7 + 3 * cfg.d_sh;
  const int n_sh = 3 * cfg.d_sh;
  const int64_t g0 = (int64_t)blockIdx.x * AD_THREADS;
  const int rows = (int)min((int64_t)AD_THREADS, G - g0);
  {
    const float* src = g_harmonics ? g_harmonics + g0 * n_sh : nullptr;
    for (int e = threadIdx.x; e < rows * n_sh; e += AD_THREADS) {
      const int rr = e / n_sh, k = e - rr * n_sh;
      s_out[rr * C + 7 + k] = src ? src[e] : 0.f;
    }
  }
  __syncthreads();
  const int64_t g = g0 + threadIdx.x;
  if (g < G) {
    const int64_t hw = (int64_t)cfg.H * cfg.W;
    const int64_t bv = g / hw;
    const int pix = (int)(g - bv * hw);
    const int row = pix / cfg.W, col = pix - row * cfg.W;
    float p[AD_POSE_F];
#pragma unroll
    for (int i = 0; i < AD_POSE_F; i++) p[i] = __ldg(pose + bv * AD_POSE_F + i);
    float r7[7];
#pragma unroll
    for (int k = 0; k < 7; k++) r7[k] = raw[g * C + k];
    float mean[3], cv[9];
    AdapterFwd f;
    const float dep = depth[g];
    adapter_forward_one(cfg, r7, dep, p, row, col, mean, cv, f);
    float gm[3] = {0.f, 0.f, 0.f}, gc[9];
    if (cfg.means_grad && g_means) {
#pragma unroll
      for (int k = 0; k < 3; k++) gm[k] = g_means[3 * g + k];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) gc[k] = g_cov ? g_cov[9 * g + k] : 0.f;
    float d7[7], dd;
    adapter_backward_one(cfg, r7, dep, p, f, gm, gc, d7, dd);
    float* o = s_out + threadIdx.x * C;
#pragma unroll
    for (int k = 0; k < 7; k++) o[k] = d7[k];
    d_depth[g] = dd;
    const float* D = rot + bv * AD_ROT_F;
    for (int ch = 0; ch < 3; ch++) adapter_rotate_sh<true>(cfg.sh_degree, D, o + 7 + ch * cfg.d_sh);
  }
  __syncthreads();
  float* dst = d_raw + g0 * C;
  for (int e = threadIdx.x; e < rows * C; e += AD_THREADS) dst[e] = s_out[e];
}

static bool adapter_cfg_ok(const AdapterCfg& c) {
  return c.H > 0 && c.W > 0 && c.sh_degree >= 0 && c.sh_degree <= 4 && c.d_sh == (c.sh_degree + 1) * (c.sh_degree + 1);
}

int launch_adapter_forward(const AdapterCfg& cfg, int64_t G, const float* raw, const float* depth, const float* pose,
                           const float* rot, float* means, float* cov, float* harmonics, float* scales, float* rotations,
                           cudaStream_t st) {
  if (!adapter_cfg_ok(cfg)) return S360_ERR_BAD_ARGUMENT;
  if (G == 0) return 0;
  const size_t smem = (size_t)AD_THREADS * (7 + 3 * cfg.d_sh) * sizeof(float);
  if (smem > 40 * 1024) cudaFuncSetAttribute(adapter_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  adapter_forward_kernel<<<(unsigned)((G + AD_THREADS - 1) / AD_THREADS), AD_THREADS, smem, st>>>(
      cfg, G, raw, depth, pose, rot, means, cov, harmonics, scales, rotations);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_adapter_backward(const AdapterCfg& cfg, int64_t G, const float* raw, const float* depth, const float* pose,
                            const float* rot, const float* g_means, const float* g_cov, const float* g_harmonics,
                            float* d_raw, float* d_depth, cudaStream_t st) {
  if (!adapter_cfg_ok(cfg)) return S360_ERR_BAD_ARGUMENT;
  if (G == 0) return 0;
  const size_t smem = (size_t)AD_THREADS * (7 + 3 * cfg.d_sh) * sizeof(float);
  if (smem > 40 * 1024) cudaFuncSetAttribute(adapter_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  adapter_backward_kernel<<<(unsigned)((G + AD_THREADS - 1) / AD_THREADS), AD_THREADS, smem, st>>>(
      cfg, G, raw, depth, pose, rot, g_means, g_cov, g_harmonics, d_raw, d_depth);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace s360

using namespace s360;

extern "C" {

static AdapterCfg make_cfg(int32_t H, int32_t W, int32_t sh_degree, float scale_min, float scale_max, int32_t means_grad) {
  AdapterCfg c;
  c.H = H; c.W = W; c.sh_degree = sh_degree; c.d_sh = (sh_degree + 1) * (sh_degree + 1);
  c.scale_min = scale_min; c.scale_max = scale_max;
  c.pixel_size = 1.f / (float)(W > H ? W : H);   // gaussian_adapter_erp.py:72
  c.eps = 1e-8f;                                   // :60, gaussians.py:11
  c.means_grad = means_grad;
  return c;
}

int s360_adapter_forward(int32_t views, int32_t H, int32_t W, int32_t sh_degree, float scale_min, float scale_max,
                         const float* raw, const float* depth, const float* pose, const float* sh_rot, float* means,
                         float* covariances, float* harmonics, float* scales, float* rotations, void* stream) {
  if (views < 0 || H <= 0 || W <= 0) return S360_ERR_BAD_ARGUMENT;
  const int64_t G = (int64_t)views * H * W;
  if (G > 0 && (!raw || !depth || !pose || !sh_rot || !means || !covariances || !harmonics)) return S360_ERR_BAD_ARGUMENT;
  return launch_adapter_forward(make_cfg(H, W, sh_degree, scale_min, scale_max, 0), G, raw, depth, pose, sh_rot, means,
                                covariances, harmonics, scales, rotations, (cudaStream_t)stream);
}

int s360_adapter_backward(int32_t views, int32_t H, int32_t W, int32_t sh_degree, float scale_min, float scale_max,
                          int32_t means_grad, const float* raw, const float* depth, const float* pose, const float* sh_rot,
                          const float* dL_dmeans, const float* dL_dcovariances, const float* dL_dharmonics, float* dL_draw,
                          float* dL_ddepth, void* stream) {
  if (views < 0 || H <= 0 || W <= 0) return S360_ERR_BAD_ARGUMENT;
  const int64_t G = (int64_t)views * H * W;
  if (G > 0 && (!raw || !depth || !pose || !sh_rot || !dL_draw || !dL_ddepth)) return S360_ERR_BAD_ARGUMENT;
  return launch_adapter_backward(make_cfg(H, W, sh_degree, scale_min, scale_max, means_grad), G, raw, depth, pose, sh_rot,
                                 dL_dmeans, dL_dcovariances, dL_dharmonics, dL_draw, dL_ddepth, (cudaStream_t)stream);
}

}  // extern "C"

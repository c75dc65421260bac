// loss.cu -- fused MSE loss + seed gradient for the rendered image.
// The reference's photometric loss is  weight * ((prediction - target)^2).mean()
// (/root/reference/src/loss/loss_mse.py:22-31); its gradient w.r.t. the rendered colour is the seed of the rasterizer's
// backward pass.  One pass over the image produces both, instead of the half-dozen
// elementwise / reduction launches autograd would issue.
#include "common.cuh"

namespace s360 {

__global__ void __launch_bounds__(256)
mse_loss_grad_kernel(const float4* __restrict__ color, const float4* __restrict__ target, const float* __restrict__ color_s,
                     const float* __restrict__ target_s, int64_t n4, int64_t n, float scale, float4* __restrict__ grad,
                     float* __restrict__ grad_s, float* __restrict__ loss) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 c = color[i], t = target[i];
    const float4 d = make_float4(c.x - t.x, c.y - t.y, c.z - t.z, c.w - t.w);
    acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    grad[i] = make_float4(2.f * scale * d.x, 2.f * scale * d.y, 2.f * scale * d.z, 2.f * scale * d.w);
  }
  // scalar tail (n not a multiple of 4), handled by the first threads
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = color_s[i] - target_s[i];
    acc += d * d;
    grad_s[i] = 2.f * scale * d;
  }
  __shared__ float s_w[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) t += s_w[w];
    atomicAdd(loss, t * scale);
  }
}

}  // namespace s360

using namespace s360;

extern "C" int s360_mse_loss_grad(const float* color, const float* target, int64_t n, float weight, float* loss,
                                  float* grad, void* stream) {
  if (n < 0 || !loss || (n > 0 && (!color || !target || !grad))) return S360_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = (int)cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (rc || n == 0) return rc;
  const bool aligned = ((reinterpret_cast<uintptr_t>(color) | reinterpret_cast<uintptr_t>(target) |
                         reinterpret_cast<uintptr_t>(grad)) & 15u) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const int blocks = (int)((n / 4 + 255) / 256 < 1184 ? ((n / 4 + 255) / 256 > 0 ? (n / 4 + 255) / 256 : 1) : 1184);
  mse_loss_grad_kernel<<<blocks, 256, 0, st>>>((const float4*)color, (const float4*)target, color, target, n4, n,
                                               weight / (float)n, (float4*)grad, grad, loss);
  count_launch();
  return (int)cudaGetLastError();
}

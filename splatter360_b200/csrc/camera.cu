// camera.cu -- batched inverse of 4x4 camera matrices on the device.
// The reference turns every camera-to-world pose into the rasterizer's view matrix with `extrinsics.inverse()`
// (/root/reference/src/model/decoder/cuda_splatting.py:84, :176, :262): per call a cuSOLVER/cuBLAS batched LU (several
// launches, pointer arrays uploaded from the host, a status read-back) for a handful of 4x4 matrices.  One thread per
// matrix here: Gauss-Jordan with partial pivoting in double precision, rounded once to float -- within half an ulp of the
// exact inverse, one launch, nothing on the host, capturable in a CUDA graph.  Singular input yields inf/nan like
// `torch.linalg.inv_ex` (no status).
#include "camera_math.cuh"

namespace s360 {

__global__ void __launch_bounds__(64)
invert4x4_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  invert4x4(in + i * 16, out + i * 16);
}

}  // namespace s360

using namespace s360;

extern "C" int s360_invert4x4(const float* in, float* out, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!in || !out))) return S360_ERR_BAD_ARGUMENT;
  if (n == 0) return 0;
  invert4x4_kernel<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(in, out, n);
  count_launch();
  return (int)cudaGetLastError();
}

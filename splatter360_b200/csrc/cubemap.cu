// cubemap.cu -- cube faces -> equirectangular stitch.
// The reference turns six rendered cube faces into a panorama with `change_order` (torch.flip / stack,
// /root/reference/src/model/model_wrapper_erp.py:135-158), a torch.cat into a [b,c,f,6f] strip (:395-398) and
// Cube2Equirec.forward = a 5-D F.grid_sample over a precomputed sample grid (/root/reference/src/geometry/layers.py:
// 108-116).  Here it is one gather kernel reading the rasterizer's [6,3,f,f] output directly (face reorder and the
// up/down flips are index arithmetic), and one scatter kernel for the gradient.  The sample grid is the module's
// own buffer (u, v in [-1,1], face = (z + 1) * 2.5), so the face selection logic stays pinned to the reference.
// HBM-bound: 12 B grid + 4*C taps (L2-resident faces) + 4*C B out per panorama pixel.
// Optional z-depth -> distance conversion (the reference's depth video: depth_to_distance_map_batch on the reordered
// faces, /root/reference/src/geometry/z_depth_to_distance.py:4-34, before the stitch, model_wrapper_erp.py:447-463):
// every tap is scaled by sqrt(((r - cx)/fx)^2 + ((c - cy)/fy)^2 + 1) of its (reordered-face) texel.
#include "common.cuh"

namespace s360 {

// source offset of texel (row r, col c) of face `face` (order [F R B L U D]) and channel ch
//   layout 0: strip  [B, C, f, 6f]            -- the reference module's input
//   layout 1: faces  [B, 6, C, f, f] in the dataset order [U B L F R D]; U and D flipped along both image axes
__device__ __forceinline__ size_t face_texel(int layout, int b, int ch, int C, int f, int face, int r, int c) {
  if (layout == 0) return (((size_t)b * C + ch) * f + r) * (6 * (size_t)f) + (size_t)face * f + c;
  // [F R B L U D] <- dataset index [3 4 1 2 0 5]
  const int src = face == 0 ? 3 : face == 1 ? 4 : face == 2 ? 1 : face == 3 ? 2 : face == 4 ? 0 : 5;
  if (face >= 4) { r = f - 1 - r; c = f - 1 - c; }
  return ((((size_t)b * 6 + src) * C + ch) * f + r) * (size_t)f + c;
}

// distance / z-depth of texel (r, c).  Literal to the reference: integer pixel coordinates, and its
// torch.meshgrid(arange(width), arange(height)) is 'ij'-indexed, so "u" (paired with cx, fx) runs along the ROWS of the
// square face and "v" (cy, fy) along the columns -- indistinguishable for its cube faces (cx = cy, fx = fy).
struct D2D { float fx, fy, cx, cy; int on; };
__device__ __forceinline__ float d2d_factor(const D2D& d, int r, int c) {
  if (!d.on) return 1.f;
  const float x = ((float)r - d.cx) / d.fx, y = ((float)c - d.cy) / d.fy;
  return sqrtf(x * x + y * y + 1.f);
}

struct Tap {
  int face, x0, x1, y0, y1;
  float wx, wy;
};
__device__ __forceinline__ Tap make_tap(const float* __restrict__ grid, size_t pix, int f) {
  const float u = grid[3 * pix], v = grid[3 * pix + 1], z = grid[3 * pix + 2];
  Tap t;
  t.face = min(5, max(0, __float2int_rn((z + 1.f) * 2.5f)));
  // align_corners=True, padding_mode="border": clip the unnormalised coordinate, then interpolate
  const float fx = fminf(fmaxf((u + 1.f) * 0.5f * (float)(f - 1), 0.f), (float)(f - 1));
  const float fy = fminf(fmaxf((v + 1.f) * 0.5f * (float)(f - 1), 0.f), (float)(f - 1));
  t.x0 = (int)floorf(fx); t.y0 = (int)floorf(fy);
  t.x1 = min(t.x0 + 1, f - 1); t.y1 = min(t.y0 + 1, f - 1);
  t.wx = fx - (float)t.x0; t.wy = fy - (float)t.y0;
  return t;
}

__global__ void __launch_bounds__(256)
cube2equirec_forward_kernel(const float* __restrict__ faces, const float* __restrict__ grid, int layout, int B, int C,
                            int f, int H, int W, float* __restrict__ out, const D2D d2d) {
  const size_t npix = (size_t)H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // pixel of one batch item
  const int b = blockIdx.y;
  if (i >= npix) return;
  const Tap t = make_tap(grid, i, f);
  const float w00 = (1.f - t.wx) * (1.f - t.wy) * d2d_factor(d2d, t.y0, t.x0), w01 = t.wx * (1.f - t.wy) * d2d_factor(d2d, t.y0, t.x1),
              w10 = (1.f - t.wx) * t.wy * d2d_factor(d2d, t.y1, t.x0), w11 = t.wx * t.wy * d2d_factor(d2d, t.y1, t.x1);
  for (int ch = 0; ch < C; ch++) {
    const float a = faces[face_texel(layout, b, ch, C, f, t.face, t.y0, t.x0)];
    const float bb = faces[face_texel(layout, b, ch, C, f, t.face, t.y0, t.x1)];
    const float c = faces[face_texel(layout, b, ch, C, f, t.face, t.y1, t.x0)];
    const float d = faces[face_texel(layout, b, ch, C, f, t.face, t.y1, t.x1)];
    out[((size_t)b * C + ch) * npix + i] = a * w00 + bb * w01 + c * w10 + d * w11;
  }
}

__global__ void __launch_bounds__(256)
cube2equirec_backward_kernel(const float* __restrict__ dout, const float* __restrict__ grid, int layout, int B, int C,
                             int f, int H, int W, float* __restrict__ dfaces, const D2D d2d) {
  const size_t npix = (size_t)H * W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= npix) return;
  const Tap t = make_tap(grid, i, f);
  const float w00 = (1.f - t.wx) * (1.f - t.wy) * d2d_factor(d2d, t.y0, t.x0), w01 = t.wx * (1.f - t.wy) * d2d_factor(d2d, t.y0, t.x1),
              w10 = (1.f - t.wx) * t.wy * d2d_factor(d2d, t.y1, t.x0), w11 = t.wx * t.wy * d2d_factor(d2d, t.y1, t.x1);
  for (int ch = 0; ch < C; ch++) {
    const float g = dout[((size_t)b * C + ch) * npix + i];
    atomicAdd(dfaces + face_texel(layout, b, ch, C, f, t.face, t.y0, t.x0), g * w00);
    if (w01 != 0.f) atomicAdd(dfaces + face_texel(layout, b, ch, C, f, t.face, t.y0, t.x1), g * w01);
    if (w10 != 0.f) atomicAdd(dfaces + face_texel(layout, b, ch, C, f, t.face, t.y1, t.x0), g * w10);
    if (w11 != 0.f) atomicAdd(dfaces + face_texel(layout, b, ch, C, f, t.face, t.y1, t.x1), g * w11);
  }
}

}  // namespace s360

using namespace s360;

extern "C" {

static D2D make_d2d(const float* k) {
  D2D d;
  d.on = k != nullptr;
  d.fx = k ? k[0] : 1.f; d.fy = k ? k[1] : 1.f; d.cx = k ? k[2] : 0.f; d.cy = k ? k[3] : 0.f;
  return d;
}

int s360_cube2equirec_forward(const float* faces, const float* grid, int32_t layout, int32_t B, int32_t C,
                              int32_t face_w, int32_t H, int32_t W, const float* depth_to_distance, float* out,
                              void* stream) {
  if (!faces || !grid || !out || (layout != 0 && layout != 1) || B < 0 || C <= 0 || face_w <= 0 || H <= 0 || W <= 0 || B > 65535)
    return S360_ERR_BAD_ARGUMENT;
  if (B == 0) return 0;
  const size_t npix = (size_t)H * W;
  dim3 grid_dim((unsigned)((npix + 255) / 256), (unsigned)B);
  cube2equirec_forward_kernel<<<grid_dim, 256, 0, (cudaStream_t)stream>>>(faces, grid, layout, B, C, face_w, H, W, out, make_d2d(depth_to_distance));
  count_launch();
  return (int)cudaGetLastError();
}

int s360_cube2equirec_backward(const float* dL_dout, const float* grid, int32_t layout, int32_t B, int32_t C,
                               int32_t face_w, int32_t H, int32_t W, const float* depth_to_distance, float* dL_dfaces,
                               void* stream) {
  if (!dL_dout || !grid || !dL_dfaces || (layout != 0 && layout != 1) || B < 0 || C <= 0 || face_w <= 0 || H <= 0 || W <= 0 || B > 65535)
    return S360_ERR_BAD_ARGUMENT;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = (int)cudaMemsetAsync(dL_dfaces, 0, (size_t)B * C * 6 * face_w * face_w * sizeof(float), st);
  if (rc) return rc;
  const size_t npix = (size_t)H * W;
  dim3 grid_dim((unsigned)((npix + 255) / 256), (unsigned)B);
  cube2equirec_backward_kernel<<<grid_dim, 256, 0, st>>>(dL_dout, grid, layout, B, C, face_w, H, W, dL_dfaces, make_d2d(depth_to_distance));
  count_launch();
  return (int)cudaGetLastError();
}

}  // extern "C"

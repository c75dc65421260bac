// common.cuh -- shared device helpers and state layouts for libsplatter360 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/splatter360.h"

namespace s360 {

constexpr int TILE = 16;                 // tile edge in pixels (upstream BLOCK_X = BLOCK_Y = 16)
constexpr int TILE_PIX = TILE * TILE;    // 256 threads per tile
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float T_EPS = 0.0001f;
constexpr float PI_F = 3.14159265358979323846f;

// Per-Gaussian geometry record: 3 x float4 = 48 B, gathered by the render kernels.
//   r0 = {x, y, conicA, conicB}   r1 = {conicC, opacity, hx, hy}   r2 = {r, g, b, depth}
// hx, hy: half extents of the axis-aligned box outside which the Gaussian cannot contribute.
constexpr int REC_F4 = 3;

// Geometry state carved out of the caller's `geom` buffer.
struct GeomState {
  float4* rec;        // [P*3]
  uint2* rect;        // [P] packed tile rect: x = (x0 & 0xffff) | nx << 16 ; y = y0 | ny << 16
  uint8_t* clamped;   // [P] bits 0..2 SH clamp per channel, bit 3 = jacobian clamp x, bit 4 = clamp y
  float* sh_jac;      // [P*9] d rgb_c / d dir_a of the SH colour (per pair in the batched path): K1 has the SH block in shared
                      // memory anyway, and with these 36 B the backward never reads the 300-B SH row again
};

struct ImageState {
  float* final_T;       // [H*W]
  uint32_t* n_contrib;  // [H*W]
  uint2* ranges;        // [tiles]
  uint32_t* order;      // [tiles] tile ids, heaviest first: the render CTAs take their tile from here
  uint32_t* work;       // [tiles] instances the forward pass actually walked (drives the backward schedule)
  uint32_t* order_bwd;  // [tiles]
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Pure per-Gaussian math (projection, covariance, SH, their derivatives) is host + device: the kernels inline it, and
// tests/host_harness compiles the very same functions for the CPU to check them against the oracle without a GPU.
#define S360_HD __host__ __device__ __forceinline__
S360_HD float s360_inf() {
#ifdef __CUDA_ARCH__
  return __int_as_float(0x7f800000);
#else
  return __builtin_huge_valf();
#endif
}
S360_HD uint32_t s360_float_bits(float x) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(x);
#else
  uint32_t u; __builtin_memcpy(&u, &x, 4); return u;
#endif
}

inline size_t geom_base_bytes(int P) {
  return align_up((size_t)P * REC_F4 * sizeof(float4), 256) + align_up((size_t)P * sizeof(uint2), 256) +
         align_up((size_t)P, 256);
}
inline GeomState carve_geom_base(void* buf, int P) {
  char* p = (char*)buf;
  GeomState g;
  g.rec = (float4*)p;      p += align_up((size_t)P * REC_F4 * sizeof(float4), 256);
  g.rect = (uint2*)p;      p += align_up((size_t)P * sizeof(uint2), 256);
  g.clamped = (uint8_t*)p; p += align_up((size_t)P, 256);
  g.sh_jac = nullptr;
  return g;
}
inline GeomState carve_geom(void* buf, int P) {
  GeomState g = carve_geom_base(buf, P);
  g.sh_jac = (float*)((char*)buf + geom_base_bytes(P));
  return g;
}
inline size_t geom_bytes(int P) { return geom_base_bytes(P) + align_up((size_t)P * 9 * sizeof(float), 256); }
// Per-Gaussian bookkeeping of the batched multi-view path: which views the Gaussian has a pair in, and where its
// pairs start in the compacted pair buffers (pair of view v = base + popc(mask & ((1 << v) - 1))).
struct PairState {
  uint32_t* base;   // [P]
  uint32_t* mask;   // [P]
  uint32_t* count;  // [1] pairs stored = min(pairs needed, pair capacity)
};
// geometry state of the batched path: pair-indexed GeomState for `cap` pairs followed by the per-Gaussian PairState
inline GeomState carve_geom_multi(void* buf, int P, int64_t cap, PairState* ps) {
  GeomState g = carve_geom_base(buf, (int)cap);
  char* p = (char*)buf + geom_base_bytes((int)cap);
  ps->base = (uint32_t*)p; p += align_up((size_t)P * 4, 256);
  ps->mask = (uint32_t*)p; p += align_up((size_t)P * 4, 256);
  ps->count = (uint32_t*)p; p += 256;
  g.sh_jac = (float*)p;   // [cap*9] per PAIR
  return g;
}
inline size_t geom_multi_bytes(int P, int64_t cap) {
  return geom_base_bytes((int)cap) + 2 * align_up((size_t)P * 4, 256) + 256 + align_up((size_t)cap * 9 * sizeof(float), 256);
}

// V views are laid out as one stacked image: pixel state [V][H][W], tiles [V][gy][gx] (V = 1: the single view)
inline ImageState carve_image(void* buf, int H, int W, int V = 1) {
  char* p = (char*)buf;
  ImageState s;
  size_t npix = (size_t)V * H * W;
  size_t tiles = (size_t)V * ((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
  s.final_T = (float*)p;      p += align_up(npix * 4, 256);
  s.n_contrib = (uint32_t*)p; p += align_up(npix * 4, 256);
  s.ranges = (uint2*)p;       p += align_up(tiles * 8, 256);
  s.order = (uint32_t*)p;     p += align_up(tiles * 4, 256);
  s.work = (uint32_t*)p;      p += align_up(tiles * 4, 256);
  s.order_bwd = (uint32_t*)p; p += align_up(tiles * 4, 256);
  return s;
}
inline size_t image_bytes(int H, int W, int V = 1) {
  size_t npix = (size_t)V * H * W;
  size_t tiles = (size_t)V * ((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
  return align_up(npix * 4, 256) * 2 + align_up(tiles * 8, 256) + 3 * align_up(tiles * 4, 256);
}

// ---------------------------------------------------------------------------------------------
// TMA bulk-copy (cp.async.bulk) + mbarrier helpers: one elected thread moves a whole contiguous
// block of per-Gaussian attributes between HBM and shared memory (SASS: UBLKCP / SYNCS).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
// global -> shared, completion signalled on `bar` (bytes % 16 == 0, both addresses 16-B aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global; call after fence_async_smem() + __syncthreads(); the issuing thread must wait
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// launch bookkeeping (introspection only)
void count_launch(int n = 1);

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled while its predecessor on the
// stream is still draining -- its CTAs become resident as the predecessor's retire and park at pdl_enter() until the
// predecessor's grid has completed and its writes are visible -- so the launch ramp of every kernel in the latency-bound
// sort / binning chain overlaps the tail of the kernel before it (also inside a captured CUDA graph).  Rules kept here:
// (1) pdl_enter() is the FIRST statement of every kernel launched through launch_pdl(), executed by all threads, so that
// "this grid completed" always implies "everything before it completed"; (2) a kernel never touches global memory before
// it.  pdl_enter() also releases the kernel's own dependents (they park the same way).  Launched normally, both
// instructions are no-ops.  S360_PDL=0 in the environment turns the launch attribute off.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// Fused depth channel: what the reference renders in a second full pass by feeding
// per-Gaussian depth as colour (cuda_splatting.py:226-269).  The geometry record stores the sort depth of the
// rescaled scene (camera z for pinhole, radial distance for erp); dividing by scene_scale recovers the value the
// reference computes from the unscaled means.
struct DepthSpec {
  int mode;          // S360_DEPTH_*
  float inv_scale;   // 1 / scene_scale
  float near, far;   // unscaled, for relative_disparity / log
};
S360_HD float depth_value(const DepthSpec& d, float rec_depth) {
  const float z = rec_depth * d.inv_scale;
  if (d.mode == S360_DEPTH_DISPARITY) return 1.f / z;
  if (d.mode == S360_DEPTH_RELATIVE_DISPARITY) {
    const float eps = 1e-10f;
    const float dn = 1.f / (d.near + eps), df = 1.f / (d.far + eps), dz = 1.f / (z + eps);
    return 1.f - (dz - df) / (dn - df + eps);
  }
  if (d.mode == S360_DEPTH_LOG) return logf(fmaxf(fminf(z, d.near), d.far));   // literal: .minimum(near).maximum(far).log()
  return z;
}
// d depth_value / d rec_depth (what autograd gives through the reference's torch expressions)
S360_HD float depth_value_grad(const DepthSpec& d, float rec_depth) {
  const float z = rec_depth * d.inv_scale;
  if (d.mode == S360_DEPTH_DISPARITY) return -d.inv_scale / (z * z);
  if (d.mode == S360_DEPTH_RELATIVE_DISPARITY) {
    const float eps = 1e-10f;
    const float dn = 1.f / (d.near + eps), df = 1.f / (d.far + eps), zi = 1.f / (z + eps);
    return d.inv_scale * zi * zi / (dn - df + eps);
  }
  if (d.mode == S360_DEPTH_LOG) {
    // gradient reaches z only where the clamps select it: z <= near and min(z, near) >= far
    const float m = fminf(z, d.near);
    return (z <= d.near && m >= d.far) ? d.inv_scale / z : 0.f;
  }
  return d.inv_scale;
}

// ---------------------------------------------------------------------------------------------
// camera block loaded once per thread from the tiny device-resident matrices
struct Cam {
  float V[16];
  float PM[16];
  float cam[3];
};

__device__ __forceinline__ void load_cam(const S360View& v, Cam& c, bool need_proj) {
#pragma unroll
  for (int i = 0; i < 16; i++) c.V[i] = __ldg(v.viewmatrix + i);
  if (need_proj) {
#pragma unroll
    for (int i = 0; i < 16; i++) c.PM[i] = __ldg(v.projmatrix + i);
  }
#pragma unroll
  for (int i = 0; i < 3; i++) c.cam[i] = __ldg(v.campos + i);
}

// Screen-space geometry of one Gaussian, shared by the forward and backward preprocess kernels.
struct Geo {
  float t[3];       // view-space centre
  float tc[3];      // centre used inside J (after fov / pole clamp)
  bool clampx, clampy;
  float J[2][3];
  float Mm[2][3];   // J * R
  float a, b, c;    // cov2D incl. low-pass
};

template <int MODE>
S360_HD void geo_compute(const S360View& v, const float* V, float mx, float my, float mz,
                                            const float* cov, Geo& g) {
  // R[i][k] = V[4k + i]
  g.t[0] = V[0] * mx + V[4] * my + V[8] * mz + V[12];
  g.t[1] = V[1] * mx + V[5] * my + V[9] * mz + V[13];
  g.t[2] = V[2] * mx + V[6] * my + V[10] * mz + V[14];
  const float x = g.t[0], y = g.t[1], z = g.t[2];
  g.clampx = g.clampy = false;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) g.J[i][k] = 0.f;
  if (MODE == S360_MODE_PINHOLE) {
    const float fx = (float)v.image_width / (2.f * v.tanfovx), fy = (float)v.image_height / (2.f * v.tanfovy);
    const float limx = v.fov_clamp * v.tanfovx, limy = v.fov_clamp * v.tanfovy;
    const float txtz = x / z, tytz = y / z;
    g.clampx = (txtz < -limx) || (txtz > limx);
    g.clampy = (tytz < -limy) || (tytz > limy);
    const float cx = fminf(limx, fmaxf(-limx, txtz)) * z;
    const float cy = fminf(limy, fmaxf(-limy, tytz)) * z;
    g.tc[0] = cx; g.tc[1] = cy; g.tc[2] = z;
    g.J[0][0] = fx / z; g.J[0][2] = -(fx * cx) / (z * z);
    g.J[1][1] = fy / z; g.J[1][2] = -(fy * cy) / (z * z);
  } else {
    const float su = -(float)v.image_width / (2.f * PI_F), sv = -(float)v.image_height / PI_F;
    const float rho = sqrtf(x * x + z * z);
    const float r = sqrtf(x * x + y * y + z * z);
    const float rmin = v.pole_eps * r;
    float xc = x, zc = z;
    if (rho < rmin) {
      g.clampx = true;
      if (rho > 0.f) { const float s = rmin / rho; xc = x * s; zc = z * s; }
      else { xc = 0.f; zc = rmin; }
    }
    g.tc[0] = xc; g.tc[1] = y; g.tc[2] = zc;
    const float q = xc * xc + zc * zc, rc = sqrtf(q), r2 = q + y * y;
    g.J[0][0] = su * zc / q;              g.J[0][2] = -su * xc / q;
    g.J[1][0] = -sv * xc * y / (rc * r2); g.J[1][1] = sv * rc / r2; g.J[1][2] = -sv * zc * y / (rc * r2);
  }
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int k = 0; k < 3; k++)
      g.Mm[i][k] = g.J[i][0] * V[4 * k + 0] + g.J[i][1] * V[4 * k + 1] + g.J[i][2] * V[4 * k + 2];
  const float S[3][3] = {{cov[0], cov[1], cov[2]}, {cov[1], cov[3], cov[4]}, {cov[2], cov[4], cov[5]}};
  float Sm0[3], Sm1[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    Sm0[k] = S[k][0] * g.Mm[0][0] + S[k][1] * g.Mm[0][1] + S[k][2] * g.Mm[0][2];
    Sm1[k] = S[k][0] * g.Mm[1][0] + S[k][1] * g.Mm[1][1] + S[k][2] * g.Mm[1][2];
  }
  g.a = g.Mm[0][0] * Sm0[0] + g.Mm[0][1] * Sm0[1] + g.Mm[0][2] * Sm0[2] + v.lowpass;
  g.b = g.Mm[0][0] * Sm1[0] + g.Mm[0][1] * Sm1[1] + g.Mm[0][2] * Sm1[2];
  g.c = g.Mm[1][0] * Sm1[0] + g.Mm[1][1] * Sm1[1] + g.Mm[1][2] * Sm1[2] + v.lowpass;
}

// ---------------------------------------------------------------------------------------------
// real spherical harmonics, 3DGS sign convention, bands 0..4
constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
// band 2..4 constants as host + device lookups (a namespace-scope constexpr ARRAY is not visible to device code)
S360_HD constexpr float sh_c2(int i) {
  constexpr float t[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                          0.5462742152960396f};
  return t[i];
}
S360_HD constexpr float sh_c3(int i) {
  constexpr float t[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                          -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
  return t[i];
}
S360_HD constexpr float sh_c4(int i) {
  constexpr float t[9] = {2.5033429417967046f, -1.7701307697799304f, 0.9461746957575601f, -0.6690465435572892f,
                          0.10578554691520431f, -0.6690465435572892f, 0.47308734787878004f, -1.7701307697799304f,
                          0.6258357354491761f};
  return t[i];
}

// ONE list of the 25 basis polynomials and their partial derivatives w.r.t. (x, y, z) taken as independent variables:
// f(k, b_k, db_k/dx, db_k/dy, db_k/dz) is called for every active coefficient, in order; returns their number.  Everything
// below is a thin wrapper (unused values are dead code to the compiler), so no 25-float array needs to be live anywhere.
template <class F>
S360_HD int sh_terms(int deg, float x, float y, float z, F&& f) {
  f(0, SH_C0, 0.f, 0.f, 0.f);
  if (deg < 1) return 1;
  f(1, -SH_C1 * y, 0.f, -SH_C1, 0.f);
  f(2, SH_C1 * z, 0.f, 0.f, SH_C1);
  f(3, -SH_C1 * x, -SH_C1, 0.f, 0.f);
  if (deg < 2) return 4;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  f(4, sh_c2(0) * xy, sh_c2(0) * y, sh_c2(0) * x, 0.f);
  f(5, sh_c2(1) * yz, 0.f, sh_c2(1) * z, sh_c2(1) * y);
  f(6, sh_c2(2) * (2.f * zz - xx - yy), sh_c2(2) * -2.f * x, sh_c2(2) * -2.f * y, sh_c2(2) * 4.f * z);
  f(7, sh_c2(3) * xz, sh_c2(3) * z, 0.f, sh_c2(3) * x);
  f(8, sh_c2(4) * (xx - yy), sh_c2(4) * 2.f * x, sh_c2(4) * -2.f * y, 0.f);
  if (deg < 3) return 9;
  f(9, sh_c3(0) * y * (3.f * xx - yy), sh_c3(0) * 6.f * xy, sh_c3(0) * (3.f * xx - 3.f * yy), 0.f);
  f(10, sh_c3(1) * xy * z, sh_c3(1) * yz, sh_c3(1) * xz, sh_c3(1) * xy);
  f(11, sh_c3(2) * y * (4.f * zz - xx - yy), sh_c3(2) * -2.f * xy, sh_c3(2) * (4.f * zz - xx - 3.f * yy), sh_c3(2) * 8.f * yz);
  f(12, sh_c3(3) * z * (2.f * zz - 3.f * xx - 3.f * yy), sh_c3(3) * -6.f * xz, sh_c3(3) * -6.f * yz,
    sh_c3(3) * (6.f * zz - 3.f * xx - 3.f * yy));
  f(13, sh_c3(4) * x * (4.f * zz - xx - yy), sh_c3(4) * (4.f * zz - 3.f * xx - yy), sh_c3(4) * -2.f * xy, sh_c3(4) * 8.f * xz);
  f(14, sh_c3(5) * z * (xx - yy), sh_c3(5) * 2.f * xz, sh_c3(5) * -2.f * yz, sh_c3(5) * (xx - yy));
  f(15, sh_c3(6) * x * (xx - 3.f * yy), sh_c3(6) * (3.f * xx - 3.f * yy), sh_c3(6) * -6.f * xy, 0.f);
  if (deg < 4) return 16;
  f(16, sh_c4(0) * xy * (xx - yy), sh_c4(0) * (3.f * xx * y - yy * y), sh_c4(0) * (xx * x - 3.f * x * yy), 0.f);
  f(17, sh_c4(1) * yz * (3.f * xx - yy), sh_c4(1) * 6.f * xy * z, sh_c4(1) * z * (3.f * xx - 3.f * yy), sh_c4(1) * y * (3.f * xx - yy));
  f(18, sh_c4(2) * xy * (7.f * zz - 1.f), sh_c4(2) * y * (7.f * zz - 1.f), sh_c4(2) * x * (7.f * zz - 1.f), sh_c4(2) * 14.f * xy * z);
  f(19, sh_c4(3) * yz * (7.f * zz - 3.f), 0.f, sh_c4(3) * z * (7.f * zz - 3.f), sh_c4(3) * y * (21.f * zz - 3.f));
  f(20, sh_c4(4) * (zz * (35.f * zz - 30.f) + 3.f), 0.f, 0.f, sh_c4(4) * (140.f * zz * z - 60.f * z));
  f(21, sh_c4(5) * xz * (7.f * zz - 3.f), sh_c4(5) * z * (7.f * zz - 3.f), 0.f, sh_c4(5) * x * (21.f * zz - 3.f));
  f(22, sh_c4(6) * (xx - yy) * (7.f * zz - 1.f), sh_c4(6) * 2.f * x * (7.f * zz - 1.f), sh_c4(6) * -2.f * y * (7.f * zz - 1.f),
    sh_c4(6) * (xx - yy) * 14.f * z);
  f(23, sh_c4(7) * xz * (xx - 3.f * yy), sh_c4(7) * z * (3.f * xx - 3.f * yy), sh_c4(7) * -6.f * xy * z, sh_c4(7) * x * (xx - 3.f * yy));
  f(24, sh_c4(8) * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy)), sh_c4(8) * (4.f * xx * x - 12.f * x * yy),
    sh_c4(8) * (-12.f * xx * y + 4.f * yy * y), 0.f);
  return 25;
}

// basis values for all 25 slots (unused bands are left untouched); returns number of active coeffs
S360_HD int sh_basis(int deg, float x, float y, float z, float* b) {
  return sh_terms(deg, x, y, z, [&](int k, float v, float, float, float) { b[k] = v; });
}

// partial derivatives of the basis polynomials w.r.t. (x, y, z) taken as independent variables
S360_HD void sh_basis_grad(int deg, float x, float y, float z, float* bx, float* by, float* bz) {
  sh_terms(deg, x, y, z, [&](int k, float, float gx, float gy, float gz) { bx[k] = gx; by[k] = gy; bz[k] = gz; });
}

// f(k, b_k) for every active basis function, one at a time; returns the number of coefficients
template <class F>
S360_HD int sh_basis_each(int deg, float x, float y, float z, F&& f) {
  return sh_terms(deg, x, y, z, [&](int k, float v, float, float, float) { f(k, v); });
}

// d += sum_k grad b_k(x, y, z) * s(k): the direction gradient of an SH colour; s(k) is a callable
template <class S>
S360_HD void sh_grad_dot(int deg, float x, float y, float z, S&& s, float* d) {
  sh_terms(deg, x, y, z, [&](int k, float, float gx, float gy, float gz) {
    if (k == 0) return;
    const float sk = s(k);
    d[0] += gx * sk; d[1] += gy * sk; d[2] += gz * sk; });
}

// colour sums and their Jacobian in ONE pass over the coefficients: rgb[c] = sum_k b_k sh(k, c) (no +0.5, no clamp) and
// J[c][a] = sum_k (d b_k / d a) sh(k, c), the derivative w.r.t. the (unnormalised) view direction.  sh(k, c) is a callable and
// is called once per (k, c).
template <class S>
S360_HD void sh_colour_and_jacobian(int deg, float x, float y, float z, S&& sh, float* rgb /* [3] */, float* J /* [3][3] */) {
#pragma unroll
  for (int i = 0; i < 3; i++) rgb[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 9; i++) J[i] = 0.f;
  sh_terms(deg, x, y, z, [&](int k, float v, float gx, float gy, float gz) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float s = sh(k, c);
      rgb[c] += v * s;
      J[3 * c] += gx * s; J[3 * c + 1] += gy * s; J[3 * c + 2] += gz * s;
    }
  });
}

// ---------------------------------------------------------------------------------------------
// host-side launchers (defined in the .cu files, called from api.cu)
int launch_preprocess(const S360View& v, const float* means, const float* cov, const float* opac,
                      const float* shs, const float* colors, GeomState g, int32_t* radii,
                      uint32_t* depth_keys, uint32_t* ids, S360Counters* counters, cudaStream_t st);
int launch_preprocess_backward(const S360View& v, const float* means, const float* cov, const float* opac, const float* shs,
                               GeomState g, const int32_t* radii, const float* acc, float* d_means,
                               float* d_means2D, float* d_cov, float* d_opac, float* d_shs, float* d_colors,
                               int has_depth, int depth_mode, float depth_near, float depth_far, cudaStream_t st);
int launch_mark_visible(const S360View& v, const float* means, uint8_t* present, cudaStream_t st);
// batched multi-view K1 / K8+K9 (v.viewmatrix / projmatrix / campos point at NV consecutive cameras)
int launch_preprocess_multi(const S360View& v, int NV, int64_t pair_capacity, const float* means, const float* cov,
                            const float* opac, const float* shs, const float* colors, GeomState g, PairState ps,
                            int32_t* radii, uint32_t* depth_keys, uint32_t* ids, S360Counters* counters,
                            uint32_t* status, cudaStream_t st);
int launch_zero_acc(float* acc, const uint32_t* n_dev, int64_t cap, cudaStream_t st);
int launch_preprocess_multi_backward(const S360View& v, int NV, const float* means, const float* cov, const float* opac,
                                     const float* shs, GeomState g, PairState ps, const float* acc, float* d_means,
                                     float* d_cov, float* d_opac, float* d_shs, float* d_colors, int has_depth,
                                     int depth_mode, float depth_near, float depth_far, cudaStream_t st);

// onesweep radix sort of (u32 key, u32 value) pairs on bits [0, nbits).  keys_a/vals_a hold the input;
// *_b are same-sized alternates; *result_in_b says where the result landed.  n_dev (device u32, may be NULL)
// overrides n with min(*n_dev, n).  With hist_ready the caller has zeroed the scratch with
// radix_prepare_hist() and filled the returned per-pass digit histograms itself.
size_t radix_scratch_bytes(int64_t n);
uint32_t* radix_prepare_hist(void* scratch, int64_t n, int nbits, cudaStream_t st);
int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int64_t n,
                     const uint32_t* n_dev, int nbits, void* scratch, cudaStream_t st, int* result_in_b,
                     bool hist_ready, uint32_t* vals_final = nullptr);

// n_items: entries of the geometry state the stage walks (P Gaussians, or the pair capacity of the batched path);
// n_dev (device u32, may be NULL) overrides it with min(*n_dev, n_items); NV: views stacked on the virtual image.
int launch_scan_offsets(int64_t n_items, const uint32_t* n_dev, GeomState g, const uint32_t* depth_order,
                        uint32_t* offsets, S360Counters* counters, uint32_t* block_sums, cudaStream_t st);
int launch_emit(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, GeomState g,
                const uint32_t* depth_order, const uint32_t* offsets, S360Counters* counters, int64_t capacity,
                uint32_t* keys, uint32_t* vals, uint32_t* tile_count, cudaStream_t st);
int tile_hist_copies();
int launch_tile_scan(const S360View& v, int NV, const uint32_t* tile_count, int copies, int64_t capacity, uint2* ranges,
                     uint32_t* order, uint32_t* work, uint32_t* hist, int npasses, cudaStream_t st);
// matrix binning (binning.cu): depth order -> point_list sorted by (tile, depth, id) + tile ranges + schedule, without
// materialising (tile, id) keys; usable when matrix_binning_ok(items, tiles)
bool matrix_binning_ok(int64_t n_items, int64_t tiles, int gx);
size_t matrix_scratch_bytes(int64_t n_items, int64_t tiles);
uint32_t* mb_tile_totals(void* scratch, int64_t n_items, int ntiles);
int launch_mb_count(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, GeomState g,
                    const uint32_t* depth_order, void* scratch, cudaStream_t st);
int launch_mb_colscan(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, void* scratch, cudaStream_t st);
int launch_mb_scatter(const S360View& v, int NV, int64_t n_items, const uint32_t* n_dev, GeomState g,
                      const uint32_t* depth_order, S360Counters* counters, int64_t capacity, uint32_t* point_list,
                      const uint2* ranges, void* scratch, cudaStream_t st);
int launch_tile_order(const S360View& v, int NV, const uint32_t* work, uint32_t* order, cudaStream_t st);

// out_color [NV,3,H,W], out_depth [NV,H,W], dL_dcolor [NV,3,H,W]
int launch_render_forward(const S360View& v, int NV, GeomState g, const uint32_t* point_list, ImageState img,
                          float* out_color, float* out_depth, int depth_mode, float depth_near, float depth_far,
                          cudaStream_t st);
int launch_render_backward(const S360View& v, int NV, GeomState g, const uint32_t* point_list, ImageState img,
                           const float* dL_dcolor, const float* dL_ddepth /* [NV,H,W] or NULL */, int depth_mode,
                           float depth_near, float depth_far, float* acc, cudaStream_t st);

int read_counters(unsigned long long* out, int reset, cudaStream_t st);   // render.cu, instrumented builds only

constexpr int ACC_STRIDE = 12;  // floats per Gaussian in the backward accumulator (9 used; a tenth for the depth channel)

// Build switch (prepared, NOT yet measured -- DESIGN.md sec. 7): 1 = the render backward accumulates the moments of
// q' = (o G) dL/dalpha, i.e. of the unclamped alpha it has already evaluated, instead of q = G dL/dalpha, which saves the
// second pair of ex2 per survivor; the per-Gaussian backward then drops its opacity factors and divides dL/do by o.
#ifndef S360_BWD_QPRIME
#define S360_BWD_QPRIME 0
#endif

}  // namespace s360

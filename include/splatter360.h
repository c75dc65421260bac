/*
 * splatter360.h -- C-ABI of libsplatter360.so, the B200-native differentiable Gaussian-splat
 * rasterizer (pinhole + native equirectangular).
 *
 * This is the drop-in boundary for the hot path of thucz/splatter360.  The reference reaches its
 * rasterizer through a pip-installed torch C++ extension (pybind, no C header):
 *
 *   from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
 *       /root/reference/src/model/decoder/cuda_splatting.py:5-8
 *   rasterizer(means3D=..., means2D=..., shs=..., colors_precomp=..., opacities=..., cov3D_precomp=...)
 *       /root/reference/src/model/decoder/cuda_splatting.py:113-124, 192-217
 *
 * whose extension entry points are  _C.rasterize_gaussians / _C.rasterize_gaussians_backward /
 * _C.mark_visible  (upstream is not vendored in the reference).  The
 * functions below are what a binding for that path binds instead; each names the upstream entry
 * it replaces.  INTEGRATION.md shows the ctypes stub and the reference-side import switch.
 *
 * Contract
 *   - plain C types only; every pointer is a DEVICE pointer unless it says "host".
 *   - the library never allocates, frees or keeps state: all buffers (inputs, outputs, saved state,
 *     scratch) are caller-owned; sizes come from the s360_*_bytes() queries.
 *   - every call is asynchronous and ordered on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 on success, a positive cudaError_t, or a negative S360_ERR_* code.
 *     No C++ exceptions cross the boundary.  s360_error_string() maps codes to text.
 *   - re-entrant; safe from several host threads on distinct streams.
 *
 * Data flow of one forward call (one view):
 *   s360_forward_preprocess   K1 project/cull/cov2D/SH  -> geom state, radii, depth-ordered ids,
 *                             per-Gaussian instance offsets, *num_rendered (device u32)
 *   (caller sizes the instance buffers: read *num_rendered, or pass a capacity it trusts)
 *   s360_forward_render       stable bucketing of the depth-ordered instances by tile (count matrix over
 *                             (chunk of the depth order, tile) -> column scan -> ranked scatter; or, for very
 *                             large tile counts, emit + stable radix sort) -> tile ranges -> front-to-back compositing
 *   s360_backward             per-tile back-to-front gradient pass + fused per-Gaussian backward
 *                             (optionally also the gradient of the fused depth channel)
 *
 * Batched form (s360_multi_*): V views of the same Gaussians through the same stages ONCE -- the reference's loop
 * over views / cube faces as a single pass; and s360_cube2equirec_* stitches six faces into the panorama.
 */
#ifndef SPLATTER360_H
#define SPLATTER360_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S360_ABI_VERSION 6

#define S360_MODE_PINHOLE 0 /* upstream semantics                     */
#define S360_MODE_ERP 1     /* native equirectangular splatting      */

/* fused depth channel of s360_forward_render: per-Gaussian value blended like a colour
 * (DepthRenderingMode of /root/reference/src/model/decoder/cuda_splatting.py:223-269) */
#define S360_DEPTH_DEPTH 0               /* camera z (pinhole) / radial distance (erp), unscaled */
#define S360_DEPTH_DISPARITY 1           /* 1 / depth                                            */
#define S360_DEPTH_RELATIVE_DISPARITY 2  /* depth_to_relative_disparity(depth, near, far)        */
#define S360_DEPTH_LOG 3                 /* log(max(min(depth, near), far)) as written upstream  */

#define S360_ERR_BAD_ARGUMENT (-1)
#define S360_ERR_WORKSPACE_OVERFLOW (-2) /* reported through S360Counters.overflow, see below */
#define S360_ERR_UNSUPPORTED (-3)

/* Mirrors GaussianRasterizationSettings (12 fields; cuda_splatting.py:99-112) plus the constants
 * upstream hard-codes, so the oracle and the kernels can be frozen to whatever the installed fork
 * does. */
typedef struct S360View {
  int32_t P;             /* number of Gaussians                                              */
  int32_t M;             /* SH coefficients per channel stored per Gaussian (25 in splatter360) */
  int32_t sh_degree;     /* settings.sh_degree                                               */
  int32_t image_height;  /* settings.image_height                                            */
  int32_t image_width;   /* settings.image_width                                             */
  int32_t mode;          /* S360_MODE_*                                                      */
  int32_t max_sh_degree; /* highest SH band evaluated (4; stock upstream stops at 3)         */
  int32_t tight_bbox;    /* 1: drop tiles the alpha>=1/255 ellipse cannot reach (image-exact) */
  float tanfovx;         /* settings.tanfovx                                                 */
  float tanfovy;         /* settings.tanfovy                                                 */
  float near_cull;       /* 0.2 : cull view z (pinhole) / radial distance (erp) <= this      */
  float fov_clamp;       /* 1.3 : |x/z| clamp factor inside J (pinhole)                      */
  float lowpass;         /* 0.3 : added to the cov2D diagonal                                */
  float pole_eps;        /* erp : horizontal radius clamp rho >= pole_eps * r                */
  float scene_scale;     /* means *= s, cov3D *= s^2 on load (the reference's 1/near rescale,
                            cuda_splatting.py:64-71); gradients are returned w.r.t. the unscaled inputs */
  int32_t sh_layout;     /* 0: shs[P,M,3] (rasterizer API)   1: [P,3,M] (reference's harmonics layout) */
  int32_t cov_layout;    /* 0: cov3D[P,6]                     1: [P,3,3] (upper triangle read / written) */
  int32_t reserved0;
  const float* viewmatrix; /* [16] settings.viewmatrix  (p_view = [x y z 1] . V, row-major)  */
  const float* projmatrix; /* [16] settings.projmatrix  (unused in erp mode)                 */
  const float* campos;     /* [3]  settings.campos                                           */
  const float* bg;         /* [3]  settings.bg                                               */
} S360View;

/* Device-side counters written by the forward pass (16 bytes). */
typedef struct S360Counters {
  uint32_t num_rendered; /* total (tile, Gaussian) instances N this view needs               */
  uint32_t overflow;     /* bit 0: s360_forward_render needed N > instance_capacity;
                            bit 1: the batched path needed more pairs than pair_capacity      */
  uint32_t num_visible;  /* Gaussians with radius > 0; batched path: (view, Gaussian) pairs needed */
  uint32_t reserved;
} S360Counters;

/* ---- sizes of caller-owned buffers -------------------------------------------------------- */
/* geometry state kept from forward to backward (per Gaussian). */
size_t s360_geom_bytes(int32_t P);
/* transient scratch of s360_forward_preprocess (depth sort + scan). */
size_t s360_preprocess_scratch_bytes(int32_t P);
/* transient scratch of s360_forward_render for P Gaussians and `instance_capacity` instances. */
size_t s360_binning_scratch_bytes(int32_t P, int64_t instance_capacity, int32_t image_height, int32_t image_width);
/* image state kept from forward to backward (final transmittance, contributor count, tile ranges). */
size_t s360_image_bytes(int32_t image_height, int32_t image_width);
/* transient scratch of s360_backward (per-Gaussian screen-space gradient accumulators). */
size_t s360_backward_scratch_bytes(int32_t P);

/* ---- forward, stage 1 (replaces the first half of upstream _C.rasterize_gaussians:
 *      preprocessCUDA + InclusiveSum) ------------------------------------------------------- */
int s360_forward_preprocess(
    const S360View* view,        /* host */
    const float* means3D,        /* [P,3]                                                      */
    const float* cov3D,          /* [P,6] (xx,xy,xz,yy,yz,zz)  cov3D_precomp                   */
    const float* opacities,      /* [P]                                                        */
    const float* shs,            /* [P,M,3] or NULL                                            */
    const float* colors_precomp, /* [P,3]  or NULL (exactly one of shs/colors_precomp)         */
    void* geom,                  /* s360_geom_bytes(P)                                         */
    int32_t* radii,              /* [P] out                                                    */
    uint32_t* depth_order,       /* [P] out: Gaussian ids sorted by (depth, id)                */
    uint32_t* inst_offsets,      /* [P] scratch handed on to s360_forward_render (exclusive scan of tiles touched in
                                    depth order when the emit + radix-sort binning path is taken; untouched by the
                                    matrix binning path, which needs no per-Gaussian offsets)                 */
    S360Counters* counters,      /* out (device)                                               */
    void* scratch,               /* s360_preprocess_scratch_bytes(P)                           */
    void* stream);

/* The same stage in two halves, so that a host that needs the instance count can read
 * counters->num_rendered (final after s360_forward_project) while s360_forward_order is still running:
 *   s360_forward_project : K1 only -> geom, radii, counters (scratch keeps the unsorted depth keys)
 *   s360_forward_order   : depth sort + scan -> depth_order, inst_offsets (same scratch, same stream) */
int s360_forward_project(const S360View* view, const float* means3D, const float* cov3D, const float* opacities,
                         const float* shs, const float* colors_precomp, void* geom, int32_t* radii,
                         S360Counters* counters, void* scratch, void* stream);
int s360_forward_order(const S360View* view, const void* geom, uint32_t* depth_order, uint32_t* inst_offsets,
                       S360Counters* counters, void* scratch, void* stream);

/* ---- forward, stage 2 (replaces duplicateWithKeys + SortPairs + identifyTileRanges + renderCUDA) */
int s360_forward_render(
    const S360View* view,        /* host */
    const void* geom,
    const uint32_t* depth_order, /* from stage 1 */
    const uint32_t* inst_offsets,/* from stage 1 */
    S360Counters* counters,      /* device; num_rendered read, overflow written                */
    int64_t instance_capacity,   /* entries available in point_list / binning scratch          */
    uint32_t* point_list,        /* [instance_capacity] out: Gaussian ids sorted by (tile, depth, id) */
    void* image_state,           /* s360_image_bytes(H,W) out                                  */
    float* out_color,            /* [3,H,W] out                                                */
    float* out_depth,            /* [H,W] out or NULL: fused depth channel (gradient: s360_backward's dL_ddepth) */
    int32_t depth_mode,          /* S360_DEPTH_*                                               */
    float depth_near, float depth_far, /* unscaled near / far for relative_disparity and log   */
    void* scratch,               /* s360_binning_scratch_bytes(P,instance_capacity,H,W)        */
    void* stream);

/* ---- backward (replaces upstream _C.rasterize_gaussians_backward) ---------------------------- */
int s360_backward(
    const S360View* view,        /* host */
    const float* means3D, const float* cov3D, const float* opacities,
    const float* shs, const float* colors_precomp,
    const void* geom, const int32_t* radii,
    const uint32_t* point_list, const void* image_state,
    const float* dL_dcolor,      /* [3,H,W]                                                    */
    const float* dL_ddepth,      /* [H,W] gradient w.r.t. the fused depth channel, or NULL     */
    int32_t depth_mode, float depth_near, float depth_far, /* as given to s360_forward_render  */
    float* dL_dmeans3D,          /* [P,3] out                                                  */
    float* dL_dmeans2D,          /* [P,3] out (NDC units, z = 0)  */
    float* dL_dcov3D,            /* [P,6] out                                                  */
    float* dL_dopacity,          /* [P]   out                                                  */
    float* dL_dshs,              /* [P,M,3] out or NULL                                        */
    float* dL_dcolors,           /* [P,3] out or NULL                                          */
    void* scratch,               /* s360_backward_scratch_bytes(P)                             */
    void* stream);

/* ---- batched multi-view path ------------------------------------------
 * Renders V views of the SAME Gaussians in one pass.  The reference loops the views and calls the rasterizer once
 * per view and batch item (/root/reference/src/model/decoder/decoder_splatting_cuda.py:44-59,
 * cuda_splatting.py:91-126) -- six 90-degree cube faces per panorama (model_wrapper_erp.py:336-345) -- so every
 * call re-reads all P Gaussians.  Here each Gaussian is read once and projected into all V views; every
 * (view, Gaussian) pair that touches a tile gets a slot in pair-indexed geometry buffers (in Gaussian-major order,
 * so equal depths keep index order exactly as V separate calls would), and the sort / emission / compositing /
 * backward stages run ONCE over the pairs on a virtual image of V stacked views.  Results are those of V
 * separate s360_forward_* calls; gradients are the sum over the views (what autograd produces in the reference).
 *
 * `view` carries the settings common to all views (P, M, image size, mode, tanfov, scene_scale, layouts, bg);
 * its viewmatrix / projmatrix / campos point at V consecutive [16] / [16] / [3] blocks.  1 <= V <= S360_MAX_VIEWS.
 * pair_capacity: slots in the pair buffers; V * P can never overflow, a tighter value saves memory and is
 * reported through counters->overflow bit 1 / counters->num_visible (pairs needed) when too small.
 * dL_dmeans2D is not produced by the batched path. */
#define S360_MAX_VIEWS 32
size_t s360_multi_geom_bytes(int32_t P, int64_t pair_capacity);
size_t s360_multi_preprocess_scratch_bytes(int32_t P, int64_t pair_capacity);
size_t s360_multi_binning_scratch_bytes(int64_t pair_capacity, int64_t instance_capacity, int32_t V, int32_t image_height,
                                        int32_t image_width);
size_t s360_multi_image_bytes(int32_t V, int32_t image_height, int32_t image_width);
size_t s360_multi_backward_scratch_bytes(int64_t pair_capacity);
/* K1 for all views -> pair geometry, radii [V,P] (may be NULL), counters (num_rendered final, num_visible = pairs) */
int s360_multi_forward_project(const S360View* view, int32_t V, int64_t pair_capacity, const float* means3D,
                               const float* cov3D, const float* opacities, const float* shs,
                               const float* colors_precomp, void* geom, int32_t* radii, S360Counters* counters,
                               void* scratch, void* stream);
/* depth sort + scan over the pairs -> depth_order [pair_capacity], inst_offsets [pair_capacity] */
int s360_multi_forward_order(const S360View* view, int32_t V, int64_t pair_capacity, const void* geom,
                             uint32_t* depth_order, uint32_t* inst_offsets, S360Counters* counters, void* scratch,
                             void* stream);
/* emission + tile sort + compositing of all views: out_color [V,3,H,W], out_depth [V,H,W] or NULL;
 * point_list holds pair slots sorted by (view, tile, depth, Gaussian) */
int s360_multi_forward_render(const S360View* view, int32_t V, int64_t pair_capacity, const void* geom,
                              const uint32_t* depth_order, const uint32_t* inst_offsets, S360Counters* counters,
                              int64_t instance_capacity, uint32_t* point_list, void* image_state, float* out_color,
                              float* out_depth, int32_t depth_mode, float depth_near, float depth_far, void* scratch,
                              void* stream);
/* dL_dcolor [V,3,H,W] -> gradients summed over the views */
int s360_multi_backward(const S360View* view, int32_t V, int64_t pair_capacity, const float* means3D,
                        const float* cov3D, const float* opacities, const float* shs, const float* colors_precomp,
                        const void* geom, const uint32_t* point_list, const void* image_state, const float* dL_dcolor,
                        const float* dL_ddepth /* [V,H,W] or NULL */, int32_t depth_mode, float depth_near,
                        float depth_far, float* dL_dmeans3D, float* dL_dcov3D, float* dL_dopacity, float* dL_dshs,
                        float* dL_dcolors, void* scratch, void* stream);

/* ---- camera-to-world -> view matrix: batched inverse of n row-major 4x4 matrices (in != out).  Replaces
 * `extrinsics.inverse()` of /root/reference/src/model/decoder/cuda_splatting.py:84, :176, :262 (a batched LU with a host
 * status read-back) by one launch; partial pivoting, double-precision arithmetic rounded once to float; no singularity
 * status (like torch.linalg.inv_ex). */
int s360_invert4x4(const float* in, float* out, int64_t n, void* stream);

/* ---- visibility mask (replaces upstream _C.mark_visible) ------------------------------------ */
int s360_mark_visible(const S360View* view, const float* means3D, uint8_t* present, void* stream);

/* ---- fused photometric loss: loss = weight * mean((color - target)^2), grad = d loss / d color.
 * Replaces the elementwise/reduction chain of /root/reference/src/loss/loss_mse.py:22-31 on the rendered image and
 * yields the seed gradient of s360_backward in the same pass.  n = number of floats. */
int s360_mse_loss_grad(const float* color, const float* target, int64_t n, float weight, float* loss /* [1] out */,
                       float* grad /* [n] out */, void* stream);

/* ---- cube faces -> equirectangular panorama.  Replaces change_order
 * (/root/reference/src/model/model_wrapper_erp.py:135-158) + the strip torch.cat (:395-398) + Cube2Equirec.forward
 * (/root/reference/src/geometry/layers.py:108-116, a 5-D F.grid_sample) with one gather kernel; the backward is the
 * matching scatter.  grid [H,W,3] is the reference module's sample grid (u, v in [-1,1], z = face / 2.5 - 1, faces
 * in the order [F R B L U D]); bilinear, align_corners, border clamp.
 *   layout 0: faces [B,C,f,6f]   the strip the reference module takes
 *   layout 1: faces [B,6,C,f,f]  rasterizer output in the dataset face order [U B L F R D]; the reorder and the
 *                                180-degree turn of U and D are applied by index arithmetic */
int s360_cube2equirec_forward(const float* faces, const float* grid, int32_t layout, int32_t B, int32_t C,
                              int32_t face_w, int32_t H, int32_t W,
                              const float* depth_to_distance /* host [4] = fx, fy, cx, cy in pixels, or NULL */,
                              float* out /*[B,C,H,W]*/, void* stream);
int s360_cube2equirec_backward(const float* dL_dout /*[B,C,H,W]*/, const float* grid, int32_t layout, int32_t B,
                               int32_t C, int32_t face_w, int32_t H, int32_t W, const float* depth_to_distance,
                               float* dL_dfaces /* out, same layout as faces; zeroed by the call */, void* stream);
/* depth_to_distance != NULL: the faces hold z-depth and the panorama is wanted in radial distance -- every tap is
 * scaled by sqrt(((r - cx)/fx)^2 + ((c - cy)/fy)^2 + 1) of its texel (row r, column c; the reference's meshgrid is
 * 'ij'-indexed) on the reordered face, i.e. the reference's
 * depth_to_distance_map_batch (/root/reference/src/geometry/z_depth_to_distance.py:4-34) applied between change_order
 * and the stitch exactly as its depth video does (model_wrapper_erp.py:447-463). */

/* ---- fused Gaussian adapter: encoder head outputs -> rasterizer inputs in one pass.
 * Replaces GaussianAdapterERP.forward (/root/reference/src/model/encoder/common/gaussian_adapter_erp.py:49-119): scale
 * activation, quaternion normalisation, build_covariance (gaussians.py:8-44), rotation of the covariance into the world frame,
 * ERP unprojection of the depth (sphere_projection.py:6-87, hm3d convention), SH mask and SH rotation (sh_rotation.py:10-30).
 * One Gaussian per context pixel, in (batch x view, row, col) order: G = views * H * W.
 *   raw      [G, 7 + 3 d_sh]  scale features 3 | quaternion xyzw 4 | SH [3][d_sh],  d_sh = (sh_degree + 1)^2, sh_degree <= 4
 *   depth    [G]
 *   pose     [views, 12]      camera-to-world R (row-major 3x3) and t of the context view
 *   sh_rot   [views, 165]     the SH rotation the reference applies (Wigner D^0..D^4 of R, row-major blocks of size 1, 9, 25,
 *                             49, 81), each column pre-multiplied by the reference's SH mask 0.1 * 0.25^l (1 for l = 0)
 *   outputs  means [G,3], covariances [G,3,3], harmonics [G,3,d_sh]; scales [G,3] / rotations [G,4] (normalised xyzw) may be NULL
 * Backward: cotangents (any may be NULL = zero) -> dL_draw [G, 7 + 3 d_sh], dL_ddepth [G].  means_grad = 0 reproduces the
 * reference, whose unprojection runs under torch.no_grad() (sphere_projection.py:14): the means carry no gradient. */
int s360_adapter_forward(int32_t views, int32_t H, int32_t W, int32_t sh_degree, float scale_min, float scale_max,
                         const float* raw, const float* depth, const float* pose, const float* sh_rot, float* means,
                         float* covariances, float* harmonics, float* scales, float* rotations, void* stream);
int s360_adapter_backward(int32_t views, int32_t H, int32_t W, int32_t sh_degree, float scale_min, float scale_max,
                          int32_t means_grad, const float* raw, const float* depth, const float* pose, const float* sh_rot,
                          const float* dL_dmeans, const float* dL_dcovariances, const float* dL_dharmonics, float* dL_draw,
                          float* dL_ddepth, void* stream);

/* ---- debugging / introspection ---------------------------------------------------------------- */
/* Unpack the geometry state for per-stage parity tests.  Any output may be NULL. */
int s360_debug_unpack_geom(int32_t P, const void* geom, float* xy /*[P,2]*/, float* depth /*[P]*/,
                           float* conic_opacity /*[P,4]*/, float* rgb /*[P,3]*/,
                           uint32_t* tiles_touched /*[P]*/, uint8_t* clamped /*[P,3]*/, void* stream);
/* Unpack the image state.  Any output may be NULL. */
int s360_debug_unpack_image(int32_t image_height, int32_t image_width, const void* image_state,
                            float* final_T /*[H,W]*/, uint32_t* n_contrib /*[H,W]*/,
                            uint32_t* tile_ranges /*[tiles,2]*/, void* stream);

/* Batched path: per-Gaussian pair bookkeeping (the pair of view v is slot pair_base + popcount(view_mask & ((1 << v) - 1)))
 * and the number of pairs stored.  Pair slots index the pair geometry: s360_debug_unpack_geom(pair_capacity, geom, ...)
 * unpacks it (tile rows there count from the top of the stacked image).  Any output may be NULL. */
int s360_debug_unpack_pairs(int32_t P, int64_t pair_capacity, const void* geom, uint32_t* pair_base /*[P]*/,
                            uint32_t* view_mask /*[P]*/, uint32_t* num_pairs /*[1]*/, void* stream);

/* ---- optional per-stage timing (CUDA events recorded on the caller's stream around each stage).
 * Process-wide switch, off by default; used by bench.py for the roofline numbers.  The only mutable
 * state the library holds besides the launch counter. */
#define S360_STAGE_PREPROCESS 0
#define S360_STAGE_DEPTH_SORT 1
#define S360_STAGE_SCAN 2
#define S360_STAGE_EMIT 3
#define S360_STAGE_TILE_SORT 4
#define S360_STAGE_TILE_RANGES 5
#define S360_STAGE_RENDER_FWD 6
#define S360_STAGE_RENDER_BWD 7
#define S360_STAGE_PREPROCESS_BWD 8
#define S360_NUM_STAGES 9
int s360_profile_enable(int on); /* returns the previous setting */
/* Sum of elapsed milliseconds and number of recordings per stage since the last reset.
 * Synchronises on the recorded events.  ms / counts: host arrays of S360_NUM_STAGES, may be NULL. */
int s360_profile_read(double* ms, uint64_t* counts, int reset);

/* Work counters of the render kernels (host array of 16 uint64; see render.cu).  Only instrumented builds
 * (-DS360_COUNTERS=1, tools/counters.py) count; the shipped library returns S360_ERR_UNSUPPORTED and zeros. */
int s360_debug_counters(uint64_t* out /* host [16] */, int reset, void* stream);

int s360_abi_version(void);
const char* s360_error_string(int code);
/* number of kernels launched by this library since load (for bench.py's gpu_launches). */
uint64_t s360_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SPLATTER360_H */

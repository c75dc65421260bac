/*
 * raster_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * Plain-C restatement of the tile-based differentiable Gaussian rasterizer that
 * splatter360 calls through `diff_gaussian_rasterization`
 * (/root/reference/src/model/decoder/cuda_splatting.py:5-8, 99-126; requirements.txt:17).
 *
 * PARITY UNPINNED: the arithmetic lives in an un-vendored, un-pinned pip dependency
 * (dcharatan/diff-gaussian-rasterization-modified, fork of
 * graphdeco-inria/diff-gaussian-rasterization) whose source is not under
 * /root/reference, and the reference has no tests or golden vectors for this path
 * (SURVEY.md sec. 4, 8c).  This file restates the *published* 3DGS rasterizer algorithm
 *  and is cross-checked against an independent float64 PyTorch
 * autograd restatement (oracle/torch_oracle.py).  The camera/argument construction
 * around it IS pinned against the importable reference Python (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path never does.
 *
 * Two projection modes:
 *   mode 0 "pinhole": the upstream algorithm (Appendix A).
 *   mode 1 "erp":     native equirectangular splatting (camera
 *                     convention = /root/reference/src/geometry/utils360.py:93-104,
 *                     148-153, 193-198, 250-263).  No reference implementation exists.
 *
 * Stages (named after SURVEY.md sec. 2b):  K1 preprocess -> K2..K5 binning (scan,
 * duplicate, stable sort by (tile, depth, id), tile ranges) -> K6 render ->
 * K7 render backward -> K8 cov2D backward -> K9 preprocess/SH backward.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ORACLE_F64: the SAME source with every float as double (liboracle_f64.so; arrays and OrcCfg fields are then doubles).
 * The float32 constants of the algorithm (0.3f, 1.3f, SH constants ...) keep their float32 values; the depth sort key is the
 * float32 rounding of the depth, as in every float32 implementation.  This is the "truth" the float32 paths (this oracle
 * and the CUDA kernels) are measured against where their mutual difference is of the order of float32 rounding noise. */
#ifdef ORACLE_F64
#define float double
#define sqrtf sqrt
#define fmaxf fmax
#define fminf fmin
#define ceilf ceil
#define floorf floor
#define expf exp
#define atan2f atan2
#define logf log
#endif

#define TILE 16

typedef struct {
  int32_t P;              /* number of Gaussians                                   */
  int32_t M;              /* SH coefficients per channel stored per Gaussian       */
  int32_t D;              /* active SH degree (settings.sh_degree)                 */
  int32_t H, W;           /* image size                                            */
  int32_t mode;           /* 0 pinhole, 1 erp                                      */
  int32_t max_sh_degree;  /* highest band the evaluator implements (3 stock, 4 fork?) */
  int32_t use_sh;         /* 1: colours from shs[P,M,3]; 0: colors_precomp[P,3]    */
  float tanfovx, tanfovy;
  float near_cull;        /* 0.2 upstream                                          */
  float fov_clamp;        /* 1.3 upstream                                          */
  float lowpass;          /* 0.3 upstream                                          */
  float pole_eps;         /* erp: rho >= pole_eps * r clamp                        */
  float view[16];         /* p_view = (x,y,z,1) . view   (row-major 4x4)           */
  float proj[16];         /* p_hom  = (x,y,z,1) . proj                             */
  float campos[3];
  float bg[3];
} OrcCfg;

/* ---------------------------------------------------------------- SH basis */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};
static const float SH_C4[9] = {2.5033429417967046f,  -1.7701307697799304f, 0.9461746957575601f,
                               -0.6690465435572892f, 0.10578554691520431f, -0.6690465435572892f,
                               0.47308734787878004f, -1.7701307697799304f, 0.6258357354491761f};

/* basis value b[k] and partial derivatives (as polynomials in independent x,y,z) */
static int sh_basis(int deg, float x, float y, float z, float* b, float* bx, float* by, float* bz) {
  int n = 1;
  b[0] = SH_C0; bx[0] = by[0] = bz[0] = 0.f;
  if (deg > 0) {
    b[1] = -SH_C1 * y; bx[1] = 0; by[1] = -SH_C1; bz[1] = 0;
    b[2] = SH_C1 * z;  bx[2] = 0; by[2] = 0; bz[2] = SH_C1;
    b[3] = -SH_C1 * x; bx[3] = -SH_C1; by[3] = 0; bz[3] = 0;
    n = 4;
  }
  float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  if (deg > 1) {
    b[4] = SH_C2[0] * xy; bx[4] = SH_C2[0] * y; by[4] = SH_C2[0] * x; bz[4] = 0;
    b[5] = SH_C2[1] * yz; bx[5] = 0; by[5] = SH_C2[1] * z; bz[5] = SH_C2[1] * y;
    b[6] = SH_C2[2] * (2.f * zz - xx - yy);
    bx[6] = SH_C2[2] * -2.f * x; by[6] = SH_C2[2] * -2.f * y; bz[6] = SH_C2[2] * 4.f * z;
    b[7] = SH_C2[3] * xz; bx[7] = SH_C2[3] * z; by[7] = 0; bz[7] = SH_C2[3] * x;
    b[8] = SH_C2[4] * (xx - yy); bx[8] = SH_C2[4] * 2.f * x; by[8] = SH_C2[4] * -2.f * y; bz[8] = 0;
    n = 9;
  }
  if (deg > 2) {
    b[9] = SH_C3[0] * y * (3.f * xx - yy);
    bx[9] = SH_C3[0] * 6.f * xy; by[9] = SH_C3[0] * (3.f * xx - 3.f * yy); bz[9] = 0;
    b[10] = SH_C3[1] * xy * z;
    bx[10] = SH_C3[1] * yz; by[10] = SH_C3[1] * xz; bz[10] = SH_C3[1] * xy;
    b[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
    bx[11] = SH_C3[2] * -2.f * xy; by[11] = SH_C3[2] * (4.f * zz - xx - 3.f * yy); bz[11] = SH_C3[2] * 8.f * yz;
    b[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
    bx[12] = SH_C3[3] * -6.f * xz; by[12] = SH_C3[3] * -6.f * yz; bz[12] = SH_C3[3] * (6.f * zz - 3.f * xx - 3.f * yy);
    b[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
    bx[13] = SH_C3[4] * (4.f * zz - 3.f * xx - yy); by[13] = SH_C3[4] * -2.f * xy; bz[13] = SH_C3[4] * 8.f * xz;
    b[14] = SH_C3[5] * z * (xx - yy);
    bx[14] = SH_C3[5] * 2.f * xz; by[14] = SH_C3[5] * -2.f * yz; bz[14] = SH_C3[5] * (xx - yy);
    b[15] = SH_C3[6] * x * (xx - 3.f * yy);
    bx[15] = SH_C3[6] * (3.f * xx - 3.f * yy); by[15] = SH_C3[6] * -6.f * xy; bz[15] = 0;
    n = 16;
  }
  if (deg > 3) {
    b[16] = SH_C4[0] * xy * (xx - yy);
    bx[16] = SH_C4[0] * (3.f * xx * y - yy * y); by[16] = SH_C4[0] * (xx * x - 3.f * x * yy); bz[16] = 0;
    b[17] = SH_C4[1] * yz * (3.f * xx - yy);
    bx[17] = SH_C4[1] * 6.f * xy * z; by[17] = SH_C4[1] * z * (3.f * xx - 3.f * yy); bz[17] = SH_C4[1] * y * (3.f * xx - yy);
    b[18] = SH_C4[2] * xy * (7.f * zz - 1.f);
    bx[18] = SH_C4[2] * y * (7.f * zz - 1.f); by[18] = SH_C4[2] * x * (7.f * zz - 1.f); bz[18] = SH_C4[2] * 14.f * xy * z;
    b[19] = SH_C4[3] * yz * (7.f * zz - 3.f);
    bx[19] = 0; by[19] = SH_C4[3] * z * (7.f * zz - 3.f); bz[19] = SH_C4[3] * y * (21.f * zz - 3.f);
    b[20] = SH_C4[4] * (zz * (35.f * zz - 30.f) + 3.f);
    bx[20] = 0; by[20] = 0; bz[20] = SH_C4[4] * (140.f * zz * z - 60.f * z);
    b[21] = SH_C4[5] * xz * (7.f * zz - 3.f);
    bx[21] = SH_C4[5] * z * (7.f * zz - 3.f); by[21] = 0; bz[21] = SH_C4[5] * x * (21.f * zz - 3.f);
    b[22] = SH_C4[6] * (xx - yy) * (7.f * zz - 1.f);
    bx[22] = SH_C4[6] * 2.f * x * (7.f * zz - 1.f); by[22] = SH_C4[6] * -2.f * y * (7.f * zz - 1.f);
    bz[22] = SH_C4[6] * (xx - yy) * 14.f * z;
    b[23] = SH_C4[7] * xz * (xx - 3.f * yy);
    bx[23] = SH_C4[7] * z * (3.f * xx - 3.f * yy); by[23] = SH_C4[7] * -6.f * xy * z; bz[23] = SH_C4[7] * x * (xx - 3.f * yy);
    b[24] = SH_C4[8] * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
    bx[24] = SH_C4[8] * (4.f * xx * x - 12.f * x * yy); by[24] = SH_C4[8] * (-12.f * xx * y + 4.f * yy * y); bz[24] = 0;
    n = 25;
  }
  return n;
}

/* --------------------------------------------------------- per-Gaussian geometry */
typedef struct {
  float t[3];      /* view-space centre (before any clamp)                  */
  float tc[3];     /* centre used inside J (after fov / pole clamp)         */
  int   clampx, clampy; /* pinhole: fov clamp active; erp: clampx = pole clamp  */
  float J[2][3];   /* screen Jacobian (pixel units)                         */
  float Mm[2][3];  /* J * R                                                 */
  float a, b, c;   /* cov2D incl. low-pass                                  */
} Geo;

static void view_point(const float* V, const float* p, float* t) {
  t[0] = V[0] * p[0] + V[4] * p[1] + V[8] * p[2] + V[12];
  t[1] = V[1] * p[0] + V[5] * p[1] + V[9] * p[2] + V[13];
  t[2] = V[2] * p[0] + V[6] * p[1] + V[10] * p[2] + V[14];
}

/* R[i][k]: view_i = sum_k R[i][k] * world_k */
static void view_rot(const float* V, float R[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < 3; k++) R[i][k] = V[4 * k + i];
}

static void geo_compute(const OrcCfg* c, const float* mean, const float* cov6, Geo* g) {
  float R[3][3];
  view_rot(c->view, R);
  view_point(c->view, mean, g->t);
  float x = g->t[0], y = g->t[1], z = g->t[2];
  g->clampx = g->clampy = 0;
  memset(g->J, 0, sizeof(g->J));
  if (c->mode == 0) {
    /* Appendix A, K1: t.x = clamp(t.x/t.z, +-1.3 tanfovx) * t.z */
    float fx = (float)c->W / (2.f * c->tanfovx), fy = (float)c->H / (2.f * c->tanfovy);
    float limx = c->fov_clamp * c->tanfovx, limy = c->fov_clamp * c->tanfovy;
    float txtz = x / z, tytz = y / z;
    if (txtz < -limx || txtz > limx) g->clampx = 1;
    if (tytz < -limy || tytz > limy) g->clampy = 1;
    float cx = fminf(limx, fmaxf(-limx, txtz)) * z;
    float cy = fminf(limy, fmaxf(-limy, tytz)) * z;
    g->tc[0] = cx; g->tc[1] = cy; g->tc[2] = z;
    g->J[0][0] = fx / z; g->J[0][2] = -(fx * cx) / (z * z);
    g->J[1][1] = fy / z; g->J[1][2] = -(fy * cy) / (z * z);
  } else {
    /* Appendix B2 */
    const float PI = 3.14159265358979323846f;
    float su = -(float)c->W / (2.f * PI), sv = -(float)c->H / PI;
    float rho = sqrtf(x * x + z * z);
    float r = sqrtf(x * x + y * y + z * z);
    float rmin = c->pole_eps * r;
    float xc = x, zc = z;
    if (rho < rmin) {
      g->clampx = 1;
      if (rho > 0.f) { xc = x * (rmin / rho); zc = z * (rmin / rho); }
      else { xc = 0.f; zc = rmin; }
    }
    g->tc[0] = xc; g->tc[1] = y; g->tc[2] = zc;
    float q = xc * xc + zc * zc, rc = sqrtf(q), r2 = q + y * y;
    g->J[0][0] = su * zc / q;              g->J[0][2] = -su * xc / q;
    g->J[1][0] = -sv * xc * y / (rc * r2); g->J[1][1] = sv * rc / r2; g->J[1][2] = -sv * zc * y / (rc * r2);
  }
  for (int i = 0; i < 2; i++)
    for (int k = 0; k < 3; k++)
      g->Mm[i][k] = g->J[i][0] * R[0][k] + g->J[i][1] * R[1][k] + g->J[i][2] * R[2][k];
  float S[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
  float Sm0[3], Sm1[3];
  for (int k = 0; k < 3; k++) {
    Sm0[k] = S[k][0] * g->Mm[0][0] + S[k][1] * g->Mm[0][1] + S[k][2] * g->Mm[0][2];
    Sm1[k] = S[k][0] * g->Mm[1][0] + S[k][1] * g->Mm[1][1] + S[k][2] * g->Mm[1][2];
  }
  g->a = g->Mm[0][0] * Sm0[0] + g->Mm[0][1] * Sm0[1] + g->Mm[0][2] * Sm0[2] + c->lowpass;
  g->b = g->Mm[0][0] * Sm1[0] + g->Mm[0][1] * Sm1[1] + g->Mm[0][2] * Sm1[2];
  g->c = g->Mm[1][0] * Sm1[0] + g->Mm[1][1] * Sm1[1] + g->Mm[1][2] * Sm1[2] + c->lowpass;
}

/* tile rectangle of one Gaussian.  x range may be "unwrapped" (erp): tiles x0..x0+nx-1 mod gx */
typedef struct { int x0, nx, y0, ny; } Rect;

static Rect get_rect(const OrcCfg* c, float px, float py, int ex, int ey, int gx, int gy) {
  Rect r;
  int ymin = (int)((py - ey) / TILE), ymax = (int)((py + ey + TILE - 1) / TILE);
  ymin = ymin < 0 ? 0 : (ymin > gy ? gy : ymin);
  ymax = ymax < 0 ? 0 : (ymax > gy ? gy : ymax);
  r.y0 = ymin; r.ny = ymax - ymin;
  if (c->mode == 0) {
    int xmin = (int)((px - ex) / TILE), xmax = (int)((px + ex + TILE - 1) / TILE);
    xmin = xmin < 0 ? 0 : (xmin > gx ? gx : xmin);
    xmax = xmax < 0 ? 0 : (xmax > gx ? gx : xmax);
    r.x0 = xmin; r.nx = xmax - xmin;
  } else {
    int xmin = (int)floorf((px - ex) / TILE), xmax = (int)floorf((px + ex + TILE - 1) / TILE);
    int nx = xmax - xmin;
    if (nx > gx) nx = gx;
    if (nx < 0) nx = 0;
    r.x0 = xmin; r.nx = nx;
  }
  if (r.ny < 0) r.ny = 0;
  return r;
}

static inline int wrap_tile(int t, int gx) { int m = t % gx; return m < 0 ? m + gx : m; }

/* ------------------------------------------------------------ stable radix sort */
static void sort_pairs_u64(uint64_t* keys, uint32_t* vals, size_t n, int nbits) {
  uint64_t* k2 = (uint64_t*)malloc(n * sizeof(uint64_t));
  uint32_t* v2 = (uint32_t*)malloc(n * sizeof(uint32_t));
  for (int shift = 0; shift < nbits; shift += 8) {
    size_t cnt[257];
    memset(cnt, 0, sizeof(cnt));
    for (size_t i = 0; i < n; i++) cnt[((keys[i] >> shift) & 255) + 1]++;
    for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
    for (size_t i = 0; i < n; i++) {
      size_t dst = cnt[(keys[i] >> shift) & 255]++;
      k2[dst] = keys[i]; v2[dst] = vals[i];
    }
    uint64_t* tk = keys; keys = k2; k2 = tk;
    uint32_t* tv = vals; vals = v2; v2 = tv;
  }
  /* after an even number of passes data is back in the caller's arrays */
  int passes = (nbits + 7) / 8;
  if (passes & 1) {
    memcpy(k2, keys, n * sizeof(uint64_t));
    memcpy(v2, vals, n * sizeof(uint32_t));
    uint64_t* tk = keys; keys = k2; k2 = tk;
    uint32_t* tv = vals; vals = v2; v2 = tv;
  }
  free(k2); free(v2);
}

/* =====================================================================
 * oracle_render: forward (+ optional backward) of one view.
 *
 * Inputs : means[P,3], cov6[P,6] (xx,xy,xz,yy,yz,zz), opac[P], shs[P,M,3] or colors[P,3].
 * Outputs (any may be NULL unless noted):
 *   out_color[3,H,W] (required), out_radii[P], out_final_T[H,W], out_n_contrib[H,W],
 *   g_xy[P,2], g_depth[P], g_conic_op[P,4], g_rgb[P,3], g_tiles[P], g_clamped[P,3] (uint8),
 *   inst_tile[cap], inst_gid[cap] (sorted instance list), tile_ranges[tiles,2], num_rendered.
 * Backward (if dL_dpix != NULL): dL_dpix[3,H,W] ->
 *   d_means[P,3], d_means2D[P,3], d_cov6[P,6], d_opac[P], d_shs[P,M,3], d_colors[P,3].
 * Returns 0, or -1 on allocation failure, -2 if inst capacity too small.
 * ===================================================================== */
/* Depth channel (what the reference renders in a second rasterisation with depth as colour,
 * /root/reference/src/model/decoder/cuda_splatting.py:226-269): per-Gaussian value of the sort depth z (camera z in
 * pinhole mode, radial distance in erp mode) divided by `scale`:
 *   0 depth: z   1 disparity: 1/z   2 relative_disparity: 1 - (1/(z+eps) - 1/(far+eps)) / (1/(near+eps) - 1/(far+eps) + eps)
 *   3 log: log(max(min(z, near), far))   (literal .minimum(near).maximum(far).log()),
 * blended with the colour weights, no background.  Returns the value and its derivative w.r.t. the sort depth
 * (what autograd gives through the reference's torch expressions). */
static float depth_value(int mode, float sortdepth, float inv_scale, float near, float far, float* dval_dsort) {
  const float z = sortdepth * inv_scale, eps = 1e-10f;
  float v, d;
  if (mode == 1) { v = 1.f / z; d = -1.f / (z * z); }
  else if (mode == 2) {
    float dn = 1.f / (near + eps), df = 1.f / (far + eps), zi = 1.f / (z + eps);
    v = 1.f - (zi - df) / (dn - df + eps); d = zi * zi / (dn - df + eps);
  } else if (mode == 3) {
    float m = fminf(z, near);
    v = logf(fmaxf(m, far)); d = (z <= near && m >= far) ? 1.f / z : 0.f;
  } else { v = z; d = 1.f; }
  if (dval_dsort) *dval_dsort = d * inv_scale;
  return v;
}

int oracle_render_ex(const OrcCfg* c, const float* means, const float* cov6, const float* opac,
                  const float* shs, const float* colors,
                  float* out_color, int32_t* out_radii, float* out_final_T, uint32_t* out_n_contrib,
                  float* g_xy, float* g_depth, float* g_conic_op, float* g_rgb, uint32_t* g_tiles,
                  uint8_t* g_clamped,
                  int64_t inst_cap, uint32_t* inst_tile, uint32_t* inst_gid, uint32_t* tile_ranges,
                  int64_t* num_rendered,
                  const float* dL_dpix, float* d_means, float* d_means2D, float* d_cov6, float* d_opac,
                  float* d_shs, float* d_colors,
                  int depth_mode /* -1: no depth channel */, float depth_inv_scale, float depth_near, float depth_far,
                  float* out_depth /* [H,W] or NULL */, const float* dL_ddepth /* [H,W] or NULL */) {
  const int with_depth = depth_mode >= 0 && depth_mode <= 3;
  const int P = c->P, H = c->H, W = c->W;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE, ntiles = gx * gy;
  const int deg = c->D < c->max_sh_degree ? c->D : c->max_sh_degree;

  float* xy = (float*)calloc((size_t)P * 2 + 1, sizeof(float));
  float* depth = (float*)calloc((size_t)P + 1, sizeof(float));
  float* conop = (float*)calloc((size_t)P * 4 + 1, sizeof(float));
  float* rgb = (float*)calloc((size_t)P * 3 + 1, sizeof(float));
  int32_t* ext = (int32_t*)calloc((size_t)P * 2 + 1, sizeof(int32_t)); /* ex, ey; ex==0 -> invisible */
  uint32_t* tiles = (uint32_t*)calloc((size_t)P + 1, sizeof(uint32_t));
  uint8_t* clamped = (uint8_t*)calloc((size_t)P * 3 + 1, 1);
  if (!xy || !depth || !conop || !rgb || !ext || !tiles || !clamped) return -1;

  /* ------------------------------------------------------------- K1 preprocess */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    const float* m = means + 3 * (size_t)i;
    float t[3];
    view_point(c->view, m, t);
    float sortkey;
    if (c->mode == 0) {
      if (t[2] <= c->near_cull) continue;
      sortkey = t[2];
    } else {
      float r = sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
      if (r <= c->near_cull) continue;
      sortkey = r;
    }
    Geo g;
    geo_compute(c, m, cov6 + 6 * (size_t)i, &g);
    float det = g.a * g.c - g.b * g.b;
    if (det == 0.f) continue;
    float det_inv = 1.f / det;
    float cA = g.c * det_inv, cB = -g.b * det_inv, cC = g.a * det_inv;
    int ex, ey;
    float px, py;
    if (c->mode == 0) {
      float mid = 0.5f * (g.a + g.c);
      float lam1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
      float lam2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
      ex = ey = (int)ceilf(3.f * sqrtf(fmaxf(lam1, lam2)));
      const float* PM = c->proj;
      float hx = PM[0] * m[0] + PM[4] * m[1] + PM[8] * m[2] + PM[12];
      float hy = PM[1] * m[0] + PM[5] * m[1] + PM[9] * m[2] + PM[13];
      float hw = PM[3] * m[0] + PM[7] * m[1] + PM[11] * m[2] + PM[15];
      float pw = 1.f / (hw + 0.0000001f);
      px = ((hx * pw + 1.f) * W - 1.f) * 0.5f;
      py = ((hy * pw + 1.f) * H - 1.f) * 0.5f;
    } else {
      const float PI = 3.14159265358979323846f;
      ex = (int)ceilf(3.f * sqrtf(g.a));
      ey = (int)ceilf(3.f * sqrtf(g.c));
      if (ex > W / 2) ex = W / 2;
      float su = -(float)W / (2.f * PI), sv = -(float)H / PI;
      px = su * atan2f(t[0], t[2]) + 0.5f * W - 0.5f;
      py = sv * atan2f(t[1], sqrtf(t[0] * t[0] + t[2] * t[2])) + 0.5f * H - 0.5f;
    }
    Rect r = get_rect(c, px, py, ex, ey, gx, gy);
    if (r.nx * r.ny == 0) continue;
    /* colour */
    float col[3];
    if (c->use_sh) {
      float dx = m[0] - c->campos[0], dy = m[1] - c->campos[1], dz = m[2] - c->campos[2];
      float inv = 1.f / sqrtf(dx * dx + dy * dy + dz * dz);
      dx *= inv; dy *= inv; dz *= inv;
      float b[25], bx[25], by[25], bz[25];
      int n = sh_basis(deg, dx, dy, dz, b, bx, by, bz);
      const float* sh = shs + (size_t)i * c->M * 3;
      for (int ch = 0; ch < 3; ch++) {
        float acc = 0.f;
        for (int k = 0; k < n; k++) acc += b[k] * sh[3 * k + ch];
        acc += 0.5f;
        clamped[3 * (size_t)i + ch] = acc < 0.f;
        col[ch] = fmaxf(acc, 0.f);
      }
    } else {
      for (int ch = 0; ch < 3; ch++) col[ch] = colors[3 * (size_t)i + ch];
    }
    xy[2 * (size_t)i] = px; xy[2 * (size_t)i + 1] = py;
    depth[i] = sortkey;
    conop[4 * (size_t)i] = cA; conop[4 * (size_t)i + 1] = cB; conop[4 * (size_t)i + 2] = cC;
    conop[4 * (size_t)i + 3] = opac[i];
    for (int ch = 0; ch < 3; ch++) rgb[3 * (size_t)i + ch] = col[ch];
    ext[2 * (size_t)i] = ex; ext[2 * (size_t)i + 1] = ey;
    tiles[i] = (uint32_t)(r.nx * r.ny);
  }

  /* --------------------------------------------------- K2..K5 scan/duplicate/sort */
  int64_t N = 0;
  uint64_t* offs = (uint64_t*)malloc(((size_t)P + 1) * sizeof(uint64_t));
  if (!offs) return -1;
  for (int i = 0; i < P; i++) { offs[i] = (uint64_t)N; N += tiles[i]; }
  if (num_rendered) *num_rendered = N;
  uint64_t* keys = (uint64_t*)malloc(((size_t)N + 1) * sizeof(uint64_t));
  uint32_t* gids = (uint32_t*)malloc(((size_t)N + 1) * sizeof(uint32_t));
  uint32_t* ranges = (uint32_t*)calloc((size_t)ntiles * 2, sizeof(uint32_t));
  if (!keys || !gids || !ranges) return -1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    if (tiles[i] == 0) continue;
    Rect r = get_rect(c, xy[2 * (size_t)i], xy[2 * (size_t)i + 1], ext[2 * (size_t)i], ext[2 * (size_t)i + 1], gx, gy);
    uint64_t o = offs[i];
    uint32_t dbits;
    {
#ifdef ORACLE_F64
#undef float
      const float d32 = (float)depth[i];
#define float double
#else
      const float d32 = depth[i];
#endif
      memcpy(&dbits, &d32, 4);
    }
    for (int ty = r.y0; ty < r.y0 + r.ny; ty++)
      for (int k = 0; k < r.nx; k++) {
        int tx = c->mode == 0 ? r.x0 + k : wrap_tile(r.x0 + k, gx);
        uint64_t key = (uint64_t)(ty * gx + tx);
        keys[o] = (key << 32) | dbits;
        gids[o] = (uint32_t)i;
        o++;
      }
  }
  int tbits = 0;
  while ((1 << tbits) < ntiles) tbits++;
  sort_pairs_u64(keys, gids, (size_t)N, 32 + tbits + 1);
  for (int64_t i = 0; i < N; i++) {
    uint32_t t = (uint32_t)(keys[i] >> 32);
    if (i == 0 || t != (uint32_t)(keys[i - 1] >> 32)) ranges[2 * t] = (uint32_t)i;
    if (i == N - 1 || t != (uint32_t)(keys[i + 1] >> 32)) ranges[2 * t + 1] = (uint32_t)(i + 1);
  }
  if (inst_tile || inst_gid) {
    if (N > inst_cap) return -2;
    for (int64_t i = 0; i < N; i++) {
      if (inst_tile) inst_tile[i] = (uint32_t)(keys[i] >> 32);
      if (inst_gid) inst_gid[i] = gids[i];
    }
  }
  if (tile_ranges) memcpy(tile_ranges, ranges, (size_t)ntiles * 2 * sizeof(uint32_t));

  /* ------------------------------------------------------------------ K6 render */
  float* final_T = (float*)malloc((size_t)H * W * sizeof(float));
  uint32_t* n_contrib = (uint32_t*)malloc((size_t)H * W * sizeof(uint32_t));
  if (!final_T || !n_contrib) return -1;
  const float halfW = 0.5f * W;
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < ntiles; tile++) {
    int ty = tile / gx, tx = tile % gx;
    uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
        if (pxi >= W || pyi >= H) continue;
        float T = 1.f, C[3] = {0, 0, 0}, Dz = 0.f;
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = s; k < e; k++) {
          contributor++;
          uint32_t g = gids[k];
          float dx = xy[2 * (size_t)g] - (float)pxi, dy = xy[2 * (size_t)g + 1] - (float)pyi;
          if (c->mode == 1) { if (dx > halfW) dx -= W; else if (dx < -halfW) dx += W; }
          const float* co = conop + 4 * (size_t)g;
          float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0.f) continue;
          float alpha = fminf(0.99f, co[3] * expf(power));
          if (alpha < 1.f / 255.f) continue;
          float test_T = T * (1.f - alpha);
          if (test_T < 0.0001f) break; /* this Gaussian is not blended */
          for (int ch = 0; ch < 3; ch++) C[ch] += rgb[3 * (size_t)g + ch] * alpha * T;
          if (with_depth) Dz += depth_value(depth_mode, depth[g], depth_inv_scale, depth_near, depth_far, NULL) * alpha * T;
          T = test_T;
          last = contributor;
        }
        size_t pid = (size_t)pyi * W + pxi;
        final_T[pid] = T; n_contrib[pid] = last;
        for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pid] = C[ch] + T * c->bg[ch];
        if (with_depth && out_depth) out_depth[pid] = Dz;
      }
  }

  if (out_radii)
    for (int i = 0; i < P; i++) out_radii[i] = ext[2 * (size_t)i] > ext[2 * (size_t)i + 1] ? ext[2 * (size_t)i] : ext[2 * (size_t)i + 1];
  if (out_radii)
    for (int i = 0; i < P; i++) if (tiles[i] == 0) out_radii[i] = 0;
  if (out_final_T) memcpy(out_final_T, final_T, (size_t)H * W * sizeof(float));
  if (out_n_contrib) memcpy(out_n_contrib, n_contrib, (size_t)H * W * sizeof(uint32_t));
  if (g_xy) memcpy(g_xy, xy, (size_t)P * 2 * sizeof(float));
  if (g_depth) memcpy(g_depth, depth, (size_t)P * sizeof(float));
  if (g_conic_op) memcpy(g_conic_op, conop, (size_t)P * 4 * sizeof(float));
  if (g_rgb) memcpy(g_rgb, rgb, (size_t)P * 3 * sizeof(float));
  if (g_tiles) memcpy(g_tiles, tiles, (size_t)P * sizeof(uint32_t));
  if (g_clamped) memcpy(g_clamped, clamped, (size_t)P * 3);

  /* =============================================================== backward */
  if (dL_dpix) {
    /* per-Gaussian screen-space accumulators: [0..2] dL/drgb, [3,4] dL/dmean2D (pixel units),
       [5..7] dL/dconic (A, B_true, C), [8] dL/dopacity */
    const int with_dgrad = with_depth && dL_ddepth != NULL;
    double* acc = (double*)calloc((size_t)P * 10 + 1, sizeof(double));
    if (!acc) return -1;
    /* ---------------------------------------------------------- K7 render bwd */
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < ntiles; tile++) {
      int ty = tile / gx, tx = tile % gx;
      uint32_t s = ranges[2 * tile], e = ranges[2 * tile + 1];
      for (int ly = 0; ly < TILE; ly++)
        for (int lx = 0; lx < TILE; lx++) {
          int pxi = tx * TILE + lx, pyi = ty * TILE + ly;
          if (pxi >= W || pyi >= H) continue;
          size_t pid = (size_t)pyi * W + pxi;
          const float T_final = final_T[pid];
          float T = T_final;
          uint32_t last_contributor = n_contrib[pid];
          float dpix[3], accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f;
          float dpd = with_dgrad ? dL_ddepth[pid] : 0.f, accum_rec_d = 0.f, last_d = 0.f;
          float bg_dot = 0.f;
          for (int ch = 0; ch < 3; ch++) { dpix[ch] = dL_dpix[(size_t)ch * H * W + pid]; bg_dot += c->bg[ch] * dpix[ch]; }
          for (uint32_t k = s + last_contributor; k-- > s;) {
            uint32_t g = gids[k];
            float dx = xy[2 * (size_t)g] - (float)pxi, dy = xy[2 * (size_t)g + 1] - (float)pyi;
            if (c->mode == 1) { if (dx > halfW) dx -= W; else if (dx < -halfW) dx += W; }
            const float* co = conop + 4 * (size_t)g;
            float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > 0.f) continue;
            float G = expf(power);
            float alpha = fminf(0.99f, co[3] * G);
            if (alpha < 1.f / 255.f) continue;
            T = T / (1.f - alpha);
            float dchannel_dcolor = alpha * T;
            float dL_dalpha = 0.f;
            double* a9 = acc + 10 * (size_t)g;
            if (with_dgrad) {   /* fourth blended channel, no background term */
              float dv = depth_value(depth_mode, depth[g], depth_inv_scale, depth_near, depth_far, NULL);
              accum_rec_d = last_alpha * last_d + (1.f - last_alpha) * accum_rec_d;
              last_d = dv;
              dL_dalpha += (dv - accum_rec_d) * dpd;
              double v = (double)(dchannel_dcolor * dpd);
#pragma omp atomic
              a9[9] += v;
            }
            for (int ch = 0; ch < 3; ch++) {
              float cc = rgb[3 * (size_t)g + ch];
              accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
              last_color[ch] = cc;
              dL_dalpha += (cc - accum_rec[ch]) * dpix[ch];
              double v = (double)(dchannel_dcolor * dpix[ch]);
#pragma omp atomic
              a9[ch] += v;
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            float dL_dG = co[3] * dL_dalpha;
            float gdx = G * dx, gdy = G * dy;
            float dG_ddelx = -gdx * co[0] - gdy * co[1];
            float dG_ddely = -gdy * co[2] - gdx * co[1];
            double v3 = (double)(dL_dG * dG_ddelx), v4 = (double)(dL_dG * dG_ddely);
            double v5 = (double)(-0.5f * gdx * dx * dL_dG), v6 = (double)(-gdx * dy * dL_dG), v7 = (double)(-0.5f * gdy * dy * dL_dG);
            double v8 = (double)(G * dL_dalpha);
#pragma omp atomic
            a9[3] += v3;
#pragma omp atomic
            a9[4] += v4;
#pragma omp atomic
            a9[5] += v5;
#pragma omp atomic
            a9[6] += v6;
#pragma omp atomic
            a9[7] += v7;
#pragma omp atomic
            a9[8] += v8;
          }
        }
    }

    /* ------------------------------------------- K8 cov2D bwd + K9 preprocess bwd */
    float R[3][3];
    view_rot(c->view, R);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
      float dm[3] = {0, 0, 0}, dcov[6] = {0, 0, 0, 0, 0, 0}, dop = 0.f, dm2[3] = {0, 0, 0};
      float dcol[3] = {0, 0, 0};
      if (tiles[i] != 0) {
        const double* a9 = acc + 10 * (size_t)i;
        const float* m = means + 3 * (size_t)i;
        const float* cv = cov6 + 6 * (size_t)i;
        float gu = (float)a9[3], gv = (float)a9[4];           /* pixel units */
        float gA = (float)a9[5], gB = (float)a9[6], gC = (float)a9[7];
        dop = (float)a9[8];
        for (int ch = 0; ch < 3; ch++) dcol[ch] = (float)a9[ch];
        /* screen-space mean gradient is reported in NDC units (Appendix A, K7) */
        dm2[0] = gu * 0.5f * W; dm2[1] = gv * 0.5f * H;
        Geo g;
        geo_compute(c, m, cv, &g);
        /* dL/d(a,b,c) from dL/dconic; 1/(det^2 + 1e-7) as upstream */
        float denom = g.a * g.c - g.b * g.b;
        float inv2 = 1.f / (denom * denom + 0.0000001f);
        float da = inv2 * (-g.c * g.c * gA + g.b * g.c * gB + (denom - g.a * g.c) * gC);
        float dc = inv2 * (-g.a * g.a * gC + g.a * g.b * gB + (denom - g.a * g.c) * gA);
        float db = inv2 * (2.f * g.b * g.c * gA - (denom + 2.f * g.b * g.b) * gB + 2.f * g.a * g.b * gC);
        const float (*Mm)[3] = g.Mm;
        dcov[0] = Mm[0][0] * Mm[0][0] * da + Mm[0][0] * Mm[1][0] * db + Mm[1][0] * Mm[1][0] * dc;
        dcov[3] = Mm[0][1] * Mm[0][1] * da + Mm[0][1] * Mm[1][1] * db + Mm[1][1] * Mm[1][1] * dc;
        dcov[5] = Mm[0][2] * Mm[0][2] * da + Mm[0][2] * Mm[1][2] * db + Mm[1][2] * Mm[1][2] * dc;
        dcov[1] = 2.f * Mm[0][0] * Mm[0][1] * da + (Mm[0][0] * Mm[1][1] + Mm[0][1] * Mm[1][0]) * db + 2.f * Mm[1][0] * Mm[1][1] * dc;
        dcov[2] = 2.f * Mm[0][0] * Mm[0][2] * da + (Mm[0][0] * Mm[1][2] + Mm[0][2] * Mm[1][0]) * db + 2.f * Mm[1][0] * Mm[1][2] * dc;
        dcov[4] = 2.f * Mm[0][2] * Mm[0][1] * da + (Mm[0][1] * Mm[1][2] + Mm[0][2] * Mm[1][1]) * db + 2.f * Mm[1][1] * Mm[1][2] * dc;
        float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
        float dM[2][3];
        for (int k = 0; k < 3; k++) {
          float Sm0 = S[k][0] * Mm[0][0] + S[k][1] * Mm[0][1] + S[k][2] * Mm[0][2];
          float Sm1 = S[k][0] * Mm[1][0] + S[k][1] * Mm[1][1] + S[k][2] * Mm[1][2];
          dM[0][k] = 2.f * da * Sm0 + db * Sm1;
          dM[1][k] = 2.f * dc * Sm1 + db * Sm0;
        }
        float dJ[2][3];
        for (int r = 0; r < 2; r++)
          for (int k = 0; k < 3; k++) dJ[r][k] = R[k][0] * dM[r][0] + R[k][1] * dM[r][1] + R[k][2] * dM[r][2];
        float dt[3] = {0, 0, 0};
        if (c->mode == 0) {
          float fx = (float)W / (2.f * c->tanfovx), fy = (float)H / (2.f * c->tanfovy);
          float tz = 1.f / g.tc[2], tz2 = tz * tz, tz3 = tz2 * tz;
          float xm = g.clampx ? 0.f : 1.f, ym = g.clampy ? 0.f : 1.f;
          dt[0] = xm * -fx * tz2 * dJ[0][2];
          dt[1] = ym * -fy * tz2 * dJ[1][2];
          dt[2] = -fx * tz2 * dJ[0][0] - fy * tz2 * dJ[1][1] + (2.f * fx * g.tc[0]) * tz3 * dJ[0][2] + (2.f * fy * g.tc[1]) * tz3 * dJ[1][2];
          /* K9: projected-mean gradient through the full projection matrix */
          const float* PM = c->proj;
          float hw = PM[3] * m[0] + PM[7] * m[1] + PM[11] * m[2] + PM[15];
          float mw = 1.f / (hw + 0.0000001f);
          float mul1 = (PM[0] * m[0] + PM[4] * m[1] + PM[8] * m[2] + PM[12]) * mw * mw;
          float mul2 = (PM[1] * m[0] + PM[5] * m[1] + PM[9] * m[2] + PM[13]) * mw * mw;
          dm[0] = (PM[0] * mw - PM[3] * mul1) * dm2[0] + (PM[1] * mw - PM[3] * mul2) * dm2[1];
          dm[1] = (PM[4] * mw - PM[7] * mul1) * dm2[0] + (PM[5] * mw - PM[7] * mul2) * dm2[1];
          dm[2] = (PM[8] * mw - PM[11] * mul1) * dm2[0] + (PM[9] * mw - PM[11] * mul2) * dm2[1];
        } else {
          const float PI = 3.14159265358979323846f;
          float su = -(float)W / (2.f * PI), sv = -(float)H / PI;
          float x = g.tc[0], y = g.tc[1], z = g.tc[2];
          if (!g.clampx) {
            float q = x * x + z * z, rho = sqrtf(q), r2 = q + y * y;
            float q2 = q * q;
            float f = 1.f / (rho * r2);
            float dfx = -x * (r2 + 2.f * q) / (rho * q * r2 * r2);
            float dfz = -z * (r2 + 2.f * q) / (rho * q * r2 * r2);
            float dfy = -2.f * y / (rho * r2 * r2);
            /* J00 = su z/q ; J02 = -su x/q */
            float dJ00x = -2.f * su * x * z / q2, dJ00z = su * (x * x - z * z) / q2;
            float dJ02x = su * (x * x - z * z) / q2, dJ02z = 2.f * su * x * z / q2;
            /* J10 = -sv x y f ; J12 = -sv z y f ; J11 = sv rho / r2 */
            float dJ10x = -sv * y * (f + x * dfx), dJ10y = -sv * x * (f + y * dfy), dJ10z = -sv * x * y * dfz;
            float dJ12x = -sv * z * y * dfx, dJ12y = -sv * z * (f + y * dfy), dJ12z = -sv * y * (f + z * dfz);
            float dJ11x = sv * x * (r2 - 2.f * q) / (rho * r2 * r2), dJ11z = sv * z * (r2 - 2.f * q) / (rho * r2 * r2);
            float dJ11y = -2.f * sv * rho * y / (r2 * r2);
            dt[0] = dJ[0][0] * dJ00x + dJ[0][2] * dJ02x + dJ[1][0] * dJ10x + dJ[1][1] * dJ11x + dJ[1][2] * dJ12x;
            dt[1] = dJ[1][0] * dJ10y + dJ[1][1] * dJ11y + dJ[1][2] * dJ12y;
            dt[2] = dJ[0][0] * dJ00z + dJ[0][2] * dJ02z + dJ[1][0] * dJ10z + dJ[1][1] * dJ11z + dJ[1][2] * dJ12z;
          }
          /* position path: (u,v) = f(t), Jacobian = g.J (clamped J near the pole) */
          dt[0] += g.J[0][0] * gu + g.J[1][0] * gv;
          dt[1] += g.J[0][1] * gu + g.J[1][1] * gv;
          dt[2] += g.J[0][2] * gu + g.J[1][2] * gv;
        }
        if (with_dgrad) {   /* depth value depends on the mean through the sort depth (view z, or |t| in erp mode) */
          float tv[3], dd;
          view_point(c->view, m, tv);
          depth_value(depth_mode, depth[i], depth_inv_scale, depth_near, depth_far, &dd);
          float coef = (float)a9[9] * dd;
          if (c->mode == 0) dt[2] += coef;
          else { float r = sqrtf(tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2]); for (int k = 0; k < 3; k++) dt[k] += coef * tv[k] / r; }
        }
        for (int k = 0; k < 3; k++) dm[k] += R[0][k] * dt[0] + R[1][k] * dt[1] + R[2][k] * dt[2];
        /* SH backward */
        if (c->use_sh) {
          float ox = m[0] - c->campos[0], oy = m[1] - c->campos[1], oz = m[2] - c->campos[2];
          float len2 = ox * ox + oy * oy + oz * oz;
          float inv = 1.f / sqrtf(len2);
          float dx = ox * inv, dy = oy * inv, dz = oz * inv;
          float b[25], bx[25], by[25], bz[25];
          int n = sh_basis(deg, dx, dy, dz, b, bx, by, bz);
          const float* sh = shs + (size_t)i * c->M * 3;
          float* dsh = d_shs ? d_shs + (size_t)i * c->M * 3 : NULL;
          float drgb[3];
          for (int ch = 0; ch < 3; ch++) drgb[ch] = clamped[3 * (size_t)i + ch] ? 0.f : dcol[ch];
          float ddir[3] = {0, 0, 0};
          for (int k = 0; k < n; k++)
            for (int ch = 0; ch < 3; ch++) {
              if (dsh) dsh[3 * k + ch] = b[k] * drgb[ch];
              ddir[0] += bx[k] * sh[3 * k + ch] * drgb[ch];
              ddir[1] += by[k] * sh[3 * k + ch] * drgb[ch];
              ddir[2] += bz[k] * sh[3 * k + ch] * drgb[ch];
            }
          if (dsh) for (int k = n; k < c->M; k++) for (int ch = 0; ch < 3; ch++) dsh[3 * k + ch] = 0.f;
          /* through normalisation: d(dir)/d(orig) = (I - dir dir^T)/len */
          float dot = dx * ddir[0] + dy * ddir[1] + dz * ddir[2];
          dm[0] += (ddir[0] - dx * dot) * inv;
          dm[1] += (ddir[1] - dy * dot) * inv;
          dm[2] += (ddir[2] - dz * dot) * inv;
        }
      } else if (c->use_sh && d_shs) {
        memset(d_shs + (size_t)i * c->M * 3, 0, (size_t)c->M * 3 * sizeof(float));
      }
      if (d_means) for (int k = 0; k < 3; k++) d_means[3 * (size_t)i + k] = dm[k];
      if (d_means2D) for (int k = 0; k < 3; k++) d_means2D[3 * (size_t)i + k] = dm2[k];
      if (d_cov6) for (int k = 0; k < 6; k++) d_cov6[6 * (size_t)i + k] = dcov[k];
      if (d_opac) d_opac[i] = dop;
      if (d_colors) for (int k = 0; k < 3; k++) d_colors[3 * (size_t)i + k] = c->use_sh ? 0.f : dcol[k];
    }
    free(acc);
  }

  free(xy); free(depth); free(conop); free(rgb); free(ext); free(tiles); free(clamped);
  free(offs); free(keys); free(gids); free(ranges); free(final_T); free(n_contrib);
  return 0;
}

int oracle_render(const OrcCfg* c, const float* means, const float* cov6, const float* opac,
                  const float* shs, const float* colors,
                  float* out_color, int32_t* out_radii, float* out_final_T, uint32_t* out_n_contrib,
                  float* g_xy, float* g_depth, float* g_conic_op, float* g_rgb, uint32_t* g_tiles,
                  uint8_t* g_clamped,
                  int64_t inst_cap, uint32_t* inst_tile, uint32_t* inst_gid, uint32_t* tile_ranges,
                  int64_t* num_rendered,
                  const float* dL_dpix, float* d_means, float* d_means2D, float* d_cov6, float* d_opac,
                  float* d_shs, float* d_colors) {
  return oracle_render_ex(c, means, cov6, opac, shs, colors, out_color, out_radii, out_final_T, out_n_contrib, g_xy, g_depth,
                          g_conic_op, g_rgb, g_tiles, g_clamped, inst_cap, inst_tile, inst_gid, tile_ranges, num_rendered,
                          dL_dpix, d_means, d_means2D, d_cov6, d_opac, d_shs, d_colors, -1, 1.f, 0.f, 0.f, NULL, NULL);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_cfg_size(void) { return (int)sizeof(OrcCfg); }

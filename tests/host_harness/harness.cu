// harness.cu -- TEST INFRASTRUCTURE.  Compiles the per-Gaussian math the kernels inline (splatter360_b200/csrc/persplat.cuh:
// project_view, sh_to_rgb, view_backward, depth_value; camera_math.cuh: invert4x4) for the HOST, so that tests/test_host_math.py can check the very
// code the GPU runs against the CPU oracle without a GPU.  Plain C entry points, host pointers everywhere.
#include "../../splatter360_b200/csrc/persplat.cuh"
#include "../../splatter360_b200/csrc/render_cull.cuh"
#include "../../splatter360_b200/csrc/adapter_math.cuh"
#include "../../splatter360_b200/csrc/camera_math.cuh"

using namespace s360;

extern "C" {

// K1 geometry + SH colour of every Gaussian for one view (view->viewmatrix / projmatrix / campos are HOST pointers here)
int s360h_project(const S360View* view, const float* means, const float* cov, const float* opac, const float* shs,
                  float* xy /*[P,2]*/, float* conic_op /*[P,4]*/, float* depth /*[P]*/, int32_t* radii /*[P]*/,
                  uint32_t* tiles /*[P]*/, float* rgb /*[P,3]*/, uint8_t* clamped /*[P,3]*/, float* half_extent /*[P,2] or NULL*/) {
  const S360View& v = *view;
  const float wf = view_frobenius2(v.viewmatrix);
  for (int i = 0; i < v.P; i++) {
    const float sc = v.scene_scale;
    const float mx = means[3 * i] * sc, my = means[3 * i + 1] * sc, mz = means[3 * i + 2] * sc;
    float cv[6];
    load_cov6(v, cov, i, cv);
    Proj pr;
    if (v.mode == S360_MODE_PINHOLE) project_view<S360_MODE_PINHOLE>(v, v.viewmatrix, v.projmatrix, mx, my, mz, cv, opac, i, wf, pr);
    else project_view<S360_MODE_ERP>(v, v.viewmatrix, v.projmatrix, mx, my, mz, cv, opac, i, wf, pr);
    radii[i] = pr.radius;
    tiles[i] = pr.tiles;
    xy[2 * i] = pr.px; xy[2 * i + 1] = pr.py;
    conic_op[4 * i] = pr.cA; conic_op[4 * i + 1] = pr.cB; conic_op[4 * i + 2] = pr.cC; conic_op[4 * i + 3] = pr.op;
    depth[i] = pr.sortkey;
    if (half_extent) { half_extent[2 * i] = pr.hx; half_extent[2 * i + 1] = pr.hy; }
    float col[3] = {0.f, 0.f, 0.f};
    uint8_t cl = 0;
    if (pr.upstream_visible && shs) sh_to_rgb(v, shs + (size_t)i * v.M * 3, mx, my, mz, v.campos, col, cl);
    for (int c = 0; c < 3; c++) { rgb[3 * i + c] = col[c]; clamped[3 * i + c] = (cl >> c) & 1; }
  }
  return 0;
}

// K8 geometry backward of every Gaussian from given moment sums acc[P,9] = {dL/drgb[3], sum q dx, sum q dy, sum q dx^2,
// sum q dxdy, sum q dy^2, sum q}: dL/dmean (geometry path only, no SH direction term), dL/dmean2D (NDC), dL/dcov6
int s360h_view_backward(const S360View* view, const float* means, const float* cov, const float* opac, const float* acc,
                        float* d_means /*[P,3]*/, float* d_means2D /*[P,2]*/, float* d_cov6 /*[P,6]*/) {
  const S360View& v = *view;
  const DepthSpec ds = {0, 1.f, 0.f, 0.f};
  for (int i = 0; i < v.P; i++) {
    const float sc = v.scene_scale;
    const float mx = means[3 * i] * sc, my = means[3 * i + 1] * sc, mz = means[3 * i + 2] * sc;
    float cv[6];
    load_cov6(v, cov, i, cv);
    const float* a = acc + 9 * (size_t)i;
    const float4 a0 = make_float4(a[0], a[1], a[2], a[3]), a1 = make_float4(a[4], a[5], a[6], a[7]);
    float dm[3], dm2[2], dcv[6];
    if (v.mode == S360_MODE_PINHOLE) view_backward<S360_MODE_PINHOLE, false>(v, v.viewmatrix, v.projmatrix, mx, my, mz, cv, opac[i], a0, a1, dm, dm2, dcv, ds, 0.f);
    else view_backward<S360_MODE_ERP, false>(v, v.viewmatrix, v.projmatrix, mx, my, mz, cv, opac[i], a0, a1, dm, dm2, dcv, ds, 0.f);
    for (int k = 0; k < 3; k++) d_means[3 * i + k] = dm[k] * sc;
    d_means2D[2 * i] = dm2[0]; d_means2D[2 * i + 1] = dm2[1];
    for (int k = 0; k < 6; k++) d_cov6[6 * i + k] = dcv[k] * sc * sc;
  }
  return 0;
}

// fused depth channel: per-Gaussian value and its derivative w.r.t. the sort depth
int s360h_depth_value(int mode, float inv_scale, float near, float far, int n, const float* sort_depth, float* value, float* grad) {
  const DepthSpec ds = {mode, inv_scale, near, far};
  for (int i = 0; i < n; i++) { value[i] = depth_value(ds, sort_depth[i]); grad[i] = depth_value_grad(ds, sort_depth[i]); }
  return 0;
}

// render kernels: stage an instance from its 48-B record and run the exact ellipse / 8x8-block test for a block whose
// centre is block_c; also returns the staged evaluation parameters {A', B', C', log2 o}
int s360h_cull(int n, const float* rec /*[n,12]*/, const float* block_c /*[n,2]*/, uint8_t* hit, float* ev_out /*[n,4]*/) {
  for (int i = 0; i < n; i++) {
    // rec: {x, y, conicA, conicB, conicC, opacity, hx, hy, r, g, b, depth}, packed the way K1 packs it
    const float* r = rec + 12 * (size_t)i;
    float4 r0, r1, r2;
    pack_record(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r + 8, r[11], 1 << 20, r0, r1, r2);
    float4 cull, ev, col;
    stage_instance(r0, r1, r2, cull, ev, col);
    hit[i] = rect_can_contribute(cull, ev, col.w, cull.x - block_c[2 * i], cull.y - block_c[2 * i + 1]) ? 1 : 0;
    ev_out[4 * i] = ev.x; ev_out[4 * i + 1] = ev.y; ev_out[4 * i + 2] = ev.z; ev_out[4 * i + 3] = ev.w;
  }
  return 0;
}

// fused Gaussian adapter: forward and backward of every Gaussian with the kernels' own per-Gaussian functions
// (raw [G, 7 + 3 d_sh], pose [views, 12], rot [views, 165]; cotangents g_* -> d_raw, d_depth)
int s360h_adapter(int views, int H, int W, int sh_degree, float smin, float smax, int means_grad, const float* raw,
                  const float* depth, const float* pose, const float* rot, float* means, float* cov, float* harmonics,
                  const float* g_means, const float* g_cov, const float* g_sh, float* d_raw, float* d_depth) {
  AdapterCfg c;
  c.H = H; c.W = W; c.sh_degree = sh_degree; c.d_sh = (sh_degree + 1) * (sh_degree + 1);
  c.scale_min = smin; c.scale_max = smax; c.pixel_size = 1.f / (float)(W > H ? W : H); c.eps = 1e-8f; c.means_grad = means_grad;
  const int C = 7 + 3 * c.d_sh;
  for (long g = 0; g < (long)views * H * W; g++) {
    const long bv = g / ((long)H * W);
    const int pix = (int)(g - bv * H * W), row = pix / W, col = pix - row * W;
    AdapterFwd f;
    adapter_forward_one(c, raw + g * C, depth[g], pose + bv * AD_POSE_F, row, col, means + 3 * g, cov + 9 * g, f);
    for (int ch = 0; ch < 3; ch++) {
      float* sh = harmonics + g * 3 * c.d_sh + ch * c.d_sh;
      for (int k = 0; k < c.d_sh; k++) sh[k] = raw[g * C + 7 + ch * c.d_sh + k];
      adapter_rotate_sh<false>(sh_degree, rot + bv * AD_ROT_F, sh);
    }
    if (d_raw) {
      adapter_backward_one(c, raw + g * C, depth[g], pose + bv * AD_POSE_F, f, g_means + 3 * g, g_cov + 9 * g, d_raw + g * C, d_depth[g]);
      for (int ch = 0; ch < 3; ch++) {
        float* o = d_raw + g * C + 7 + ch * c.d_sh;
        for (int k = 0; k < c.d_sh; k++) o[k] = g_sh[g * 3 * c.d_sh + ch * c.d_sh + k];
        adapter_rotate_sh<true>(sh_degree, rot + bv * AD_ROT_F, o);
      }
    }
  }
  return 0;
}

// batched 4x4 inverse exactly as camera.cu's kernel computes it, one matrix after the other
int s360h_invert4x4(const float* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; i++) invert4x4(in + i * 16, out + i * 16);
  return 0;
}

}  // extern "C"

// adapter_math.cuh -- per-Gaussian math of the fused Gaussian adapter: what
// /root/reference/src/model/encoder/common/gaussian_adapter_erp.py:49-119 does with ~25 torch ops per call
// (scale activation :63-77, quaternion normalisation :82, SH mask :85-86, build_covariance gaussians.py:8-44, rotation into
// the world frame :89-91, ERP unprojection sphere_projection.py:6-87, SH rotation sh_rotation.py:10-30) for ONE Gaussian.
// Host + device: adapter.cu inlines it; tests/host_harness builds it for the CPU, where tests/test_adapter.py checks it
// against the golden vector generated from the reference's own module.
#pragma once
#include "common.cuh"

namespace s360 {

constexpr int AD_MAX_DSH = 25;                // SH coefficients per channel, degree <= 4
constexpr int AD_POSE_F = 12;                 // per view: R (row-major 3x3), t
// per view block of the SH rotation: D^0 (1) | D^1 (9) | D^2 (25) | D^3 (49) | D^4 (81), each row-major, already
// multiplied by the SH mask of its COLUMN (sh_world = D (mask * sh_raw))
S360_HD constexpr int ad_block_offset(int l) { return l == 0 ? 0 : l == 1 ? 1 : l == 2 ? 10 : l == 3 ? 35 : 84; }
constexpr int AD_ROT_F = 165;

struct AdapterCfg {
  int H, W;              // context image size (one Gaussian per pixel)
  int d_sh;              // coefficients per channel: (sh_degree + 1)^2
  int sh_degree;
  float scale_min, scale_max, pixel_size, eps;
  int means_grad;        // 0: like the reference, the means carry no gradient (its unprojection runs under no_grad)
};

// quaternion (x, y, z, w) -> rotation matrix, as gaussians.py:9-33 (two_s = 2 / (q.q + eps))
S360_HD void ad_quat_to_matrix(const float* q, float eps, float* R) {
  const float i = q[0], j = q[1], k = q[2], r = q[3];
  const float ts = 2.f / (i * i + j * j + k * k + r * r + eps);
  R[0] = 1.f - ts * (j * j + k * k); R[1] = ts * (i * j - k * r);       R[2] = ts * (i * k + j * r);
  R[3] = ts * (i * j + k * r);       R[4] = 1.f - ts * (i * i + k * k); R[5] = ts * (j * k - i * r);
  R[6] = ts * (i * k - j * r);       R[7] = ts * (j * k + i * r);       R[8] = 1.f - ts * (i * i + j * j);
}

// unit ray of ERP pixel (row, col) in the sphere-camera frame, hm3d convention (utils360.py:93-104, 148-153)
S360_HD void ad_pixel_dir(int row, int col, int H, int W, float* d) {
  // theta = 2 pi tu, phi = pi tv; on the device sincospif takes the angle in units of pi (exact argument reduction, no
  // local-memory slow path)
  const float tu = 0.5f - ((float)col + 0.5f) / (float)W, tv = -(((float)row + 0.5f) / (float)H - 0.5f);
  float st, ct, sp, cp;
#ifdef __CUDA_ARCH__
  sincospif(2.f * tu, &st, &ct);
  sincospif(tv, &sp, &cp);
#else
  st = sinf(tu * (2.f * PI_F)); ct = cosf(tu * (2.f * PI_F)); sp = sinf(tv * PI_F); cp = cosf(tv * PI_F);
#endif
  d[0] = cp * st; d[1] = sp; d[2] = cp * ct;
}

struct AdapterFwd {
  float s[3];        // scales (camera frame, before rotation)
  float qn[4];       // normalised quaternion
  float Rq[9];       // its matrix
  float M[9];        // Rc Rq
  float sig[3];      // sigmoid of the raw scale features
  float dir[3];      // Rc dir(pixel)
};

// raw: [7 + 3 d_sh] = scale features 3 | quaternion xyzw 4 | sh [3][d_sh].  pose: R (9), t (3).
S360_HD void adapter_forward_one(const AdapterCfg& c, const float* raw, float depth, const float* pose, int row, int col,
                                 float* mean, float* cov9, AdapterFwd& f) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    f.sig[k] = 1.f / (1.f + expf(-raw[k]));
    f.s[k] = (c.scale_min + (c.scale_max - c.scale_min) * f.sig[k]) * depth * c.pixel_size;
  }
  const float nrm = sqrtf(raw[3] * raw[3] + raw[4] * raw[4] + raw[5] * raw[5] + raw[6] * raw[6]);
  const float inv = 1.f / (nrm + c.eps);
#pragma unroll
  for (int k = 0; k < 4; k++) f.qn[k] = raw[3 + k] * inv;
  ad_quat_to_matrix(f.qn, c.eps, f.Rq);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) f.M[3 * i + k] = pose[3 * i] * f.Rq[k] + pose[3 * i + 1] * f.Rq[3 + k] + pose[3 * i + 2] * f.Rq[6 + k];
  // covariance = M diag(s^2) M^T
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      cov9[3 * i + j] = f.M[3 * i] * f.M[3 * j] * f.s[0] * f.s[0] + f.M[3 * i + 1] * f.M[3 * j + 1] * f.s[1] * f.s[1] +
                        f.M[3 * i + 2] * f.M[3 * j + 2] * f.s[2] * f.s[2];
  float d[3];
  ad_pixel_dir(row, col, c.H, c.W, d);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    f.dir[i] = pose[3 * i] * d[0] + pose[3 * i + 1] * d[1] + pose[3 * i + 2] * d[2];
    mean[i] = f.dir[i] * depth + pose[9 + i];
  }
}

// sh_out[i] = sum_j Dm[i][j] sh_in[j] per band (Dm = D with masked columns); in place on one channel's d_sh coefficients
template <int L, bool TRANSPOSE>
S360_HD void adapter_rotate_band(const float* rot, float* sh) {
  constexpr int n = 2 * L + 1, o = L * L;
  const float* D = rot + ad_block_offset(L);
  float in[n], out[n];
#pragma unroll
  for (int j = 0; j < n; j++) in[j] = sh[o + j];
#pragma unroll
  for (int i = 0; i < n; i++) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < n; j++) a += (TRANSPOSE ? D[j * n + i] : D[i * n + j]) * in[j];
    out[i] = a;
  }
#pragma unroll
  for (int i = 0; i < n; i++) sh[o + i] = out[i];
}
template <bool TRANSPOSE>
S360_HD void adapter_rotate_sh(int sh_degree, const float* rot, float* sh) {
  adapter_rotate_band<0, TRANSPOSE>(rot, sh);
  if (sh_degree >= 1) adapter_rotate_band<1, TRANSPOSE>(rot, sh);
  if (sh_degree >= 2) adapter_rotate_band<2, TRANSPOSE>(rot, sh);
  if (sh_degree >= 3) adapter_rotate_band<3, TRANSPOSE>(rot, sh);
  if (sh_degree >= 4) adapter_rotate_band<4, TRANSPOSE>(rot, sh);
}

// backward of adapter_forward_one for cotangents g_mean [3] (used only with cfg.means_grad), g_cov [9] (general, not
// necessarily symmetric): d_raw[0..6] (scale features, quaternion) and d_depth
S360_HD void adapter_backward_one(const AdapterCfg& c, const float* raw, float depth, const float* pose, const AdapterFwd& f,
                                  const float* g_mean, const float* g_cov, float* d_raw7, float& d_depth) {
  // Gs = G + G^T;  d(s_k) = 2 s_k m_k^T G m_k;  dM = Gs M S^2
  float Gs[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Gs[3 * i + j] = g_cov[3 * i + j] + g_cov[3 * j + i];
  float GM[9];   // Gs M
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) GM[3 * i + k] = Gs[3 * i] * f.M[k] + Gs[3 * i + 1] * f.M[3 + k] + Gs[3 * i + 2] * f.M[6 + k];
  d_depth = 0.f;
  float dM[9];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float mGm = 0.5f * (f.M[k] * GM[k] + f.M[3 + k] * GM[3 + k] + f.M[6 + k] * GM[6 + k]);   // m_k^T G m_k
    const float ds = 2.f * f.s[k] * mGm;
    d_raw7[k] = ds * (c.scale_max - c.scale_min) * f.sig[k] * (1.f - f.sig[k]) * depth * c.pixel_size;
    d_depth += ds * (c.scale_min + (c.scale_max - c.scale_min) * f.sig[k]) * c.pixel_size;
#pragma unroll
    for (int i = 0; i < 3; i++) dM[3 * i + k] = GM[3 * i + k] * f.s[k] * f.s[k];
  }
  // dRq = Rc^T dM
  float dR[9];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int k = 0; k < 3; k++) dR[3 * a + k] = pose[a] * dM[k] + pose[3 + a] * dM[3 + k] + pose[6 + a] * dM[6 + k];
  // Rq = I + ts P(q): dts = sum dR . P, dP = ts dR
  const float i = f.qn[0], j = f.qn[1], k = f.qn[2], r = f.qn[3];
  const float n2 = i * i + j * j + k * k + r * r;
  const float ts = 2.f / (n2 + c.eps);
  const float P[9] = {-(j * j + k * k), i * j - k * r, i * k + j * r, i * j + k * r, -(i * i + k * k), j * k - i * r,
                      i * k - j * r, j * k + i * r, -(i * i + j * j)};
  float dts = 0.f;
#pragma unroll
  for (int a = 0; a < 9; a++) dts += dR[a] * P[a];
  float dP[9];
#pragma unroll
  for (int a = 0; a < 9; a++) dP[a] = ts * dR[a];
  float dq[4];
  dq[0] = dP[1] * j + dP[2] * k + dP[3] * j - 2.f * i * dP[4] - r * dP[5] + dP[6] * k + dP[7] * r - 2.f * i * dP[8];
  dq[1] = -2.f * j * dP[0] + dP[1] * i + dP[2] * r + dP[3] * i + dP[5] * k - r * dP[6] + dP[7] * k - 2.f * j * dP[8];
  dq[2] = -2.f * k * dP[0] - r * dP[1] + dP[2] * i + r * dP[3] - 2.f * k * dP[4] + dP[5] * j + dP[6] * i + dP[7] * j;
  dq[3] = -k * dP[1] + j * dP[2] + k * dP[3] - i * dP[5] - j * dP[6] + i * dP[7];
  const float dn2 = -0.5f * ts * ts * dts;   // d ts / d n2 = -2 / (n2 + eps)^2
  dq[0] += 2.f * i * dn2; dq[1] += 2.f * j * dn2; dq[2] += 2.f * k * dn2; dq[3] += 2.f * r * dn2;
  // qn = u / (|u| + eps)
  const float nrm = sqrtf(raw[3] * raw[3] + raw[4] * raw[4] + raw[5] * raw[5] + raw[6] * raw[6]);
  const float inv = 1.f / (nrm + c.eps);
  const float dot = dq[0] * raw[3] + dq[1] * raw[4] + dq[2] * raw[5] + dq[3] * raw[6];
  const float coef = nrm > 0.f ? dot * inv * inv / nrm : 0.f;
#pragma unroll
  for (int a = 0; a < 4; a++) d_raw7[3 + a] = dq[a] * inv - raw[3 + a] * coef;
  if (c.means_grad) d_depth += g_mean[0] * f.dir[0] + g_mean[1] * f.dir[1] + g_mean[2] * f.dir[2];
}

}  // namespace s360

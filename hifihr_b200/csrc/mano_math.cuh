// Per-joint math of the articulated hand model (host + device).
// Semantics follow utils/manopth/rodrigues_layer.py:15-54 (axis-angle -> quaternion ->
// re-normalised quaternion -> R) and the kinematic chain of utils/my_mano.py:398-439.
#pragma once
#include "common.cuh"

// axis-angle v[3] -> R[9] (row-major).  `+1e-8` is added to every component before the norm
// and the quaternion is re-normalised, exactly as the reference does.
HFR_HD void hfr_rodrigues_fwd(const float* v, float* R) {
  const float ax = v[0] + 1e-8f, ay = v[1] + 1e-8f, az = v[2] + 1e-8f;
  const float n = sqrtf(ax * ax + ay * ay + az * az);
  const float ux = v[0] / n, uy = v[1] / n, uz = v[2] / n;
  const float h = n * 0.5f;
  const float c = cosf(h), s = sinf(h);
  float qw = c, qx = s * ux, qy = s * uy, qz = s * uz;
  const float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw /= qn; qx /= qn; qy /= qn; qz /= qn;
  const float w2 = qw * qw, x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
  const float wx = qw * qx, wy = qw * qy, wz = qw * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;    R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy;    R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy;    R[7] = 2 * wx + 2 * yz;    R[8] = w2 - x2 - y2 + z2;
}

// g (dL/dR, 9) -> gv (dL/dv, 3)
HFR_HD void hfr_rodrigues_bwd(const float* v, const float* g, float* gv) {
  const float ax = v[0] + 1e-8f, ay = v[1] + 1e-8f, az = v[2] + 1e-8f;
  const float n = sqrtf(ax * ax + ay * ay + az * az);
  const float ux = v[0] / n, uy = v[1] / n, uz = v[2] / n;
  const float h = n * 0.5f;
  const float c = cosf(h), s = sinf(h);
  const float q0 = c, q1 = s * ux, q2 = s * uy, q3 = s * uz;
  const float qn = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  const float w = q0 / qn, x = q1 / qn, y = q2 / qn, z = q3 / qn;
  // dL/d(normalised quaternion)
  const float gw = 2 * w * (g[0] + g[4] + g[8]) + 2 * (-z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7]);
  const float gx = 2 * x * (g[0] - g[4] - g[8]) + 2 * (y * g[1] + z * g[2] + y * g[3] - w * g[5] + z * g[6] + w * g[7]);
  const float gy = 2 * y * (-g[0] + g[4] - g[8]) + 2 * (x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7]);
  const float gz = 2 * z * (-g[0] - g[4] + g[8]) + 2 * (-w * g[1] + x * g[2] + w * g[3] + y * g[5] + x * g[6] + y * g[7]);
  // through q / |q|
  const float dotq = w * gw + x * gx + y * gy + z * gz;
  const float g0 = (gw - w * dotq) / qn, g1 = (gx - x * dotq) / qn, g2 = (gy - y * dotq) / qn,
              g3 = (gz - z * dotq) / qn;
  // q = (cos h, sin h * u)
  const float gh = -s * g0 + c * (ux * g1 + uy * g2 + uz * g3);
  const float gux = s * g1, guy = s * g2, guz = s * g3;
  // u = v / n ; h = n/2 ; n = |v + 1e-8|
  const float gn = -(gux * v[0] + guy * v[1] + guz * v[2]) / (n * n) + 0.5f * gh;
  gv[0] = gux / n + gn * ax / n;
  gv[1] = guy / n + gn * ay / n;
  gv[2] = guz / n + gn * az / n;
}

// 3x4 rigid transforms stored row-major as 12 floats [R | t].
// out = P ∘ [Rl | tl]   (G_child = G_parent * local)
HFR_HD void hfr_rigid_compose(const float* P, const float* Rl, const float* tl, float* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[r * 4 + c] = P[r * 4 + 0] * Rl[0 * 3 + c] + P[r * 4 + 1] * Rl[1 * 3 + c] + P[r * 4 + 2] * Rl[2 * 3 + c];
    out[r * 4 + 3] = P[r * 4 + 0] * tl[0] + P[r * 4 + 1] * tl[1] + P[r * 4 + 2] * tl[2] + P[r * 4 + 3];
  }
}

// Backward of compose: given gG (12) for the child, accumulate gP (12), and produce gRl (9), gtl (3).
HFR_HD void hfr_rigid_compose_bwd(const float* P, const float* Rl, const float* tl, const float* gG, float* gP,
                                  float* gRl, float* gtl) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      gRl[k * 3 + c] = P[0 * 4 + k] * gG[0 * 4 + c] + P[1 * 4 + k] * gG[1 * 4 + c] + P[2 * 4 + k] * gG[2 * 4 + c];
    gtl[k] = P[0 * 4 + k] * gG[0 * 4 + 3] + P[1 * 4 + k] * gG[1 * 4 + 3] + P[2 * 4 + k] * gG[2 * 4 + 3];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      gP[r * 4 + k] += gG[r * 4 + 0] * Rl[k * 3 + 0] + gG[r * 4 + 1] * Rl[k * 3 + 1] + gG[r * 4 + 2] * Rl[k * 3 + 2] +
                       gG[r * 4 + 3] * tl[k];
    gP[r * 4 + 3] += gG[r * 4 + 3];
  }
}

// Keypoint projection + keypoint / mesh-regulariser losses next to the render path (sm_100a), forward + backward.
//
//   j2d        = proj_func(joints + root_xyz, K)                 utils/traineval_util.py:338-354, utils/fh_utils.py:30-39
//   joint_2d   = base(j2d_gt, j2d)                               losses.py:244-248   (base = L1 mean or MSE, :239-242)
//   joint_3d   = base(joints, joints_gt)                         losses.py:251-255
//   vert_3d    = base(mano_verts, verts_gt)                      losses.py:261-265
//   bone_direc(_3d) = mean conf * |unit(bone) - unit(bone_gt)|^2 losses.py:268-282, utils/losses_util.py:217-282
//   edge_length = mean | |edge| - |edge_gt| | over 3 edges/face  losses.py:285-289, utils/losses_util.py:284-301
//   mscale     = mean | |j_a - j_b| - 0.0282 |                   losses.py:293-299
//   triangle   = uniform Laplacian smoothing, mean_n mean_v |mean_{j in N(v)} x_j - x_v|   losses.py:422-429,
//                utils/losses_util.py:340-364 (pytorch3d mesh_laplacian_smoothing(method="uniform"))
//
// The reference runs ~60 tiny ATen kernels for these (two bmm against constant 0/+-1 matrices, boolean-mask gathers,
// six fancy-index gathers over the faces).  Here: one CTA per sample, joints / gradients staged in shared memory,
// bones and edges walked from index tables, per-term partial sums reduced in the block and added to 8 global floats.
// The backward re-derives the few intermediates instead of storing them and writes g_joints / g_verts in one pass
// (shared-memory accumulation, coalesced write-out), ready for hfr_geom_backward.
#include "common.cuh"

namespace {
constexpr int kThreads = 256;
constexpr int kMaxJ = 64;

struct BoneTerm {
  float d;        // conf * |u - t|^2
  float gv[3];    // d(d)/d(bone vector) (without the term weight)
};

// unit(v) = v / (|v| + 1e-4)
template <int D>
__device__ __forceinline__ BoneTerm bone_term(const float* jc, const float* jp, const float* gc, const float* gp, float conf,
                                              bool want_grad) {
  float v[3] = {0.f, 0.f, 0.f}, t[3] = {0.f, 0.f, 0.f};
  float n2 = 0.f, m2 = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    v[c] = jc[c] - jp[c]; t[c] = gc[c] - gp[c];
    n2 += v[c] * v[c]; m2 += t[c] * t[c];
  }
  const float n = sqrtf(n2), s = 1.0f / (n + 1e-4f), st = 1.0f / (sqrtf(m2) + 1e-4f);
  BoneTerm o;
  o.d = 0.f;
  float gu[3] = {0.f, 0.f, 0.f}, guv = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const float e = v[c] * s - t[c] * st;
    o.d += e * e;
    gu[c] = 2.0f * e * conf;
    guv += gu[c] * v[c];
  }
  o.d *= conf;
  o.gv[0] = o.gv[1] = o.gv[2] = 0.f;
  if (want_grad) {
    const float k = n > 0.0f ? guv * s * s / n : 0.0f;
#pragma unroll
    for (int c = 0; c < D; ++c) o.gv[c] = gu[c] * s - k * v[c];
  }
  return o;
}

// row v of the uniform Laplacian applied to the vertices: mean of the neighbours minus the vertex (the diagonal is -1
// for every vertex, an isolated vertex keeps -x_v as upstream does)
__device__ __forceinline__ void lap_row(const HfrKeypointArgs& a, const float* __restrict__ pv, int v, float* L) {
  const int p0 = __ldg(a.nbr_ptr + v), p1 = __ldg(a.nbr_ptr + v + 1);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int p = p0; p < p1; ++p) {
    const int j = __ldg(a.nbr_idx + p);
    s0 += __ldg(pv + 3 * j); s1 += __ldg(pv + 3 * j + 1); s2 += __ldg(pv + 3 * j + 2);
  }
  const float id = p1 > p0 ? 1.0f / (float)(p1 - p0) : 0.0f;
  L[0] = s0 * id - __ldg(pv + 3 * v); L[1] = s1 * id - __ldg(pv + 3 * v + 1); L[2] = s2 * id - __ldg(pv + 3 * v + 2);
}

__device__ __forceinline__ float base_val(float d, int l2) { return l2 ? d * d : fabsf(d); }
__device__ __forceinline__ float base_grad(float d, int l2) { return l2 ? 2.0f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }

// joints -> smem, j2d -> smem (+ optional global).  s_j (NJ*3), s_2 (NJ*2), s_w (NJ): 1/w of the projection
__device__ void stage_joints(const HfrKeypointArgs& a, int b, float* s_j, float* s_2, float* s_w) {
  const int tid = threadIdx.x, NJ = a.NJ;
  for (int i = tid; i < NJ * 3; i += kThreads) s_j[i] = a.joints[(size_t)b * NJ * 3 + i];
  __syncthreads();
  if (a.Kmat && tid < NJ) {
    const float* K = a.Kmat + (size_t)b * 9;
    float X[3];
    for (int c = 0; c < 3; ++c) X[c] = s_j[3 * tid + c] + (a.root_xyz ? a.root_xyz[(size_t)b * 3 + c] : 0.0f);
    const float u = K[0] * X[0] + K[1] * X[1] + K[2] * X[2];
    const float v = K[3] * X[0] + K[4] * X[1] + K[5] * X[2];
    const float w = K[6] * X[0] + K[7] * X[1] + K[8] * X[2];
    s_2[2 * tid] = u / w; s_2[2 * tid + 1] = v / w; s_w[tid] = 1.0f / w;
    if (a.j2d) { a.j2d[((size_t)b * NJ + tid) * 2] = u / w; a.j2d[((size_t)b * NJ + tid) * 2 + 1] = v / w; }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) keypoint_fwd_kernel(HfrKeypointArgs a) {
  __shared__ float s_j[kMaxJ * 3], s_2[kMaxJ * 2], s_w[kMaxJ], s_red[kThreads / 32][HFR_KP_NSUMS];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NJ = a.NJ;
  stage_joints(a, b, s_j, s_2, s_w);
  float acc[HFR_KP_NSUMS];
#pragma unroll
  for (int i = 0; i < HFR_KP_NSUMS; ++i) acc[i] = 0.f;
  const bool has2d = a.Kmat && a.j2d_gt;
  if (has2d) {
    const float* g2 = a.j2d_gt + (size_t)b * NJ * 2;
    for (int i = tid; i < NJ * 2; i += kThreads) acc[HFR_KP_J2D] += base_val(g2[i] - s_2[i], a.l2);
  }
  if (a.joints_gt) {
    const float* g3 = a.joints_gt + (size_t)b * NJ * 3;
    for (int i = tid; i < NJ * 3; i += kThreads) acc[HFR_KP_J3D] += base_val(s_j[i] - g3[i], a.l2);
  }
  for (int i = tid; i < a.NB; i += kThreads) {
    const int c = a.bone_child[i], p = a.bone_parent[i];
    const float conf = a.conf ? a.conf[(size_t)b * NJ + p] * a.conf[(size_t)b * NJ + c] : 1.0f;
    if (has2d) {
      const float* g2 = a.j2d_gt + (size_t)b * NJ * 2;
      acc[HFR_KP_BONE2D] += bone_term<2>(s_2 + 2 * c, s_2 + 2 * p, g2 + 2 * c, g2 + 2 * p, conf, false).d;
    }
    if (a.joints_gt) {
      const float* g3 = a.joints_gt + (size_t)b * NJ * 3;
      acc[HFR_KP_BONE3D] += bone_term<3>(s_j + 3 * c, s_j + 3 * p, g3 + 3 * c, g3 + 3 * p, conf, false).d;
    }
  }
  if (a.verts && a.verts_gt) {
    const float* pv = a.verts + (size_t)b * a.V * 3;
    const float* gv = a.verts_gt + (size_t)b * a.V * 3;
    for (int i = tid; i < a.V * 3; i += kThreads) acc[HFR_KP_V3D] += base_val(pv[i] - gv[i], a.l2);
    for (int f = tid; f < a.F; f += kThreads) {
      const int i0 = a.faces[3 * f], i1 = a.faces[3 * f + 1], i2 = a.faces[3 * f + 2];
      float P[9], G[9];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        P[c] = __ldg(pv + 3 * i0 + c); P[3 + c] = __ldg(pv + 3 * i1 + c); P[6 + c] = __ldg(pv + 3 * i2 + c);
        G[c] = __ldg(gv + 3 * i0 + c); G[3 + c] = __ldg(gv + 3 * i1 + c); G[6 + c] = __ldg(gv + 3 * i2 + c);
      }
      const int ea[3] = {0, 0, 1}, eb[3] = {1, 2, 2};
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        float d2 = 0.f, g2 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float x = P[3 * ea[e] + c] - P[3 * eb[e] + c], y = G[3 * ea[e] + c] - G[3 * eb[e] + c];
          d2 += x * x; g2 += y * y;
        }
        acc[HFR_KP_EDGE] += fabsf(sqrtf(d2) - sqrtf(g2));
      }
    }
  }
  if (tid == 0 && a.scale_a >= 0) {
    float d2 = 0.f;
    for (int c = 0; c < 3; ++c) { const float x = s_j[3 * a.scale_a + c] - s_j[3 * a.scale_b + c]; d2 += x * x; }
    acc[HFR_KP_MSCALE] += fabsf(sqrtf(d2) - a.scale_len);
  }
  if (a.verts && a.nbr_ptr) {
    const float* pv = a.verts + (size_t)b * a.V * 3;
    for (int v = tid; v < a.V; v += kThreads) {
      float L[3];
      lap_row(a, pv, v, L);
      acc[HFR_KP_LAP] += sqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    }
  }
#pragma unroll
  for (int i = 0; i < HFR_KP_NSUMS; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  if (tid < HFR_KP_NSUMS) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += s_red[w][tid];
    if (t != 0.f) atomicAdd(a.sums + tid, t);
  }
}

__global__ void __launch_bounds__(kThreads) keypoint_bwd_kernel(HfrKeypointBwdArgs q) {
  extern __shared__ __align__(16) float s_gv[];   // V*3 vertex gradients (only when verts are given)
  __shared__ float s_j[kMaxJ * 3], s_2[kMaxJ * 2], s_w[kMaxJ], s_gj[kMaxJ * 3], s_g2[kMaxJ * 2];
  const HfrKeypointArgs& a = q.f;
  const int b = blockIdx.x, tid = threadIdx.x, NJ = a.NJ;
  HfrKeypointArgs nf = a;
  nf.j2d = nullptr;                       // the forward already wrote it
  stage_joints(nf, b, s_j, s_2, s_w);
  const float ng = (float)q.n_global;
  // d(total)/d(term) x 1/count of the term's mean
  const float w2d = q.w[HFR_KP_J2D] / (ng * NJ * 2), w3d = q.w[HFR_KP_J3D] / (ng * NJ * 3);
  const float wv = q.w[HFR_KP_V3D] / (ng * a.V * 3), wb2 = q.w[HFR_KP_BONE2D] / (ng * a.NB);
  const float wb3 = q.w[HFR_KP_BONE3D] / (ng * a.NB), we = q.w[HFR_KP_EDGE] / (ng * a.F * 3);
  const float wm = q.w[HFR_KP_MSCALE] / ng;
  const bool has2d = a.Kmat && a.j2d_gt;
  for (int i = tid; i < NJ * 3; i += kThreads) {
    float g = 0.f;
    if (a.joints_gt) g = w3d * base_grad(s_j[i] - a.joints_gt[(size_t)b * NJ * 3 + i], a.l2);
    s_gj[i] = g;
  }
  for (int i = tid; i < NJ * 2; i += kThreads) {
    float g = 0.f;
    if (has2d) g = w2d * base_grad(s_2[i] - a.j2d_gt[(size_t)b * NJ * 2 + i], a.l2);   // base(j2d_gt, j2d) is symmetric
    s_g2[i] = g;
  }
  if (q.g_j2d_in && a.Kmat)   // upstream gradient on the j2d output itself (optional)
    for (int i = tid; i < NJ * 2; i += kThreads) s_g2[i] += q.g_j2d_in[(size_t)b * NJ * 2 + i];
  __syncthreads();
  for (int i = tid; i < a.NB; i += kThreads) {
    const int c = a.bone_child[i], p = a.bone_parent[i];
    const float conf = a.conf ? a.conf[(size_t)b * NJ + p] * a.conf[(size_t)b * NJ + c] : 1.0f;
    if (has2d && wb2 != 0.f) {
      const float* g2 = a.j2d_gt + (size_t)b * NJ * 2;
      const BoneTerm t = bone_term<2>(s_2 + 2 * c, s_2 + 2 * p, g2 + 2 * c, g2 + 2 * p, conf, true);
      for (int k = 0; k < 2; ++k) { atomicAdd(&s_g2[2 * c + k], wb2 * t.gv[k]); atomicAdd(&s_g2[2 * p + k], -wb2 * t.gv[k]); }
    }
    if (a.joints_gt && wb3 != 0.f) {
      const float* g3 = a.joints_gt + (size_t)b * NJ * 3;
      const BoneTerm t = bone_term<3>(s_j + 3 * c, s_j + 3 * p, g3 + 3 * c, g3 + 3 * p, conf, true);
      for (int k = 0; k < 3; ++k) { atomicAdd(&s_gj[3 * c + k], wb3 * t.gv[k]); atomicAdd(&s_gj[3 * p + k], -wb3 * t.gv[k]); }
    }
  }
  if (tid == 0 && a.scale_a >= 0 && wm != 0.f) {
    float x[3], d2 = 0.f;
    for (int c = 0; c < 3; ++c) { x[c] = s_j[3 * a.scale_a + c] - s_j[3 * a.scale_b + c]; d2 += x[c] * x[c]; }
    const float len = sqrtf(d2), e = len - a.scale_len;
    const float g = len > 0.f ? wm * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) / len : 0.f;
    for (int c = 0; c < 3; ++c) { atomicAdd(&s_gj[3 * a.scale_a + c], g * x[c]); atomicAdd(&s_gj[3 * a.scale_b + c], -g * x[c]); }
  }
  __syncthreads();
  // projection backward: (u/w, v/w) with (u,v,w) = K X
  if (a.Kmat && tid < NJ) {
    const float* K = a.Kmat + (size_t)b * 9;
    const float iw = s_w[tid], g0 = s_g2[2 * tid], g1 = s_g2[2 * tid + 1];
    const float gu = g0 * iw, gvv = g1 * iw, gw = -(g0 * s_2[2 * tid] + g1 * s_2[2 * tid + 1]) * iw;
    for (int c = 0; c < 3; ++c) s_gj[3 * tid + c] += K[c] * gu + K[3 + c] * gvv + K[6 + c] * gw;
  }
  __syncthreads();
  for (int i = tid; i < NJ * 3; i += kThreads) q.g_joints[(size_t)b * NJ * 3 + i] = s_gj[i];
  if (!q.g_verts) return;
  const bool hasv = a.verts && a.verts_gt;
  const float* pv = hasv ? a.verts + (size_t)b * a.V * 3 : nullptr;
  const float* gvt = hasv ? a.verts_gt + (size_t)b * a.V * 3 : nullptr;
  for (int i = tid; i < a.V * 3; i += kThreads) s_gv[i] = hasv ? wv * base_grad(pv[i] - gvt[i], a.l2) : 0.f;
  __syncthreads();
  if (hasv && we != 0.f) {
    for (int f = tid; f < a.F; f += kThreads) {
      const int id[3] = {a.faces[3 * f], a.faces[3 * f + 1], a.faces[3 * f + 2]};
      float P[9], G[9];
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) { P[3 * k + c] = __ldg(pv + 3 * id[k] + c); G[3 * k + c] = __ldg(gvt + 3 * id[k] + c); }
      const int ea[3] = {0, 0, 1}, eb[3] = {1, 2, 2};
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        float x[3], d2 = 0.f, g2 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          x[c] = P[3 * ea[e] + c] - P[3 * eb[e] + c];
          const float y = G[3 * ea[e] + c] - G[3 * eb[e] + c];
          d2 += x[c] * x[c]; g2 += y * y;
        }
        const float d = sqrtf(d2), df = d - sqrtf(g2);
        const float g = d > 0.f ? we * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) / d : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          atomicAdd(&s_gv[3 * id[ea[e]] + c], g * x[c]);
          atomicAdd(&s_gv[3 * id[eb[e]] + c], -g * x[c]);
        }
      }
    }
  }
  __syncthreads();
  const float wl = q.w[HFR_KP_LAP] / (ng * a.V);
  if (a.verts && a.nbr_ptr && wl != 0.f) {
    // d|L_v| / dx: u_v = L_v / |L_v| reaches x_v with -1 and every neighbour j with 1 / deg(v).  The unit vectors go
    // to a second shared array first, then every vertex GATHERS over its (symmetric) neighbourhood - no atomics.
    float* s_u = s_gv + a.V * 3;
    const float* xv = a.verts + (size_t)b * a.V * 3;
    for (int v = tid; v < a.V; v += kThreads) {
      float L[3];
      lap_row(a, xv, v, L);
      const float nrm = sqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
      const float deg = (float)(__ldg(a.nbr_ptr + v + 1) - __ldg(a.nbr_ptr + v));
      const float k = nrm > 0.f ? wl / nrm : 0.f;
      s_gv[3 * v] -= k * L[0]; s_gv[3 * v + 1] -= k * L[1]; s_gv[3 * v + 2] -= k * L[2];
      const float kd = deg > 0.f ? k / deg : 0.f;
      s_u[3 * v] = kd * L[0]; s_u[3 * v + 1] = kd * L[1]; s_u[3 * v + 2] = kd * L[2];
    }
    __syncthreads();
    for (int i = tid; i < a.V; i += kThreads) {
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      for (int p = __ldg(a.nbr_ptr + i); p < __ldg(a.nbr_ptr + i + 1); ++p) {
        const int v = __ldg(a.nbr_idx + p);
        g0 += s_u[3 * v]; g1 += s_u[3 * v + 1]; g2 += s_u[3 * v + 2];
      }
      s_gv[3 * i] += g0; s_gv[3 * i + 1] += g1; s_gv[3 * i + 2] += g2;
    }
    __syncthreads();
  }
  for (int i = tid; i < a.V * 3; i += kThreads) q.g_verts[(size_t)b * a.V * 3 + i] = s_gv[i];
}

int check_kp(const HfrKeypointArgs* a, const char* who) {
  HFR_CHECK_ARG(a && a->B >= 0, "%s: null argument", who);
  if (a->B == 0) return HFR_OK;
  HFR_CHECK_ARG(a->joints && a->NJ > 0 && a->NJ <= kMaxJ, "%s: joints missing or NJ out of range (<= %d)", who, kMaxJ);
  HFR_CHECK_ARG(a->NB == 0 || (a->bone_parent && a->bone_child), "%s: bone tables missing", who);
  HFR_CHECK_ARG(!(a->verts && a->verts_gt) || (a->faces && a->F > 0 && a->V > 0), "%s: verts need faces", who);
  HFR_CHECK_ARG(!a->nbr_ptr || (a->nbr_idx && a->verts && a->V > 0), "%s: the Laplacian term needs verts and the neighbour lists", who);
  HFR_CHECK_ARG(a->scale_a < a->NJ && a->scale_b < a->NJ && (a->scale_a < 0 || a->scale_b >= 0), "%s: bad mscale joints", who);
  return HFR_OK;
}
}  // namespace

extern "C" int hfr_keypoint_forward(const HfrKeypointArgs* a, void* stream) {
  if (int rc = check_kp(a, "keypoint_forward")) return rc;
  if (a->B == 0) return HFR_OK;
  HFR_CHECK_ARG(a->sums, "keypoint_forward: sums missing");
  keypoint_fwd_kernel<<<a->B, kThreads, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("keypoint_forward");
  return HFR_OK;
}

extern "C" int hfr_keypoint_backward(const HfrKeypointBwdArgs* q, void* stream) {
  HFR_CHECK_ARG(q, "keypoint_backward: null argument");
  if (int rc = check_kp(&q->f, "keypoint_backward")) return rc;
  if (q->f.B == 0) return HFR_OK;
  HFR_CHECK_ARG(q->w && q->g_joints && q->n_global > 0, "keypoint_backward: w / g_joints / n_global missing");
  const size_t smem = q->g_verts ? (size_t)q->f.V * 3 * sizeof(float) * (q->f.nbr_ptr ? 2 : 1) : 0;
  HFR_CHECK_ARG(smem <= 200 * 1024, "keypoint_backward: mesh too large for shared memory");
  if (smem > 40 * 1024) cudaFuncSetAttribute(keypoint_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  keypoint_bwd_kernel<<<q->f.B, kThreads, smem, (cudaStream_t)stream>>>(*q);
  HFR_CHECK_LAUNCH("keypoint_backward");
  return HFR_OK;
}

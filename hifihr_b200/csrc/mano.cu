// Articulated hand model (MANO / NIMBLE-shaped) forward + backward for sm_100a.
//
// Replaces ManoLayer.forward (utils/my_mano.py:315-483): pose PCA -> axis-angle ->
// Rodrigues (utils/manopth/rodrigues_layer.py) -> pose map -> shape/pose blendshapes ->
// joint regression -> kinematic chain -> linear blend skinning -> tips / reorder / centring,
// and its autograd.  One CTA per sample; the blend basis is stored coefficient-major and
// padded to float4 so every thread streams 128-bit, fully coalesced rows out of L2 (the
// 1.36 MB MANO basis stays L2-resident).  No atomics; reductions are warp shuffles.
#include <stdarg.h>

#include "common.cuh"
#include "mano_math.cuh"
#include "mano_batched.cuh"

static thread_local char g_err[512] = "";
void hfr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* hfr_last_error(void) { return g_err; }
extern "C" int hfr_abi_version(void) { return HFR_ABI_VERSION; }
extern "C" int hfr_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

#ifdef HFR_MANO_TIMING   // tuning builds only (tools/mano_phases.py): clock64 of block 0 at the phase boundaries
__device__ long long g_mano_t[32];
#define MT(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_mano_t[i] = clock64(); } while (0)
extern "C" int hfr_debug_mano_times(long long* out, int n) {
  return cudaMemcpyFromSymbol(out, g_mano_t, sizeof(long long) * (n < 32 ? n : 32)) == cudaSuccess ? 0 : 1;
}
#else
#define MT(i) do { } while (0)
#endif

namespace {

constexpr int kThreads = 1024;   // one CTA per sample: enough threads to cover every vertex / basis column and to keep
                                 // >100 independent L2 loads in flight per SM (the kernels are latency-bound)
constexpr int kMisc = 160;        // scratch floats (32 warps x 3 partial sums + totals; tips + centre offset)

struct ManoSmem {
  float* full;   // 3*NJ axis-angle
  float* R;      // NJ*9
  float* coef;   // NS + 9(NJ-1): betas then pose map
  float* J;      // NJ*3
  float* G;      // NJ*12
  float* A;      // NJ*12
  float* misc;   // kMisc scratch
  int* depth;    // NJ: depth of every joint in the kinematic tree (root = 0)
  int* parent;   // NJ: parent joint (-1 = root), a shared-memory copy: the chain loops would otherwise pay a
                 //     global-load round trip per tree level
  float* vp;     // C3 (posed rest verts)
  float* gv;     // C3 (backward only)
};

__device__ __forceinline__ ManoSmem carve(float* s, const HfrHandModel& m, bool bwd) {
  ManoSmem o;
  const int NJ = m.NJ;
  auto up4 = [](int x) { return (x + 3) & ~3; };
  o.full = s; s += up4(3 * NJ);
  o.R = s; s += up4(9 * NJ);
  o.coef = s; s += up4(m.NS + 9 * (NJ - 1));
  o.J = s; s += up4(3 * NJ);
  o.G = s; s += 12 * NJ;
  o.A = s; s += 12 * NJ;
  o.misc = s; s += kMisc;
  o.depth = reinterpret_cast<int*>(s); s += up4(NJ);
  o.parent = reinterpret_cast<int*>(s); s += up4(NJ);
  o.vp = s; s += m.C3;
  o.gv = bwd ? s : nullptr;
  return o;
}

static size_t mano_smem_bytes(const HfrHandModel& m, bool bwd) {
  auto up4 = [](int x) { return (x + 3) & ~3; };
  size_t f = up4(3 * m.NJ) + up4(9 * m.NJ) + up4(m.NS + 9 * (m.NJ - 1)) + up4(3 * m.NJ) + 24 * m.NJ + kMisc + 2 * up4(m.NJ);
  f += (size_t)m.C3 * (bwd ? 2 : 1);
  return f * sizeof(float);
}

// Phases 1-3 shared by forward and backward: pose -> R, pose map, J, G, A.
// The first n_rot joints take their rotation matrix from `rots` (rot6d root / rotmat joints,
// my_mano.py:355-373); the hand coefficients start at pose[poff].
__device__ void mano_setup(const HfrHandModel& m, const ManoSmem& s, const float* __restrict__ pose,
                           const float* __restrict__ betas, const float* __restrict__ rots, int n_rot, int poff) {
  const int tid = threadIdx.x, NJ = m.NJ, NPOSE = 3 * (NJ - 1);
  if (n_rot < NJ) {
    for (int i = tid; i < 3 * NJ; i += kThreads) {
      float v;
      if (i < 3) {
        v = n_rot > 0 ? 0.0f : pose[i];
      } else {
        const int o = i - 3;
        v = m.pose_mean ? m.pose_mean[o] : 0.0f;
        if (m.NPC > 0) {
          float h = 0.0f;
          for (int k = 0; k < m.NPC; ++k) h += pose[poff + k] * m.pca_comps[k * NPOSE + o];
          v += h;
        } else {
          v += pose[poff + o];
        }
      }
      s.full[i] = v;
    }
  }
  for (int i = tid; i < m.NS; i += kThreads) s.coef[i] = betas ? betas[i] : 0.0f;
  if (tid >= kThreads - NJ) {   // (the last warp is idle here) depth of each joint in the kinematic tree
    const int j = tid - (kThreads - NJ);
    int d = 0;
    const int pj = m.parents[j];
    for (int p = pj; p >= 0; p = m.parents[p]) ++d;
    s.depth[j] = d;
    s.parent[j] = pj;
  }
  __syncthreads();
  if (tid < NJ) {
    if (tid < n_rot) {
      for (int e = 0; e < 9; ++e) s.R[9 * tid + e] = rots[9 * tid + e];
    } else {
      hfr_rodrigues_fwd(s.full + 3 * tid, s.R + 9 * tid);
    }
  }
  for (int i = tid; i < 3 * NJ; i += kThreads) {
    float acc = m.J_template[i];
    for (int k = 0; k < m.NS; ++k) acc += m.J_shapedirs[i * m.NS + k] * s.coef[k];
    s.J[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < 9 * (NJ - 1); i += kThreads) {
    const int e = i % 9;
    s.coef[m.NS + i] = s.R[9 + i] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
  }
  // kinematic chain, one tree level at a time: 12 threads per joint (one per entry of its 3x4 transform), every
  // joint of a level in parallel - 4 steps for a hand (wrist, 3 phalanges) instead of NJ dependent ones
  {
    const int j = tid / 12, e = tid - 12 * j, r = e >> 2, c = e & 3;
    const int dj = j < NJ ? s.depth[j] : -1;
    for (int d = 0; d < NJ; ++d) {
      const bool mine = dj == d;
      if (mine) {
        const int p = s.parent[j];
        float val;
        const float* Rj = s.R + 9 * j;
        if (p < 0) {
          val = c < 3 ? Rj[r * 3 + c] : s.J[3 * j + r];
        } else {
          const float* P = s.G + 12 * p;
          if (c < 3) {
            val = P[r * 4 + 0] * Rj[0 * 3 + c] + P[r * 4 + 1] * Rj[1 * 3 + c] + P[r * 4 + 2] * Rj[2 * 3 + c];
          } else {
            const float t0 = s.J[3 * j + 0] - s.J[3 * p + 0], t1 = s.J[3 * j + 1] - s.J[3 * p + 1],
                        t2 = s.J[3 * j + 2] - s.J[3 * p + 2];
            val = P[r * 4 + 0] * t0 + P[r * 4 + 1] * t1 + P[r * 4 + 2] * t2 + P[r * 4 + 3];
          }
        }
        s.G[12 * j + e] = val;
      }
      if (!__syncthreads_or(mine)) break;   // no joint at this depth: the tree is done
    }
  }
  for (int i = tid; i < 3 * NJ; i += kThreads) {  // A_j = G_j with the rest joint removed
    const int j = i / 3, r = i % 3;
    const float* G = s.G + 12 * j;
    const float* Jj = s.J + 3 * j;
    float* A = s.A + 12 * j;
    A[r * 4 + 0] = G[r * 4 + 0];
    A[r * 4 + 1] = G[r * 4 + 1];
    A[r * 4 + 2] = G[r * 4 + 2];
    A[r * 4 + 3] = G[r * 4 + 3] - (G[r * 4 + 0] * Jj[0] + G[r * 4 + 1] * Jj[1] + G[r * 4 + 2] * Jj[2]);
  }
  __syncthreads();
}

// Phase 4: v_posed = template + sum_k coef_k * dirs_k, 128-bit coalesced rows.
__device__ void mano_blend(const HfrHandModel& m, const ManoSmem& s) {
  const int C4 = m.C3 >> 2, NK = m.NS + 9 * (m.NJ - 1);
  const float4* __restrict__ dirs4 = reinterpret_cast<const float4*>(m.dirs);
  const float4* __restrict__ vt4 = reinterpret_cast<const float4*>(m.v_template);
  float4* vp4 = reinterpret_cast<float4*>(s.vp);
  for (int c4 = threadIdx.x; c4 < C4; c4 += kThreads) {
    float4 acc = vt4[c4];
#pragma unroll 8
    for (int k = 0; k < NK; ++k) {
      const float w = s.coef[k];
      const float4 d = __ldg(dirs4 + (size_t)k * C4 + c4);
      acc.x += w * d.x; acc.y += w * d.y; acc.z += w * d.z; acc.w += w * d.w;
    }
    vp4[c4] = acc;
  }
  __syncthreads();
}

// (all weights, then all joint indices, are requested before the first one is used: two global round trips per
//  vertex instead of 2 NW dependent ones)
__device__ __forceinline__ void skin_matrix(const HfrHandModel& m, const float* A, int v, float* T) {
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.0f;
  float w[8];
  int ji[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = i < m.NW ? __ldg(m.skin_w + i * m.V + v) : 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ji[i] = w[i] != 0.0f ? __ldg(m.skin_idx + i * m.V + v) : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (w[i] != 0.0f) {
      const float* Aj = A + 12 * ji[i];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] += w[i] * Aj[e];
    }
  }
}

__device__ __forceinline__ void skin_vertex(const HfrHandModel& m, const ManoSmem& s, int v, float* out) {
  float T[12];
  skin_matrix(m, s.A, v, T);
  const float x = s.vp[3 * v], y = s.vp[3 * v + 1], z = s.vp[3 * v + 2];
#pragma unroll
  for (int r = 0; r < 3; ++r) out[r] = T[r * 4 + 0] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3];
}

__global__ void __launch_bounds__(kThreads) mano_fwd_kernel(HfrHandModel m, HfrManoFwdArgs a, int pose_dim) {
  extern __shared__ __align__(16) float smem[];
  const ManoSmem s = carve(smem, m, false);
  const int b = blockIdx.x, tid = threadIdx.x, NJ = m.NJ;
  const int poff = a.pose_off > 0 ? a.pose_off : 3;
  mano_setup(m, s, a.pose ? a.pose + (size_t)b * pose_dim : nullptr, a.betas ? a.betas + (size_t)b * m.NS : nullptr,
             a.rots ? a.rots + (size_t)b * a.n_rot_in * 9 : nullptr, a.rots ? a.n_rot_in : 0, poff);
  mano_blend(m, s);
  float* tips = s.misc;        // NT*3
  float* off = s.misc + 48;    // 3
  float* palm = s.misc + 52;   // 2*3: the two palm vertices (root_palm mode)
  if (tid < m.NT) skin_vertex(m, s, m.tip_verts[tid], tips + 3 * tid);
  else if (a.root_palm && tid >= 32 && tid < 34) skin_vertex(m, s, m.palm_verts[tid - 32], palm + 3 * (tid - 32));
  __syncthreads();
  // output joint value by source id: chain joint (its global translation), tip vertex, or the palm midpoint
  auto joint_src = [&](int src, int c) -> float {
    if (src >= NJ) return tips[3 * (src - NJ) + c];
    if (src == 0 && a.root_palm) return (palm[c] + palm[3 + c]) / 2.0f;
    return s.G[12 * src + c * 4 + 3];
  };
  if (tid < 3) {
    float o = 0.0f;
    if (a.trans) {
      o = a.trans[(size_t)b * 3 + tid];
    } else if (m.center_joint >= 0) {
      o = -joint_src(m.joint_order[m.center_joint], tid);
    }
    off[tid] = o;
  }
  __syncthreads();
  float* vout = a.verts + (size_t)b * m.V * 3;
  for (int v = tid; v < m.V; v += kThreads) {
    float p[3];
    skin_vertex(m, s, v, p);
    vout[3 * v + 0] = p[0] + off[0];
    vout[3 * v + 1] = p[1] + off[1];
    vout[3 * v + 2] = p[2] + off[2];
  }
  if (a.joints) {
    float* jout = a.joints + (size_t)b * (NJ + m.NT) * 3;
    for (int i = tid; i < (NJ + m.NT) * 3; i += kThreads) {
      const int k = i / 3, c = i % 3;
      jout[i] = joint_src(m.joint_order[k], c) + off[c];
    }
  }
}

__global__ void __launch_bounds__(kThreads) mano_bwd_kernel(HfrHandModel m, HfrManoBwdArgs a, int pose_dim) {
  extern __shared__ __align__(16) float smem[];
  const ManoSmem s = carve(smem, m, true);
  const int b = blockIdx.x, tid = threadIdx.x, NJ = m.NJ, V = m.V, NJO = NJ + m.NT;
  const int lane = tid & 31, warp = tid >> 5, nwarps = kThreads / 32;
  const int NK = m.NS + 9 * (NJ - 1), NPOSE = 3 * (NJ - 1);
  const float* pose = a.pose ? a.pose + (size_t)b * pose_dim : nullptr;
  const int poff = a.pose_off > 0 ? a.pose_off : 3, n_rot = a.rots ? a.n_rot_in : 0;
  MT(0);
  mano_setup(m, s, pose, a.betas ? a.betas + (size_t)b * m.NS : nullptr,
             a.rots ? a.rots + (size_t)b * a.n_rot_in * 9 : nullptr, n_rot, poff);
  MT(1);
  mano_blend(m, s);
  MT(2);
  // backward-only shared arrays live in the tail of the dynamic allocation (after gv)
  float* gA = s.gv + m.C3;        // NJ*12, later reused as gG
  float* gJ = gA + 12 * NJ;       // NJ*3
  float* gR = gJ + 3 * NJ;        // NJ*9
  float* gcoef = gR + 9 * NJ;     // NK
  float* gfull = gcoef + ((NK + 3) & ~3);  // 3*NJ
  float* red = s.misc;            // nwarps * 3 partials, then [3*nwarps ..] = total
  // ---- load upstream grads, fold tips / centre ------------------------------------------
  const float* gv_in = a.g_verts + (size_t)b * V * 3;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i = tid; i < m.C3; i += kThreads) {
    const float g = i < 3 * V ? gv_in[i] : 0.0f;
    s.gv[i] = g;
    const int c = i % 3;
    if (c == 0) sx += g; else if (c == 1) sy += g; else sz += g;
  }
  const float* gj_in = a.g_joints ? a.g_joints + (size_t)b * NJO * 3 : nullptr;
  if (gj_in) {
    for (int i = tid; i < NJO * 3; i += kThreads) {
      const float g = gj_in[i];
      const int c = i % 3;
      if (c == 0) sx += g; else if (c == 1) sy += g; else sz += g;
    }
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if (lane == 0) { red[warp * 3] = sx; red[warp * 3 + 1] = sy; red[warp * 3 + 2] = sz; }
  for (int i = tid; i < 12 * NJ + 3 * NJ; i += kThreads) gA[i] = 0.0f;  // gA and gJ
  __syncthreads();
  if (tid < 3) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[w * 3 + tid];
    red[3 * nwarps + tid] = t;
  }
  __syncthreads();
  MT(3);
  // gGt: translation-column grads that bypass A (chain joint outputs, centre)
  float* gGt = gfull;  // the gfull region (3*NJ floats) is free until the Rodrigues backward
  for (int i = tid; i < 3 * NJ; i += kThreads) gGt[i] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    // route a gradient on output-joint source `src` to what produced it (chain joint, tip vertex, palm midpoint)
    auto route = [&](int src, int c, float g) {
      if (src >= NJ) {
        s.gv[3 * m.tip_verts[src - NJ] + c] += g;
      } else if (src == 0 && a.root_palm) {
        s.gv[3 * m.palm_verts[0] + c] += 0.5f * g;
        s.gv[3 * m.palm_verts[1] + c] += 0.5f * g;
      } else {
        gGt[3 * src + c] += g;
      }
    };
    if (gj_in) {
      for (int k = 0; k < NJO; ++k)
        for (int c = 0; c < 3; ++c) route(m.joint_order[k], c, gj_in[3 * k + c]);
    }
    if (a.trans) {
      if (a.g_trans) for (int c = 0; c < 3; ++c) a.g_trans[(size_t)b * 3 + c] = red[3 * nwarps + c];
    } else if (m.center_joint >= 0) {
      for (int c = 0; c < 3; ++c) route(m.joint_order[m.center_joint], c, -red[3 * nwarps + c]);
    }
  }
  __syncthreads();
  MT(4);
  // ---- gA[j] = sum_v w_vj * g_v (x) [vp;1]: a thread per vertex keeps its <= NW influences in
  //      registers; per joint the 12 components are summed inside the warp (joints no lane touches are
  //      skipped) and one lane adds them to the shared accumulator
  if (m.jv_ptr) {
    // joint-major: thread = (joint, entry of its 3x4 block, quarter of the joint's weight list); the four quarters are
    // neighbouring lanes and meet in two shuffles - no atomics, no per-joint warp reductions
    const int total = NJ * 48;
    for (int base = warp * 32; base < total; base += kThreads) {
      const int idx = base + lane;
      float acc = 0.0f;
      if (idx < total) {
        const int q = idx & 3, e = (idx >> 2) % 12, j = idx / 48, r = e >> 2, c = e & 3;
        const int p1 = __ldg(m.jv_ptr + j + 1);
#pragma unroll 4
        for (int p = __ldg(m.jv_ptr + j) + q; p < p1; p += 4) {
          const int v = __ldg(m.jv_vert + p);
          const float w = __ldg(m.jv_w + p);
          acc += w * s.gv[3 * v + r] * (c < 3 ? s.vp[3 * v + c] : 1.0f);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (idx < total && (idx & 3) == 0) gA[12 * (idx / 48) + (idx >> 2) % 12] = acc;
    }
  } else
  for (int v0 = warp * 32; v0 < V; v0 += kThreads) {
    const int v = v0 + lane;
    int ji[8];
    float jw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ji[i] = -1; jw[i] = 0.0f; }
    float x = 0.f, y = 0.f, z = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
    unsigned jmask = 0;
    if (v < V) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < m.NW) {
          jw[i] = m.skin_w[i * V + v];
          ji[i] = jw[i] != 0.0f ? m.skin_idx[i * V + v] : -1;
          if (ji[i] >= 0) jmask |= 1u << ji[i];
        }
      }
      x = s.vp[3 * v]; y = s.vp[3 * v + 1]; z = s.vp[3 * v + 2];
      g0 = s.gv[3 * v]; g1 = s.gv[3 * v + 1]; g2 = s.gv[3 * v + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) jmask |= __shfl_xor_sync(0xffffffffu, jmask, o);
    while (jmask) {
      const int j = __ffs(jmask) - 1;
      jmask &= jmask - 1;
      float w = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) w += (ji[i] == j) ? jw[i] : 0.0f;
      const float gr[3] = {w * g0, w * g1, w * g2};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float t0 = warp_sum(gr[r] * x), t1 = warp_sum(gr[r] * y), t2 = warp_sum(gr[r] * z), t3 = warp_sum(gr[r]);
        if (lane == 0) {
          atomicAdd(&gA[12 * j + r * 4 + 0], t0); atomicAdd(&gA[12 * j + r * 4 + 1], t1);
          atomicAdd(&gA[12 * j + r * 4 + 2], t2); atomicAdd(&gA[12 * j + r * 4 + 3], t3);
        }
      }
    }
  }
  __syncthreads();
  MT(5);
  // ---- g_vp = T_rot^T g_v, in place
  for (int v = tid; v < V; v += kThreads) {
    float T[12];
    skin_matrix(m, s.A, v, T);
    const float g0 = s.gv[3 * v], g1 = s.gv[3 * v + 1], g2 = s.gv[3 * v + 2];
    s.gv[3 * v + 0] = T[0] * g0 + T[4] * g1 + T[8] * g2;
    s.gv[3 * v + 1] = T[1] * g0 + T[5] * g1 + T[9] * g2;
    s.gv[3 * v + 2] = T[2] * g0 + T[6] * g1 + T[10] * g2;
  }
  __syncthreads();
  MT(6);
  // ---- transposed blend contraction: g_coef[k] = <dirs_k, g_vp>  (warp per coefficient row)
  {
    const int C4 = m.C3 >> 2;
    const float4* __restrict__ dirs4 = reinterpret_cast<const float4*>(m.dirs);
    const float4* gv4 = reinterpret_cast<const float4*>(s.gv);
    // two rows per warp and round (rows k and k + nwarps share the g_vp reads): 3 rounds of 16 loads in flight per
    // lane instead of 5 rounds of 8 for the 145 MANO rows
    for (int k = warp; k < NK; k += 2 * nwarps) {
      const int k2 = k + nwarps;
      const bool two = k2 < NK;
      float acc = 0.0f, acc2 = 0.0f;
      const float4* row = dirs4 + (size_t)k * C4;
      const float4* row2 = dirs4 + (size_t)(two ? k2 : k) * C4;
#pragma unroll 8
      for (int c4 = lane; c4 < C4; c4 += 32) {   // independent 128-bit loads, 16 in flight per lane
        const float4 d = __ldg(row + c4);
        const float4 d2 = __ldg(row2 + c4);
        const float4 g = gv4[c4];
        acc += d.x * g.x + d.y * g.y + d.z * g.z + d.w * g.w;
        acc2 += d2.x * g.x + d2.y * g.y + d2.z * g.z + d2.w * g.w;
      }
      acc = warp_sum(acc);
      acc2 = warp_sum(acc2);
      if (lane == 0) { gcoef[k] = acc; if (two) gcoef[k2] = acc2; }
    }
  }
  __syncthreads();
  MT(7);
  // ---- chain backward, one tree level at a time from the leaves to the root (thread per joint).  A joint's
  //      contribution to its parent goes through a private slot and the parent sums its children in index
  //      order, so the result does not depend on thread timing.
  {
    float* gG = gA;  // converted in place: gG.R = gA.R - gA.t (x) J ; gG.t = gA.t (+ direct joint grads)
    float* contrib = gfull + 3 * NJ;   // NJ x 15: d(parent transform) 12, d(local translation) 3
    const int j = tid, dj = tid < NJ ? s.depth[tid] : -1;
    if (j < NJ) {
      const float* G = s.G + 12 * j;
      const float* Jj = s.J + 3 * j;
      for (int r = 0; r < 3; ++r) {
        const float gt = gA[12 * j + r * 4 + 3];
        for (int k = 0; k < 3; ++k) {
          gJ[3 * j + k] -= G[r * 4 + k] * gt;
          gG[12 * j + r * 4 + k] -= gt * Jj[k];
        }
        gG[12 * j + r * 4 + 3] = gt + gGt[3 * j + r];
      }
    }
    int maxd = 0;
    for (int i = 0; i < NJ; ++i) maxd = max(maxd, s.depth[i]);
    __syncthreads();
    for (int d = maxd; d >= 1; --d) {
      if (dj == d) {          // children of this level: differentiate G_j = G_p o [R_j | J_j - J_p]
        const int p = s.parent[j];
        const float tl[3] = {s.J[3 * j] - s.J[3 * p], s.J[3 * j + 1] - s.J[3 * p + 1], s.J[3 * j + 2] - s.J[3 * p + 2]};
        float gP[12], gtl[3];
#pragma unroll
        for (int e = 0; e < 12; ++e) gP[e] = 0.0f;
        hfr_rigid_compose_bwd(s.G + 12 * p, s.R + 9 * j, tl, gG + 12 * j, gP, gR + 9 * j, gtl);
#pragma unroll
        for (int e = 0; e < 12; ++e) contrib[15 * j + e] = gP[e];
        for (int k = 0; k < 3; ++k) { contrib[15 * j + 12 + k] = gtl[k]; gJ[3 * j + k] += gtl[k]; }
      }
      __syncthreads();
      if (dj == d - 1) {      // their parents gather, children in index order
        for (int ch = 0; ch < NJ; ++ch) {
          if (s.parent[ch] != j) continue;
#pragma unroll
          for (int e = 0; e < 12; ++e) gG[12 * j + e] += contrib[15 * ch + e];
          for (int k = 0; k < 3; ++k) gJ[3 * j + k] -= contrib[15 * ch + 12 + k];
        }
      }
      __syncthreads();
    }
    if (dj == 0) {
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) gR[9 * j + r * 3 + c] = gG[12 * j + r * 4 + c];
        gJ[3 * j + r] += gG[12 * j + r * 4 + 3];
      }
    }
  }
  __syncthreads();
  MT(8);
  // ---- Rodrigues backward (thread per joint) and shape grads
  if (tid < NJ) {
    float g[9];
    for (int e = 0; e < 9; ++e) g[e] = gR[9 * tid + e] + (tid >= 1 ? gcoef[m.NS + 9 * (tid - 1) + e] : 0.0f);
    float gvv[3] = {0.f, 0.f, 0.f};
    if (tid < n_rot) {
      if (a.g_rots) for (int e = 0; e < 9; ++e) a.g_rots[((size_t)b * n_rot + tid) * 9 + e] = g[e];
    } else {
      hfr_rodrigues_bwd(s.full + 3 * tid, g, gvv);
    }
    // gfull aliases gGt, which thread 0 finished reading before the barrier above
    gfull[3 * tid] = gvv[0]; gfull[3 * tid + 1] = gvv[1]; gfull[3 * tid + 2] = gvv[2];
  }
  if (a.g_betas && a.betas) {
    for (int k = tid; k < m.NS; k += kThreads) {
      float acc = gcoef[k];
      for (int i = 0; i < 3 * NJ; ++i) acc += m.J_shapedirs[i * m.NS + k] * gJ[i];
      a.g_betas[(size_t)b * m.NS + k] = acc;
    }
  }
  __syncthreads();
  MT(9);
  if (!a.g_pose || n_rot >= NJ) return;
  float* gp = a.g_pose + (size_t)b * pose_dim;
  if (tid < poff) gp[tid] = (n_rot == 0 && tid < 3) ? gfull[tid] : 0.0f;   // a matrix-driven root gets its gradient via g_rots
  if (m.NPC > 0) {
    for (int k = tid; k < m.NPC; k += kThreads) {
      float acc = 0.0f;
      for (int o = 0; o < NPOSE; ++o) acc += m.pca_comps[k * NPOSE + o] * gfull[3 + o];
      gp[poff + k] = acc;
    }
  } else {
    for (int o = tid; o < NPOSE; o += kThreads) gp[poff + o] = gfull[3 + o];
  }
  MT(10);
}

static int check_model(const HfrHandModel* m) {
  HFR_CHECK_ARG(m && m->V > 0 && m->NJ >= 1 && m->NJ <= HFR_MAX_JOINTS, "hand model: bad V/NJ");
  HFR_CHECK_ARG(m->NS >= 0 && m->NS <= 64 && m->NW >= 1 && m->NW <= 8 && m->NT >= 0 && m->NT <= 16,
                "hand model: bad NS/NW/NT");
  HFR_CHECK_ARG(m->C3 >= 3 * m->V && (m->C3 & 3) == 0, "hand model: C3 must be >= 3V and a multiple of 4");
  HFR_CHECK_ARG(m->dirs && m->v_template && m->J_template && m->J_shapedirs && m->parents && m->skin_idx &&
                    m->skin_w && m->joint_order,
                "hand model: null constant pointer");
  HFR_CHECK_ARG(m->NPC == 0 || m->pca_comps, "hand model: NPC>0 needs pca_comps");
  return HFR_OK;
}

}  // namespace

extern "C" int hfr_mano_forward(const HfrHandModel* m, const HfrManoFwdArgs* a, void* stream) {
  if (int rc = check_model(m)) return rc;
  HFR_CHECK_ARG(a && a->B >= 0, "mano_forward: null argument");
  if (a->B == 0) return HFR_OK;
  HFR_CHECK_ARG(a->verts && (a->pose || (a->rots && a->n_rot_in == m->NJ)), "mano_forward: null pointer");
  HFR_CHECK_ARG(!a->rots || a->n_rot_in == 1 || a->n_rot_in == m->NJ, "mano_forward: n_rot_in must be 1 or NJ");
  HFR_CHECK_ARG(!a->root_palm || (m->palm_verts[0] >= 0 && m->palm_verts[0] < m->V && m->palm_verts[1] >= 0 &&
                                  m->palm_verts[1] < m->V), "mano_forward: root_palm needs palm_verts");
  const int pose_dim = (a->pose_off > 0 ? a->pose_off : 3) + (m->NPC > 0 ? m->NPC : 3 * (m->NJ - 1));
  if (hfr::mano_batched_ok(m, a->B, a->workspace)) return hfr::mano_batched_forward(m, a, pose_dim, (cudaStream_t)stream);
  const size_t smem = mano_smem_bytes(*m, false);
  HFR_CHECK_ARG(smem <= 227 * 1024, "mano_forward: model too large for shared memory (%zu B)", smem);
  if (smem > 48 * 1024) cudaFuncSetAttribute(mano_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mano_fwd_kernel<<<a->B, kThreads, smem, (cudaStream_t)stream>>>(*m, *a, pose_dim);
  HFR_CHECK_LAUNCH("mano_forward");
  return HFR_OK;
}

extern "C" int hfr_mano_backward(const HfrHandModel* m, const HfrManoBwdArgs* a, void* stream) {
  if (int rc = check_model(m)) return rc;
  HFR_CHECK_ARG(a && a->B >= 0, "mano_backward: null argument");
  if (a->B == 0) return HFR_OK;
  const bool all_rot = a->rots && a->n_rot_in == m->NJ;
  HFR_CHECK_ARG(a->g_verts && ((a->pose && a->g_pose) || all_rot), "mano_backward: null pointer");
  HFR_CHECK_ARG(!a->rots || a->n_rot_in == 1 || a->n_rot_in == m->NJ, "mano_backward: n_rot_in must be 1 or NJ");
  HFR_CHECK_ARG(!a->root_palm || (m->palm_verts[0] >= 0 && m->palm_verts[0] < m->V && m->palm_verts[1] >= 0 &&
                                  m->palm_verts[1] < m->V), "mano_backward: root_palm needs palm_verts");
  const int pose_dim = (a->pose_off > 0 ? a->pose_off : 3) + (m->NPC > 0 ? m->NPC : 3 * (m->NJ - 1));
  if (hfr::mano_batched_ok(m, a->B, a->workspace)) return hfr::mano_batched_backward(m, a, pose_dim, (cudaStream_t)stream);
  const int NK = m->NS + 9 * (m->NJ - 1);
  const size_t extra = (size_t)(12 * m->NJ + 3 * m->NJ + 9 * m->NJ + ((NK + 3) & ~3) + 3 * m->NJ + 15 * m->NJ + 8) * sizeof(float);
  const size_t smem = mano_smem_bytes(*m, true) + extra;
  HFR_CHECK_ARG(smem <= 227 * 1024, "mano_backward: model too large for shared memory (%zu B)", smem);
  if (smem > 48 * 1024) cudaFuncSetAttribute(mano_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mano_bwd_kernel<<<a->B, kThreads, smem, (cudaStream_t)stream>>>(*m, *a, pose_dim);
  HFR_CHECK_LAUNCH("mano_backward");
  return HFR_OK;
}

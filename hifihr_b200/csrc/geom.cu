// Per-sample geometry between the hand layer and the rasterizer (sm_100a), forward + backward.
//
//   joints  = reorder(J_regressor @ posed verts, tips)      Freihand_trainer_mano_fullsup.py:175-215
//   root    = joints[root_id]; joints -= root; verts_rel = verts - root        models_res_nimble.py:159-166
//   view    = verts_rel + root_xyz (two offset_verts_ calls)                     models_res_nimble.py:203-205
//   ndc.xy  = ([X,Y,Z,1] K).xy / Z, ndc.z = Z                                     PerspectiveCameras + MeshRasterizer.transform
//   normals = normalize(sum of corner cross products of incident faces)          Meshes.verts_normals_packed
//
// One CTA per sample, everything staged in shared memory; all scatter patterns of the
// reference (index_add for normals, J_regressor^T in the backward) are turned into gathers
// over precomputed CSR incidence lists, so there are no atomics and results are deterministic.
#include "common.cuh"
#include "raster_tile.cuh"

namespace {
constexpr int kThreads = 1024;   // one CTA per sample; every per-vertex loop runs in a single round (latency-bound gathers)

__device__ __forceinline__ void cross3(const float* u, const float* w, float* o) {
  o[0] = u[1] * w[2] - u[2] * w[1];
  o[1] = u[2] * w[0] - u[0] * w[2];
  o[2] = u[0] * w[1] - u[1] * w[0];
}

// corners (a = v, b, c in winding order) of incidence entry e of vertex v: one 8-byte load when the topology carries
// the neighbour table, else the vf_idx -> faces chain (two dependent round trips)
__device__ __forceinline__ void incident_face(const HfrTopology& t, int e, int v, int& ia, int& ib, int& ic) {
  if (t.vf_nbr) {
    const int2 q = __ldg(reinterpret_cast<const int2*>(t.vf_nbr) + e);
    ia = v; ib = q.x; ic = q.y;
  } else {
    const int code = t.vf_idx[e], f = code >> 2, c = code & 3;
    ia = t.faces[3 * f + c]; ib = t.faces[3 * f + (c + 1) % 3]; ic = t.faces[3 * f + (c + 2) % 3];
  }
}

// raw (un-normalised) vertex normal: sum over incident (face, corner) of the corner cross product
__device__ __forceinline__ void raw_normal(const HfrTopology& t, const float* sv, int v, float* n) {
  n[0] = n[1] = n[2] = 0.0f;
  const int e0 = __ldg(t.vf_ptr + v), e1 = __ldg(t.vf_ptr + v + 1);
#pragma unroll 2
  for (int e = e0; e < e1; ++e) {
    int ia, ib, ic;
    incident_face(t, e, v, ia, ib, ic);
    float u[3], w[3], x[3];
    for (int k = 0; k < 3; ++k) { u[k] = sv[3 * ib + k] - sv[3 * ia + k]; w[k] = sv[3 * ic + k] - sv[3 * ia + k]; }
    cross3(u, w, x);
    n[0] += x[0]; n[1] += x[1]; n[2] += x[2];
  }
}

// d/d(view position) (out[0..2]) and d/d(vertex normal) (out[3..5]) of vertex v of sample b, gathered from the
// (face, tile) records of hfr_shade_backward_tiled in a FIXED order: incident (face, corner) entries in CSR order,
// the tiles of a face's range row by row.  No atomics anywhere between the fragments and this sum.
__device__ __forceinline__ void gather_face_rec(const HfrTopology& t, const float* __restrict__ rec, const uint32_t* __restrict__ ws,
                                                const hfr::WsLayout& L, int b, int v, float* out) {
#pragma unroll
  for (int i = 0; i < 6; ++i) out[i] = 0.0f;
  const int e0 = __ldg(t.vf_ptr + v), e1 = __ldg(t.vf_ptr + v + 1);
  for (int e = e0; e < e1; ++e) {
    const int code = __ldg(t.vf_idx + e), f = code >> 2, c = code & 3;
    const int64_t fp = (int64_t)b * t.F + f;
    const uint32_t r = __ldg(ws + fp);
    if (r == hfr::kEmptyRange) continue;
    const int nx = (int)((r >> 8) & 255) - (int)(r & 255) + 1, ny = (int)(r >> 24) - (int)((r >> 16) & 255) + 1;
    const size_t base = (size_t)__ldg(ws + L.blk + (fp >> 8)) + __ldg(ws + L.loc + fp);
    const float2* __restrict__ q = reinterpret_cast<const float2*>(rec + base * HFR_FACE_REC_FLOATS + 6 * c);
    for (int j = 0; j < nx * ny; ++j) {       // records are 72 B apart: 9 float2 per record
      const float2 a0 = __ldg(q + 9 * j), a1 = __ldg(q + 9 * j + 1), a2 = __ldg(q + 9 * j + 2);
      out[0] += a0.x; out[1] += a0.y; out[2] += a1.x; out[3] += a1.y; out[4] += a2.x; out[5] += a2.y;
    }
  }
}

// The same sum for ONE incidence entry (one thread per (sample, entry), chip-wide): out[idx][6]
__global__ void __launch_bounds__(256) rec_gather_kernel(HfrTopology t, const float* __restrict__ rec, const uint32_t* __restrict__ ws,
                                                         hfr::WsLayout L, int B, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int E = 3 * t.F;
  if (idx >= (int64_t)B * E) return;
  const int b = (int)(idx / E), e = (int)(idx - (int64_t)b * E);
  const int code = __ldg(t.vf_idx + e), f = code >> 2, c = code & 3;
  const int64_t fp = (int64_t)b * t.F + f;
  const uint32_t r = __ldg(ws + fp);
  float o[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (r != hfr::kEmptyRange) {
    const int nx = (int)((r >> 8) & 255) - (int)(r & 255) + 1, ny = (int)(r >> 24) - (int)((r >> 16) & 255) + 1;
    const size_t base = (size_t)__ldg(ws + L.blk + (fp >> 8)) + __ldg(ws + L.loc + fp);
    const float2* __restrict__ q = reinterpret_cast<const float2*>(rec + base * HFR_FACE_REC_FLOATS + 6 * c);
    for (int j = 0; j < nx * ny; ++j) {
      const float2 a0 = __ldg(q + 9 * j), a1 = __ldg(q + 9 * j + 1), a2 = __ldg(q + 9 * j + 2);
      o[0] += a0.x; o[1] += a0.y; o[2] += a1.x; o[3] += a1.y; o[4] += a2.x; o[5] += a2.y;
    }
  }
  float2* dst = reinterpret_cast<float2*>(out + idx * 6);
  dst[0] = make_float2(o[0], o[1]); dst[1] = make_float2(o[2], o[3]); dst[2] = make_float2(o[4], o[5]);
}

// Shared by fwd/bwd: stage verts, regress joints, find root; leaves view-space verts in s_view.
// s_pos (NOUT*3) holds un-shifted output joints, s_root[3] the predicted root.
__device__ void geom_stage(const HfrTopology& t, int B, int b, int root_out, const float* __restrict__ verts,
                           const float* __restrict__ root_xyz, float* s_v, float* s_view, float* s_j, float* s_pos,
                           float* s_root) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = t.V;
  const float* vin = verts + (size_t)b * V * 3;
  for (int i = tid; i < 3 * V; i += kThreads) s_v[i] = vin[i];
  if (tid < 3) s_root[tid] = 0.0f;
  __syncthreads();
  if (root_out >= 0) {
    for (int j = warp; j < t.NJR; j += kThreads / 32) {
      float ax = 0.f, ay = 0.f, az = 0.f;
      const int e1 = __ldg(t.jr_ptr + j + 1);
#pragma unroll 4
      for (int e = __ldg(t.jr_ptr + j) + lane; e < e1; e += 32) {   // independent (col, val) loads, 4 rounds in flight
        const int v = __ldg(t.jr_col + e);
        const float w = __ldg(t.jr_val + e);
        ax += w * s_v[3 * v]; ay += w * s_v[3 * v + 1]; az += w * s_v[3 * v + 2];
      }
      ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
      if (lane == 0) { s_j[3 * j] = ax; s_j[3 * j + 1] = ay; s_j[3 * j + 2] = az; }
    }
    __syncthreads();
    for (int i = tid; i < t.NOUT * 3; i += kThreads) {
      const int k = i / 3, c = i % 3, src = t.out_src[k];
      s_pos[i] = src >= 0 ? s_j[3 * src + c] : s_v[3 * (-(src + 1)) + c];
    }
    __syncthreads();
    if (tid < 3) s_root[tid] = s_pos[3 * root_out + tid];
    __syncthreads();
  }
  const float* rx = root_xyz ? root_xyz + (size_t)b * 3 : nullptr;
  for (int i = tid; i < 3 * V; i += kThreads) {
    const int c = i % 3;
    const float rel = s_v[i] - s_root[c];
    s_view[i] = rx ? rel + rx[c] : rel;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) geom_fwd_kernel(HfrTopology t, HfrGeomFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x, tid = threadIdx.x, V = t.V;
  float* s_v = smem;
  float* s_view = s_v + 3 * V;
  float* s_j = s_view + 3 * V;           // NJR*3
  float* s_pos = s_j + 3 * (t.NJR > 0 ? t.NJR : 1);  // NOUT*3
  float* s_root = s_pos + 3 * (t.NOUT > 0 ? t.NOUT : 1);
  geom_stage(t, a.B, b, a.root_out, a.verts, a.root_xyz, s_v, s_view, s_j, s_pos, s_root);
  const size_t base = (size_t)b * V * 3;
  if (a.joints && a.root_out >= 0)
    for (int i = tid; i < t.NOUT * 3; i += kThreads) a.joints[(size_t)b * t.NOUT * 3 + i] = s_pos[i] - s_root[i % 3];
  if (a.verts_rel)
    for (int i = tid; i < 3 * V; i += kThreads) a.verts_rel[base + i] = s_v[i] - s_root[i % 3];
  if (a.verts_view)
    for (int i = tid; i < 3 * V; i += kThreads) a.verts_view[base + i] = s_view[i];
  if (a.verts_ndc) {
    const float fx = a.focal[2 * b], fy = a.focal[2 * b + 1], px = a.prp[2 * b], py = a.prp[2 * b + 1];
    for (int v = tid; v < V; v += kThreads) {
      const float X = s_view[3 * v], Y = s_view[3 * v + 1], Z = s_view[3 * v + 2];
      a.verts_ndc[base + 3 * v + 0] = XDIV(XADD(XMUL(fx, X), XMUL(px, Z)), Z);
      a.verts_ndc[base + 3 * v + 1] = XDIV(XADD(XMUL(fy, Y), XMUL(py, Z)), Z);
      a.verts_ndc[base + 3 * v + 2] = Z;
    }
    if (a.face_verts) {   // gather the packed (F,3,3) rasterizer input from shared memory
      __syncthreads();    // everyone is done reading s_v (verts_rel above)
      for (int v = tid; v < V; v += kThreads) {
        const float X = s_view[3 * v], Y = s_view[3 * v + 1], Z = s_view[3 * v + 2];
        s_v[3 * v + 0] = XDIV(XADD(XMUL(fx, X), XMUL(px, Z)), Z);
        s_v[3 * v + 1] = XDIV(XADD(XMUL(fy, Y), XMUL(py, Z)), Z);
        s_v[3 * v + 2] = Z;
      }
      __syncthreads();
      float* fv = a.face_verts + (size_t)b * t.F * 9;
      for (int i = tid; i < t.F * 9; i += kThreads) {
        const int f = i / 9, e = i % 9;
        fv[i] = s_v[3 * t.faces[3 * f + e / 3] + e % 3];
      }
    }
  }
  if (a.vnormals) {
    for (int v = tid; v < V; v += kThreads) {
      float n[3];
      raw_normal(t, s_view, v, n);
      const float len = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-6f);
      a.vnormals[base + 3 * v + 0] = n[0] / len;
      a.vnormals[base + 3 * v + 1] = n[1] / len;
      a.vnormals[base + 3 * v + 2] = n[2] / len;
    }
  }
}

__global__ void __launch_bounds__(kThreads) geom_bwd_kernel(HfrTopology t, HfrGeomBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = t.V;
  float* s_v = smem;
  float* s_view = s_v + 3 * V;
  float* s_gN = s_view + 3 * V;          // grads wrt raw normals
  float* s_g = s_v;                      // accumulated grads wrt view / rel verts (s_v is dead after geom_stage)
  float* s_j = s_gN + 3 * V;
  float* s_pos = s_j + 3 * (t.NJR > 0 ? t.NJR : 1);
  float* s_root = s_pos + 3 * (t.NOUT > 0 ? t.NOUT : 1);
  float* s_red = s_root + 4;             // kThreads/32 warps * 3 + 3
  float* s_gj = s_red + 3 * (kThreads / 32) + 4;              // NJR*3 grads wrt regressed joints
  geom_stage(t, a.B, b, a.root_out, a.verts, a.root_xyz, s_v, s_view, s_j, s_pos, s_root);
  const size_t base = (size_t)b * V * 3;
  const bool from_rec = a.face_rec != nullptr;
  const uint32_t* __restrict__ ws = reinterpret_cast<const uint32_t*>(a.raster_ws);
  const hfr::WsLayout L = hfr::ws_layout((int64_t)a.B * t.F);
  const bool rec_bad = from_rec && a.status && __ldg(a.status) != 0u;   // record store overflowed: fail loudly (NaN)
  // 1. raw-normal grads
  for (int v = tid; v < V; v += kThreads) {
    float g[3] = {0.f, 0.f, 0.f};
    float r6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (from_rec && a.rec_partial) {
      // the entries were summed chip-wide (rec_gather_kernel); add this vertex's entries in CSR order
      const int e0 = __ldg(t.vf_ptr + v), e1 = __ldg(t.vf_ptr + v + 1);
      const float2* __restrict__ q = reinterpret_cast<const float2*>(a.rec_partial + ((size_t)b * 3 * t.F) * 6);
      for (int e = e0; e < e1; ++e) {
        const float2 a0 = __ldg(q + 3 * e), a1 = __ldg(q + 3 * e + 1), a2 = __ldg(q + 3 * e + 2);
        r6[0] += a0.x; r6[1] += a0.y; r6[2] += a1.x; r6[3] += a1.y; r6[4] += a2.x; r6[5] += a2.y;
      }
      s_g[3 * v] = r6[0]; s_g[3 * v + 1] = r6[1]; s_g[3 * v + 2] = r6[2];
    } else if (from_rec) {
      gather_face_rec(t, a.face_rec, ws, L, b, v, r6);
      // s_v is dead after geom_stage: park d/d(view) there for step 2
      s_g[3 * v] = r6[0]; s_g[3 * v + 1] = r6[1]; s_g[3 * v + 2] = r6[2];
    }
    if (a.g_vnormals || from_rec) {
      float n[3];
      raw_normal(t, s_view, v, n);
      const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      const float gx = from_rec ? r6[3] : a.g_vnormals[base + 3 * v], gy = from_rec ? r6[4] : a.g_vnormals[base + 3 * v + 1],
                  gz = from_rec ? r6[5] : a.g_vnormals[base + 3 * v + 2];
      if (len >= 1e-6f) {
        const float hx = n[0] / len, hy = n[1] / len, hz = n[2] / len;
        const float d = hx * gx + hy * gy + hz * gz;
        g[0] = (gx - hx * d) / len; g[1] = (gy - hy * d) / len; g[2] = (gz - hz * d) / len;
      } else {
        g[0] = gx / 1e-6f; g[1] = gy / 1e-6f; g[2] = gz / 1e-6f;
      }
    }
    s_gN[3 * v] = g[0]; s_gN[3 * v + 1] = g[1]; s_gN[3 * v + 2] = g[2];
  }
  __syncthreads();
  // 2. per-vertex gather of every grad that lands on view-space verts
  float sx = 0.f, sy = 0.f, sz = 0.f;
  const float fx = a.g_verts_ndc ? a.focal[2 * b] : 0.f, fy = a.g_verts_ndc ? a.focal[2 * b + 1] : 0.f;
  for (int v = tid; v < V; v += kThreads) {
    float g[3] = {0.f, 0.f, 0.f};
    if (from_rec) { g[0] = s_g[3 * v]; g[1] = s_g[3 * v + 1]; g[2] = s_g[3 * v + 2]; }
    if (rec_bad) g[0] = __int_as_float(0x7fc00000);
    if (a.g_verts_view) { g[0] += a.g_verts_view[base + 3 * v]; g[1] += a.g_verts_view[base + 3 * v + 1]; g[2] += a.g_verts_view[base + 3 * v + 2]; }
    if (a.g_verts_ndc) {
      const float X = s_view[3 * v], Y = s_view[3 * v + 1], Z = s_view[3 * v + 2];
      const float gx = a.g_verts_ndc[base + 3 * v], gy = a.g_verts_ndc[base + 3 * v + 1], gz = a.g_verts_ndc[base + 3 * v + 2];
      // x = (fx X + px Z)/Z = fx X / Z + px
      g[0] += gx * fx / Z;
      g[1] += gy * fy / Z;
      g[2] += gz - (gx * fx * X + gy * fy * Y) / (Z * Z);
    }
    if (a.g_vnormals || from_rec) {
      const int e0 = __ldg(t.vf_ptr + v), e1 = __ldg(t.vf_ptr + v + 1);
#pragma unroll 2
      for (int e = e0; e < e1; ++e) {
        int ia, ib, ic;
        incident_face(t, e, v, ia, ib, ic);
        // corner terms: N_a += (b-a)x(c-a);  N_b += (c-b)x(a-b);  N_c += (a-c)x(b-c)
        float u[3], w[3], x[3];
        const float *ga = s_gN + 3 * ia, *gb = s_gN + 3 * ib, *gc = s_gN + 3 * ic;
        // from N_a: u=b-a, w=c-a : dL/da = -(w x ga) - (ga x u)
        for (int k = 0; k < 3; ++k) { u[k] = s_view[3 * ib + k] - s_view[3 * ia + k]; w[k] = s_view[3 * ic + k] - s_view[3 * ia + k]; }
        cross3(w, ga, x); g[0] -= x[0]; g[1] -= x[1]; g[2] -= x[2];
        cross3(ga, u, x); g[0] -= x[0]; g[1] -= x[1]; g[2] -= x[2];
        // from N_b: u'=c-b, w'=a-b : dL/da = gb x u'
        for (int k = 0; k < 3; ++k) u[k] = s_view[3 * ic + k] - s_view[3 * ib + k];
        cross3(gb, u, x); g[0] += x[0]; g[1] += x[1]; g[2] += x[2];
        // from N_c: u''=a-c, w''=b-c : dL/da = w'' x gc
        for (int k = 0; k < 3; ++k) w[k] = s_view[3 * ib + k] - s_view[3 * ic + k];
        cross3(w, gc, x); g[0] += x[0]; g[1] += x[1]; g[2] += x[2];
      }
    }
    if (a.g_verts_rel) { g[0] += a.g_verts_rel[base + 3 * v]; g[1] += a.g_verts_rel[base + 3 * v + 1]; g[2] += a.g_verts_rel[base + 3 * v + 2]; }
    s_g[3 * v] = g[0]; s_g[3 * v + 1] = g[1]; s_g[3 * v + 2] = g[2];
    sx += g[0]; sy += g[1]; sz += g[2];
  }
  if (a.root_out < 0) {
    __syncthreads();
    for (int i = tid; i < 3 * V; i += kThreads) a.g_verts[base + i] = s_g[i];
    return;
  }
  // 3. root / joints: rel = v - root, joints_out = pos - root
  const float* gj_in = a.g_joints ? a.g_joints + (size_t)b * t.NOUT * 3 : nullptr;
  if (gj_in)
    for (int i = tid; i < t.NOUT * 3; i += kThreads) {
      const float g = gj_in[i];
      const int c = i % 3;
      if (c == 0) sx += g; else if (c == 1) sy += g; else sz += g;
    }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if (lane == 0) { s_red[warp * 3] = sx; s_red[warp * 3 + 1] = sy; s_red[warp * 3 + 2] = sz; }
  for (int i = tid; i < 3 * t.NJR; i += kThreads) s_gj[i] = 0.0f;
  __syncthreads();
  // one thread per (output joint, axis) - the serial version paid NOUT dependent global round trips for out_src;
  // shared-memory atomics because two outputs may name the same source
  for (int i = tid; i < t.NOUT * 3; i += kThreads) {
    const int k = i / 3, c = i - 3 * k, src = __ldg(t.out_src + k);
    float g = gj_in ? gj_in[i] : 0.0f;
    if (k == a.root_out) {
      float groot = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) groot -= s_red[w * 3 + c];
      g += groot;
    }
    if (src >= 0) atomicAdd(&s_gj[3 * src + c], g); else atomicAdd(&s_g[3 * (-(src + 1)) + c], g);
  }
  __syncthreads();
  for (int v = tid; v < V; v += kThreads) {
    float g0 = s_g[3 * v], g1 = s_g[3 * v + 1], g2 = s_g[3 * v + 2];
    const int e1 = __ldg(t.vj_ptr + v + 1);
#pragma unroll 4
    for (int e = __ldg(t.vj_ptr + v); e < e1; ++e) {
      const int j = __ldg(t.vj_row + e);
      const float w = __ldg(t.vj_val + e);
      g0 += w * s_gj[3 * j]; g1 += w * s_gj[3 * j + 1]; g2 += w * s_gj[3 * j + 2];
    }
    a.g_verts[base + 3 * v] = g0; a.g_verts[base + 3 * v + 1] = g1; a.g_verts[base + 3 * v + 2] = g2;
  }
}

// face_verts[b][f][c][:] = verts[b][faces[f][c]][:], one thread per (sample, face corner), chip-wide
__global__ void __launch_bounds__(256) face_verts_fwd_kernel(HfrTopology t, int B, const float* __restrict__ verts,
                                                             float* __restrict__ out) {
  const int64_t total = (int64_t)B * t.F * 3;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int b = (int)(i / (3 * t.F)), e = (int)(i - (int64_t)b * 3 * t.F);
    const float* __restrict__ src = verts + ((size_t)b * t.V + __ldg(t.faces + e)) * 3;
    const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
    float* dst = out + (size_t)i * 3;
    dst[0] = x; dst[1] = y; dst[2] = z;
  }
}

// its adjoint: one thread per (sample, vertex) adds the vertex's incident (face, corner) entries in CSR order
__global__ void __launch_bounds__(256) face_verts_bwd_kernel(HfrTopology t, int B, const float* __restrict__ g_fv,
                                                             float* __restrict__ g_verts) {
  const int64_t total = (int64_t)B * t.V;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int b = (int)(i / t.V), v = (int)(i - (int64_t)b * t.V);
    const int e0 = __ldg(t.vf_ptr + v), e1 = __ldg(t.vf_ptr + v + 1);
    const float* __restrict__ g = g_fv + (size_t)b * t.F * 9;
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll 4
    for (int e = e0; e < e1; ++e) {
      const int fc = __ldg(t.vf_idx + e);                      // face * 4 + corner
      const float* __restrict__ q = g + (size_t)(fc >> 2) * 9 + (fc & 3) * 3;
      sx += __ldg(q); sy += __ldg(q + 1); sz += __ldg(q + 2);
    }
    float* dst = g_verts + (size_t)i * 3;
    dst[0] = sx; dst[1] = sy; dst[2] = sz;
  }
}

static int check_topo(const HfrTopology* t, int need_joints) {
  HFR_CHECK_ARG(t && t->V > 0 && t->F > 0 && t->faces && t->vf_ptr && t->vf_idx, "topology: null/empty");
  if (need_joints)
    HFR_CHECK_ARG(t->NJR > 0 && t->NOUT > 0 && t->NOUT <= 64 && t->jr_ptr && t->jr_col && t->jr_val && t->vj_ptr &&
                      t->vj_row && t->vj_val && t->out_src,
                  "topology: joint regressor tables missing");
  return HFR_OK;
}
}  // namespace

extern "C" int64_t hfr_geom_rec_partial_floats(const HfrTopology* t, int32_t B) {
  return (t && B > 0) ? (int64_t)B * 3 * t->F * 6 : 0;
}

extern "C" int hfr_geom_forward(const HfrTopology* t, const HfrGeomFwdArgs* a, void* stream) {
  HFR_CHECK_ARG(a && a->B >= 0, "geom_forward: null argument");
  if (a->B == 0) return HFR_OK;
  HFR_CHECK_ARG(a->verts, "geom_forward: null pointer");
  if (int rc = check_topo(t, a->root_out >= 0)) return rc;
  HFR_CHECK_ARG(!a->verts_ndc || (a->focal && a->prp), "geom_forward: verts_ndc needs focal/prp");
  if (a->B == 0) return HFR_OK;
  const size_t smem = (size_t)(6 * t->V + 3 * (t->NJR > 0 ? t->NJR : 1) + 3 * (t->NOUT > 0 ? t->NOUT : 1) + 8) * sizeof(float);
  HFR_CHECK_ARG(smem <= 227 * 1024, "geom_forward: mesh too large for shared memory");
  if (smem > 48 * 1024) cudaFuncSetAttribute(geom_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  geom_fwd_kernel<<<a->B, kThreads, smem, (cudaStream_t)stream>>>(*t, *a);
  HFR_CHECK_LAUNCH("geom_forward");
  return HFR_OK;
}

extern "C" int hfr_geom_backward(const HfrTopology* t, const HfrGeomBwdArgs* a, void* stream) {
  HFR_CHECK_ARG(a && a->B >= 0, "geom_backward: null argument");
  if (a->B == 0) return HFR_OK;
  HFR_CHECK_ARG(a->verts && a->g_verts, "geom_backward: null pointer");
  if (int rc = check_topo(t, a->root_out >= 0)) return rc;
  HFR_CHECK_ARG(!a->g_verts_ndc || (a->focal && a->prp), "geom_backward: g_verts_ndc needs focal/prp");
  HFR_CHECK_ARG(!a->face_rec || (a->raster_ws && !a->g_verts_ndc && !a->g_vnormals),
                "geom_backward: face records need the rasterizer workspace and replace g_verts_ndc / g_vnormals");
  if (a->B == 0) return HFR_OK;
  const size_t smem = (size_t)(9 * t->V + 6 * (t->NJR > 0 ? t->NJR : 1) + 3 * (t->NOUT > 0 ? t->NOUT : 1) + 16 + 3 * (kThreads / 32)) * sizeof(float);
  HFR_CHECK_ARG(smem <= 227 * 1024, "geom_backward: mesh too large for shared memory");
  if (smem > 48 * 1024) cudaFuncSetAttribute(geom_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (a->face_rec && a->rec_partial) {
    const int64_t total = (int64_t)a->B * 3 * t->F;
    rec_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        *t, a->face_rec, reinterpret_cast<const uint32_t*>(a->raster_ws), hfr::ws_layout((int64_t)a->B * t->F), a->B, a->rec_partial);
    HFR_CHECK_LAUNCH("geom_backward (record gather)");
  }
  geom_bwd_kernel<<<a->B, kThreads, smem, (cudaStream_t)stream>>>(*t, *a);
  HFR_CHECK_LAUNCH("geom_backward");
  return HFR_OK;
}

extern "C" int hfr_face_verts_forward(const HfrTopology* t, const HfrFaceVertsArgs* a, void* stream) {
  HFR_CHECK_ARG(a && a->B >= 0, "face_verts_forward: null argument");
  if (a->B == 0) return HFR_OK;
  if (int rc = check_topo(t, 0)) return rc;
  HFR_CHECK_ARG(a->verts && a->face_verts, "face_verts_forward: null pointer");
  const int64_t total = (int64_t)a->B * t->F * 3;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  face_verts_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*t, a->B, a->verts, a->face_verts);
  HFR_CHECK_LAUNCH("face_verts_forward");
  return HFR_OK;
}

extern "C" int hfr_face_verts_backward(const HfrTopology* t, const HfrFaceVertsArgs* a, void* stream) {
  HFR_CHECK_ARG(a && a->B >= 0, "face_verts_backward: null argument");
  if (a->B == 0) return HFR_OK;
  if (int rc = check_topo(t, 0)) return rc;
  HFR_CHECK_ARG(a->g_face_verts && a->g_verts, "face_verts_backward: null pointer");
  const int64_t total = (int64_t)a->B * t->V;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  face_verts_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*t, a->B, a->g_face_verts, a->g_verts);
  HFR_CHECK_LAUNCH("face_verts_backward");
  return HFR_OK;
}

// Backward of shading + blending (+ optionally the rasterizer) for sm_100a.
//
// Differentiates HardPhongShader / SoftPhongShader / SoftSilhouetteShader (PyTorch3D; constructed at
// models_res_nimble.py:79-96, lights :187-190) and, on the fused path, rasterize_meshes_backward
// (reached by loss.backward(), train_hrnet.py:112) in ONE pass over the Fragments.
//
// Work decomposition
//   grid = (tiles_x, tiles_y, N), 256 threads = one 16x16 pixel tile, a warp owns an 8x4 block
//   (the same mapping as the forward rasterizer), so the lanes of a warp see few distinct faces.
//   Per fragment slot k the 27 per-face gradient components (3 corners x {ndc, view position,
//   vertex normal} x xyz) are first summed over the lanes that hit the SAME face with a
//   transposed butterfly (31 shuffles, after which lane j owns component j) and only then sent
//   to memory: one RED per (warp, face, component) instead of one per (pixel, component).
//   The heavy per-fragment code exists once (runtime loop over k, register arrays read through
//   select chains), which keeps the kernel inside the instruction cache.
#include "common.cuh"
#include "raster_math.cuh"
#include "shade_pixel.cuh"

namespace hfr {

constexpr int kBwdThreads = 256, kBwdTileW = 16, kBwdTileH = 16;

template <int KMAX>
__device__ __forceinline__ float selk(const float (&a)[KMAX], int k) {
  float r = a[0];
#pragma unroll
  for (int i = 1; i < KMAX; ++i) r = (k == i) ? a[i] : r;
  return r;
}
template <int KMAX>
__device__ __forceinline__ int selk_i(const int (&a)[KMAX], int k) {
  int r = a[0];
#pragma unroll
  for (int i = 1; i < KMAX; ++i) r = (k == i) ? a[i] : r;
  return r;
}

template <int KMAX>
__global__ void __launch_bounds__(kBwdThreads) shade_bwd_kernel(HfrShadeBwdArgs a) {
  __shared__ float s_light[kBwdThreads / 32][6];
  // per-warp staging of the 27 per-fragment components, pitch 33: lane L writes column L (bank j+L),
  // lane j later sums row j over the lanes of one face group (bank j+m) - both conflict-free
  __shared__ float s_red[kBwdThreads / 32][27][33];
  const HfrShadeFwdArgs& f = a.f;
  const HfrShadeParams& P = f.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, K = P.K, V = P.V;
  const int xi = blockIdx.x * kBwdTileW + (warp & 1) * 8 + (lane & 7);
  const int yi = blockIdx.y * kBwdTileH + (warp >> 1) * 4 + (lane >> 3);
  const bool active = xi < P.W && yi < P.H;
  const bool phong = P.shade == HFR_SHADE_PHONG_UV;
  const int kshade = phong ? (P.blend == HFR_BLEND_SOFTMAX ? K : 1) : 0;
  const size_t pix = ((size_t)n * P.H + yi) * P.W + xi;
  const bool dense = a.g_bary || a.g_zbuf || a.g_dists;

  // ---- fragments of this pixel -----------------------------------------------------------
  int fl[KMAX];            // face id within the mesh, -1 = empty slot
  float z[KMAX], d[KMAX];
  unsigned vmask = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { fl[k] = -1; z[k] = -1.f; d[k] = -1.f; }
  if (active) {
    const int64_t* __restrict__ ip = f.pix_to_face + pix * K;
    if (K == KMAX && (KMAX % 2) == 0) {
#pragma unroll
      for (int k = 0; k < KMAX; k += 2) {
        const longlong2 q = __ldg(reinterpret_cast<const longlong2*>(ip + k));
        fl[k] = q.x >= 0 ? (int)(q.x - (int64_t)n * P.F) : -1;
        fl[k + 1] = q.y >= 0 ? (int)(q.y - (int64_t)n * P.F) : -1;
      }
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) { const int64_t q = __ldg(ip + k); fl[k] = q >= 0 ? (int)(q - (int64_t)n * P.F) : -1; }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) vmask |= (fl[k] >= 0 ? 1u : 0u) << k;
  }
  const bool any = vmask != 0;
  const unsigned warp_any = __ballot_sync(0xffffffffu, any);

  float acc_dhat[3] = {0.f, 0.f, 0.f}, acc_lcol[3] = {0.f, 0.f, 0.f};
  float dhat[3] = {0.f, 0.f, 0.f}, dlen = 1.f, lcol[3] = {0.f, 0.f, 0.f};
  if (phong) {
    light_dir_hat(f, n, dhat, &dlen);
    lcol[0] = __ldg(f.light_color + 3 * n); lcol[1] = __ldg(f.light_color + 3 * n + 1); lcol[2] = __ldg(f.light_color + 3 * n + 2);
  }

  if (warp_any) {
    float g_colors[KMAX * 3], g_z[KMAX], g_d[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { g_z[k] = 0.f; g_d[k] = 0.f; g_colors[3 * k] = g_colors[3 * k + 1] = g_colors[3 * k + 2] = 0.f; }
    if (any) {
      if (K == KMAX && (KMAX % 4) == 0) {
#pragma unroll
        for (int k = 0; k < KMAX; k += 4) {
          const float4 zq = __ldg(reinterpret_cast<const float4*>(f.zbuf + pix * K + k));
          const float4 dq = __ldg(reinterpret_cast<const float4*>(f.dists + pix * K + k));
          z[k] = zq.x; z[k + 1] = zq.y; z[k + 2] = zq.z; z[k + 3] = zq.w;
          d[k] = dq.x; d[k + 1] = dq.y; d[k + 2] = dq.z; d[k + 3] = dq.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) { z[k] = __ldg(f.zbuf + pix * K + k); d[k] = __ldg(f.dists + pix * K + k); }
      }
      // forward colours of the shaded slots (recomputed: cheaper than storing K*3 floats per pixel)
      float colors[KMAX * 3];
#pragma unroll
      for (int k = 0; k < KMAX * 3; ++k) colors[k] = 1.0f;
#pragma unroll 1
      for (int k = 0; k < kshade; ++k) {
        if (!((vmask >> k) & 1u)) continue;
        const float* __restrict__ bp = f.bary + (pix * K + k) * 3;
        const float bc[3] = {__ldg(bp), __ldg(bp + 1), __ldg(bp + 2)};
        FragGeom g;
        gather_frag(f, n, selk_i<KMAX>(fl, k), g);
        HfrTexTap tap; HfrPhongCtx ctx; float texel[3], col[3];
        shade_fragment(f, n, g, bc, dhat, lcol, col, &tap, &ctx, texel);
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk)
          if (kk == k) { colors[3 * kk] = col[0]; colors[3 * kk + 1] = col[1]; colors[3 * kk + 2] = col[2]; }
      }
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.g_image + pix * 4));
      const float g_rgba[4] = {g4.x, g4.y, g4.z, g4.w};
      bool valid[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) valid[k] = (vmask >> k) & 1u;
      hfr_blend_bwd<KMAX>(P, K, valid, z, d, colors, g_rgba, g_colors, g_z, g_d);
    }
    const float xf = hfr_pix_to_ndc(P.W - 1 - xi, P.W, P.H), yf = hfr_pix_to_ndc(P.H - 1 - yi, P.H, P.W);
    const size_t tbase = (P.tex_n == 1 ? 0 : (size_t)n * P.tex_h * P.tex_w * 3);

    // ---- per fragment slot: differentiate, reduce per face inside the warp, scatter ----------
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      const bool vk = (vmask >> k) & 1u;
      unsigned todo = __ballot_sync(0xffffffffu, vk);
      if (dense && active && !vk) {
        if (a.g_bary) { float* o = a.g_bary + (pix * K + k) * 3; o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
        if (a.g_zbuf) a.g_zbuf[pix * K + k] = 0.f;
        if (a.g_dists) a.g_dists[pix * K + k] = 0.f;
      }
      if (!todo) continue;
      const int face = vk ? selk_i<KMAX>(fl, k) : -1;
      float v27[27];
#pragma unroll
      for (int i = 0; i < 27; ++i) v27[i] = 0.f;
      int vid[3] = {0, 0, 0};
      if (vk) {
        const float* __restrict__ bp = f.bary + (pix * K + k) * 3;
        const float bc[3] = {__ldg(bp), __ldg(bp + 1), __ldg(bp + 2)};
        const float gz = selk<KMAX>(g_z, k), gd = selk<KMAX>(g_d, k);
        float g_bc[3] = {0.f, 0.f, 0.f};
        if (k < kshade) {
          float gcol[3];
          {
            float c0[KMAX], c1[KMAX], c2[KMAX];
#pragma unroll
            for (int kk = 0; kk < KMAX; ++kk) { c0[kk] = g_colors[3 * kk]; c1[kk] = g_colors[3 * kk + 1]; c2[kk] = g_colors[3 * kk + 2]; }
            gcol[0] = selk<KMAX>(c0, k); gcol[1] = selk<KMAX>(c1, k); gcol[2] = selk<KMAX>(c2, k);
          }
          FragGeom g;
          gather_frag(f, n, face, g);
          vid[0] = g.vid[0]; vid[1] = g.vid[1]; vid[2] = g.vid[2];
          HfrTexTap tap; HfrPhongCtx ctx; float texel[3], col[3];
          shade_fragment(f, n, g, bc, dhat, lcol, col, &tap, &ctx, texel);
          float gP[3], gNn[3], gtex[3];
          hfr_phong_bwd(P, dhat, lcol, texel, &ctx, gcol, gP, gNn, gtex, acc_dhat, acc_lcol);
          float gu = 0.f, gv = 0.f;
          hfr_tex_uv_grad(f.texture + tbase, &tap, gtex, &gu, &gv);
          if (a.g_texture) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (tap.idx[q] >= 0) {
                float* dst = a.g_texture + tbase + (size_t)tap.idx[q] * 3;
                atomicAdd(dst, tap.w[q] * gtex[0]); atomicAdd(dst + 1, tap.w[q] * gtex[1]); atomicAdd(dst + 2, tap.w[q] * gtex[2]);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            g_bc[i] = gP[0] * g.X[3 * i] + gP[1] * g.X[3 * i + 1] + gP[2] * g.X[3 * i + 2] +
                      gNn[0] * g.Nv[3 * i] + gNn[1] * g.Nv[3 * i + 1] + gNn[2] * g.Nv[3 * i + 2] +
                      gu * g.uv[2 * i] + gv * g.uv[2 * i + 1];
#pragma unroll
            for (int c = 0; c < 3; ++c) { v27[9 * i + 3 + c] = bc[i] * gP[c]; v27[9 * i + 6 + c] = bc[i] * gNn[c]; }
          }
        } else if (a.g_verts_ndc) {
          vid[0] = __ldg(f.faces + 3 * face); vid[1] = __ldg(f.faces + 3 * face + 1); vid[2] = __ldg(f.faces + 3 * face + 2);
        }
        if (a.g_bary) { float* o = a.g_bary + (pix * K + k) * 3; o[0] = g_bc[0]; o[1] = g_bc[1]; o[2] = g_bc[2]; }
        if (a.g_zbuf) a.g_zbuf[pix * K + k] = gz;
        if (a.g_dists) a.g_dists[pix * K + k] = gd;
        if (a.g_verts_ndc) {
          float vv[9], gvv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float* __restrict__ src = a.verts_ndc + ((size_t)n * V + vid[i]) * 3;
            vv[3 * i] = __ldg(src); vv[3 * i + 1] = __ldg(src + 1); vv[3 * i + 2] = __ldg(src + 2);
          }
          hfr_raster_eval_bwd(xf, yf, vv, a.perspective_correct, a.clip_barycentric, g_bc, gz, gd, gvv);
#pragma unroll
          for (int i = 0; i < 3; ++i) { v27[9 * i] = gvv[3 * i]; v27[9 * i + 1] = gvv[3 * i + 1]; v27[9 * i + 2] = gvv[3 * i + 2]; }
        }
      }
      if (!(a.g_verts_ndc || (k < kshade && (a.g_verts_view || a.g_vnormals)))) continue;
      if (P.blend == HFR_BLEND_HARD && k >= kshade) continue;   // nothing flows through the hidden slots
      // segmented reduction: stage the components in shared memory, then per distinct face among the
      // lanes of this warp lane j sums component j over the group's lanes and issues one RED
      if (vk) {
#pragma unroll
        for (int i = 0; i < 27; ++i) s_red[warp][i][lane] = v27[i];
      }
      __syncwarp();
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lf = __shfl_sync(0xffffffffu, face, leader);
        const unsigned grp = __ballot_sync(0xffffffffu, vk && face == lf);
        todo &= ~grp;
        const int l0 = __shfl_sync(0xffffffffu, vid[0], leader), l1 = __shfl_sync(0xffffffffu, vid[1], leader),
                  l2 = __shfl_sync(0xffffffffu, vid[2], leader);
        if (lane < 27) {
          float total = 0.f;
          unsigned g2 = grp;
          while (g2) {
            const int m = __ffs(g2) - 1;
            g2 &= g2 - 1;
            total += s_red[warp][lane][m];
          }
          if (total != 0.f) {
            const int corner = lane / 9, r = lane - 9 * corner, which = r / 3, c = r - 3 * which;
            const int vv = corner == 0 ? l0 : (corner == 1 ? l1 : l2);
            float* base = which == 0 ? a.g_verts_ndc : (which == 1 ? a.g_verts_view : a.g_vnormals);
            if (base) atomicAdd(base + ((size_t)n * V + vv) * 3 + c, total);
          }
        }
      }
      __syncwarp();
    }
  }
  else if (dense && active) {   // no fragment in this warp: the dense per-fragment gradients are zero
    for (int k = 0; k < K; ++k) {
      if (a.g_bary) { float* o = a.g_bary + (pix * K + k) * 3; o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
      if (a.g_zbuf) a.g_zbuf[pix * K + k] = 0.f;
      if (a.g_dists) a.g_dists[pix * K + k] = 0.f;
    }
  }
  // ---- per-sample light gradients: block reduce, one atomic per component per CTA ---------------
  if (phong && (a.g_light_dir || a.g_light_color)) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { acc_dhat[c] = warp_sum(acc_dhat[c]); acc_lcol[c] = warp_sum(acc_lcol[c]); }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { s_light[warp][c] = acc_dhat[c]; s_light[warp][3 + c] = acc_lcol[c]; }
    }
    __syncthreads();
    if (tid == 0) {
      float t[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int w = 0; w < kBwdThreads / 32; ++w)
#pragma unroll
        for (int c = 0; c < 6; ++c) t[c] += s_light[w][c];
      float gd[3];
      hfr_normalize_eps_bwd(dhat, dlen, t, gd);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.g_light_dir && gd[c] != 0.f) atomicAdd(a.g_light_dir + 3 * n + c, gd[c]);
        if (a.g_light_color && t[3 + c] != 0.f) atomicAdd(a.g_light_color + 3 * n + c, t[3 + c]);
      }
    }
  }
}

int check_shade(const HfrShadeFwdArgs* a, const char* who);

}  // namespace hfr

extern "C" int hfr_shade_backward(const HfrShadeBwdArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a, "shade_backward: null args");
  if (int rc = check_shade(&a->f, "shade_backward")) return rc;
  if (a->f.p.N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->g_image, "shade_backward: null g_image");
  HFR_CHECK_ARG(!a->g_verts_ndc || (a->verts_ndc && a->f.faces && a->f.p.F > 0 && a->f.p.V > 0),
                "shade_backward: fused raster backward needs verts_ndc and faces");
  const HfrShadeParams& p = a->f.p;
  dim3 grid((p.W + kBwdTileW - 1) / kBwdTileW, (p.H + kBwdTileH - 1) / kBwdTileH, p.N);
  cudaStream_t st = (cudaStream_t)stream;
  if (p.K == 1) shade_bwd_kernel<1><<<grid, kBwdThreads, 0, st>>>(*a);
  else if (p.K == 2) shade_bwd_kernel<2><<<grid, kBwdThreads, 0, st>>>(*a);
  else if (p.K <= 4) shade_bwd_kernel<4><<<grid, kBwdThreads, 0, st>>>(*a);
  else if (p.K <= 8) shade_bwd_kernel<8><<<grid, kBwdThreads, 0, st>>>(*a);
  else shade_bwd_kernel<16><<<grid, kBwdThreads, 0, st>>>(*a);
  HFR_CHECK_LAUNCH("shade_backward");
  return HFR_OK;
}

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

#ifndef HFR_BWD_THREADS
#define HFR_BWD_THREADS 128
#endif
// CTA = 16 x (threads/16) pixels; the tile box from the rasterizer is in 16x16 tiles
constexpr int kBwdThreads = HFR_BWD_THREADS, kBwdTileW = 16, kBwdTileH = kBwdThreads / 16;

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

#ifndef HFR_BWD_WARP_LIGHT
#define HFR_BWD_WARP_LIGHT 0
#endif
// Resident CTAs per SM (register cap = 65536 / (threads x CTAs)).  Measured on B200 (C2, K=4): 8 CTAs x 64 registers
// 377 us, 4 x 128 registers 368 us, 3 x 168 411 us - the per-slot loop is latency-bound and wants both warps and
// registers.  K=1 at 672^2 (Fragments streaming dominates) prefers the 8-CTA setting: 406 vs 459 us.
#ifndef HFR_BWD_MINB
#define HFR_BWD_MINB 4
#endif
#ifndef HFR_BWD_MINB_K1
#define HFR_BWD_MINB_K1 8
#endif
#ifndef HFR_BWD_MINB_K1_PCA
#define HFR_BWD_MINB_K1_PCA HFR_BWD_MINB_K1   // texture PCA, K = 1 (C3): resident CTAs per SM of that instantiation
#endif
template <int KMAX, bool PCA>
__global__ void __launch_bounds__(kBwdThreads, (KMAX == 1 ? (PCA ? HFR_BWD_MINB_K1_PCA : HFR_BWD_MINB_K1) : (KMAX <= 4 ? HFR_BWD_MINB : 1))) shade_bwd_kernel(HfrShadeBwdArgs a) {
  __shared__ float s_light[kBwdThreads / 32][6];
  // per-warp staging of the 27 per-fragment components, pitch 33: lane L writes column L (bank j+L),
  // lane j later sums row j over the lanes of one face group (bank j+m) - both conflict-free
  __shared__ float s_red[kBwdThreads / 32][27][33];
  // per-thread fragment table [KMAX][4][threads]: face id, sigmoid prob, softmax exponent, prod_{j!=k}(1-p_j).
  // Indexed by the runtime slot k in the loop below (register arrays would need select chains).
  extern __shared__ float s_frag[];
  float* my_frag = s_frag + threadIdx.x;
#define FRAG(k, field) my_frag[((k) * 4 + (field)) * kBwdThreads]
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
  if (a.tile_box && !dense) {   // tile outside this mesh's footprint: no fragment, no gradient
    const uint4 bx = __ldg(reinterpret_cast<const uint4*>(a.tile_box) + n);
    const int tx = blockIdx.x, ty = (blockIdx.y * kBwdTileH) >> 4;
    if (tx < (int)bx.x || tx > 255 - (int)bx.y || ty < (int)bx.z || ty > 255 - (int)bx.w) return;
  }
  // texture PCA: this sample's coefficients, staged once per CTA and zero padded to a multiple of 4
  __shared__ float s_tp[PCA ? HFR_MAX_TEX_PCA : 1];
  if (PCA) {
    if (tid < HFR_MAX_TEX_PCA) s_tp[tid] = tid < P.tex_pca ? f.tex_params[(size_t)n * P.tex_pca + tid] : 0.0f;
    __syncthreads();
  }

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
    // ---- per-pixel blend state.  The blend is differentiated WITHOUT the fragments' colours: the
    //      forward image (rgb = (sum_k w_k c_k + delta bg) / den) is an input, so den, d/d(rgb) and the
    //      coupling through z_max follow from (z, d) and the stored pixel alone; a fragment's own
    //      colour is only needed when that fragment is shaded in the loop below.
    float prob[KMAX], wexp[KMAX], others[KMAX];   // sigmoid prob, exp((zinv-zmax)/gamma), prod_{j!=k}(1-p_j)
    float gnum[3] = {0.f, 0.f, 0.f}, gden = 0.f, gzmax = 0.f, g_alpha = 0.f;
    int kmax = -1;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { prob[k] = 0.f; wexp[k] = 0.f; others[k] = 1.f; }
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
      float4 g4;
      if (a.pool_aa > 1) {   // gradient of the pooled image: avg_pool2d backward folded into the load
        const int aa = a.pool_aa, Wp = P.W / aa, Hp = P.H / aa;
        g4 = __ldg(reinterpret_cast<const float4*>(a.g_image + (((size_t)n * Hp + yi / aa) * Wp + xi / aa) * 4));
        const float inv = 1.0f / (float)(aa * aa);
        g4.x *= inv; g4.y *= inv; g4.z *= inv; g4.w = a.pool_binarize ? 0.0f : g4.w * inv;
      } else {
        g4 = __ldg(reinterpret_cast<const float4*>(a.g_image + pix * 4));
      }
      g_alpha = g4.w;
      if (P.blend == HFR_BLEND_HARD) {
        gnum[0] = g4.x; gnum[1] = g4.y; gnum[2] = g4.z;     // g_colour of slot 0, nothing else flows
      } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if ((vmask >> k) & 1u) prob[k] = hfr_sigmoid(HFR_FDIV(-d[k], P.sigma));
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          float o = 1.0f;
#pragma unroll
          for (int jj = 0; jj < KMAX; ++jj)
            if (jj != k) o *= (1.0f - prob[jj]);
          others[k] = o;
        }
        if (P.blend == HFR_BLEND_SIGMOID_ALPHA) {
          gnum[0] = g4.x; gnum[1] = g4.y; gnum[2] = g4.z;   // colour of slot 0 passes straight through
        } else {
          const float eps = 1e-10f, zr = P.zfar - P.znear;
          float zinv[KMAX], zmax_raw = 0.0f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            zinv[k] = 0.0f;
            if ((vmask >> k) & 1u) zinv[k] = (P.zfar - z[k]) / zr;   // IEEE divide, as the forward
            if (k < K && (k == 0 || zinv[k] > zmax_raw)) { zmax_raw = zinv[k]; kmax = k; }
          }
          const float zmax = fmaxf(zmax_raw, eps);
          float wsum = 0.0f;
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            if (k < K) { wexp[k] = HFR_EXP(HFR_FDIV(zinv[k] - zmax, P.gamma)); wsum += prob[k] * wexp[k]; }
          }
          const float dexp = HFR_EXP(HFR_FDIV(eps - zmax, P.gamma));
          const float delta = fmaxf(dexp, eps);
          const float den = wsum + delta, iden = HFR_RCP(den);
          const float4 im = __ldg(reinterpret_cast<const float4*>(f.image + pix * 4));
          const float rgb[3] = {im.x, im.y, im.z}, gin[3] = {g4.x, g4.y, g4.z};
          float gdelta = 0.0f, gacc = 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            gnum[c] = gin[c] * iden;
            gden -= gnum[c] * rgb[c];
            gdelta += gnum[c] * P.background[c];
            gacc += gnum[c] * (rgb[c] * den - delta * P.background[c]);   // gnum . sum_k w_k c_k
          }
          gdelta += gden;
          // sum_k d/d(zinv_k - zmax) = (gden * wsum + gnum . acc) / gamma
          gzmax = -HFR_FDIV(gden * wsum + gacc, P.gamma);
          if (dexp >= eps) gzmax -= HFR_FDIV(gdelta * delta, P.gamma);
          if (!(zmax_raw >= eps)) kmax = -1;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      FRAG(k, 0) = __int_as_float(fl[k]); FRAG(k, 1) = prob[k]; FRAG(k, 2) = wexp[k]; FRAG(k, 3) = others[k];
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
      if (P.blend == HFR_BLEND_HARD && k > 0) {   // hidden slots of a hard blend carry no gradient
        if (dense && vk) {
          if (a.g_bary) { float* o = a.g_bary + (pix * K + k) * 3; o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
          if (a.g_zbuf) a.g_zbuf[pix * K + k] = 0.f;
          if (a.g_dists) a.g_dists[pix * K + k] = 0.f;
        }
        continue;
      }
      const int face = vk ? __float_as_int(FRAG(k, 0)) : -1;
      float v27[27];
#pragma unroll
      for (int i = 0; i < 27; ++i) v27[i] = 0.f;
      int vid[3] = {0, 0, 0};
      HfrTexTap tap_k;                       // kept for the texture-coefficient gradient after the fragment block
      float gtex_k[3] = {0.f, 0.f, 0.f};
      tap_k.idx[0] = tap_k.idx[1] = tap_k.idx[2] = tap_k.idx[3] = -1;
      tap_k.w[0] = tap_k.w[1] = tap_k.w[2] = tap_k.w[3] = 0.f;
      if (vk) {
        const float* __restrict__ bp = f.bary + (pix * K + k) * 3;
        const float bc[3] = {__ldg(bp), __ldg(bp + 1), __ldg(bp + 2)};
        const float pk = FRAG(k, 1), ek = FRAG(k, 2);
        float gprob = g_alpha * FRAG(k, 3), gz = 0.f;
        float g_bc[3] = {0.f, 0.f, 0.f};
        float col[3] = {1.0f, 1.0f, 1.0f}, gcol[3];
        FragGeom g;
        HfrTexTap tap; HfrPhongCtx ctx; float texel[3];
        float duv[6];   // PCA: uv-derivative sums of the texel fetch (one visit of the taps)
        if (k < kshade) {
          gather_frag(f, n, face, g);
          vid[0] = g.vid[0]; vid[1] = g.vid[1]; vid[2] = g.vid[2];
          shade_fragment<PCA>(f, n, g, bc, dhat, lcol, col, &tap, &ctx, texel, PCA ? s_tp : nullptr, PCA ? duv : nullptr);
        }
        if (P.blend == HFR_BLEND_SOFTMAX) {
          const float wk = pk * ek;
          const float gw = gden + gnum[0] * col[0] + gnum[1] * col[1] + gnum[2] * col[2];
          gcol[0] = wk * gnum[0]; gcol[1] = wk * gnum[1]; gcol[2] = wk * gnum[2];
          gprob += gw * ek;
          const float gzinv = HFR_FDIV(gw * wk, P.gamma) + (k == kmax ? gzmax : 0.f);
          gz = HFR_FDIV(-gzinv, P.zfar - P.znear);
        } else {
          gcol[0] = gnum[0]; gcol[1] = gnum[1]; gcol[2] = gnum[2];
        }
        if (k < kshade) {
          float gP[3], gNn[3], gtex[3];
          if (P.light_point) {   // the direction depends on the fragment's position: location - P
            float gdh[3] = {0.f, 0.f, 0.f}, gd[3];
            hfr_phong_bwd(P, ctx.lhat, lcol, texel, &ctx, gcol, gP, gNn, gtex, gdh, acc_lcol);
            hfr_normalize_eps_bwd(ctx.lhat, ctx.llen, gdh, gd);
#pragma unroll
            for (int c = 0; c < 3; ++c) { acc_dhat[c] += gd[c]; gP[c] -= gd[c]; }
          } else {
            hfr_phong_bwd(P, dhat, lcol, texel, &ctx, gcol, gP, gNn, gtex, acc_dhat, acc_lcol);
          }
          float gu = 0.f, gv = 0.f;
          if (PCA) hfr_tex_uv_grad_d(&tap, duv, duv + 3, gtex, &gu, &gv);
          else hfr_tex_uv_grad(tex_source<PCA>(f, n), &tap, gtex, &gu, &gv);
          if (PCA && a.g_tex_params) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { tap_k.idx[q] = tap.idx[q]; tap_k.w[q] = tap.w[q]; }
            gtex_k[0] = gtex[0]; gtex_k[1] = gtex[1]; gtex_k[2] = gtex[2];
          }
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
        const float gd = P.blend == HFR_BLEND_HARD ? 0.f : HFR_FDIV(-gprob * pk * (1.0f - pk), P.sigma);
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
      if (PCA && a.g_tex_params && k < kshade) {
        // d(loss)/d(texture coefficients): warp-reduce each component over the lanes of this slot, one RED per warp
        const HfrTexSrc tsrc = tex_source<true>(f, n, s_tp);
        for (int c0 = 0; c0 < P.tex_pca; c0 += 4) {
          float t4[4] = {0.f, 0.f, 0.f, 0.f};
          if (vk) hfr_tex_param_grad4(tsrc, &tap_k, gtex_k, c0, t4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float t = warp_sum(t4[i]);
            if (lane == 0 && c0 + i < P.tex_pca && t != 0.0f) atomicAdd(a.g_tex_params + (size_t)n * P.tex_pca + c0 + i, t);
          }
        }
      }
      if (!(a.g_verts_ndc || (k < kshade && (a.g_verts_view || a.g_vnormals)))) continue;
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
#if HFR_BWD_WARP_LIGHT
    // one set of atomics per warp, no CTA barrier: warps of a partially covered tile finish independently
    if (lane == 0 && warp_any) {
      const float t[6] = {acc_dhat[0], acc_dhat[1], acc_dhat[2], acc_lcol[0], acc_lcol[1], acc_lcol[2]};
      float gd[3] = {t[0], t[1], t[2]};
      if (!P.light_point) hfr_normalize_eps_bwd(dhat, dlen, t, gd);   // linear in t, so per-warp application sums to the same gradient
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.g_light_dir && gd[c] != 0.f) atomicAdd(a.g_light_dir + 3 * n + c, gd[c]);
        if (a.g_light_color && t[3 + c] != 0.f) atomicAdd(a.g_light_color + 3 * n + c, t[3 + c]);
      }
    }
#else
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
      float gd[3] = {t[0], t[1], t[2]};
      if (!P.light_point) hfr_normalize_eps_bwd(dhat, dlen, t, gd);   // PointLights: t already is d/d(location)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.g_light_dir && gd[c] != 0.f) atomicAdd(a.g_light_dir + 3 * n + c, gd[c]);
        if (a.g_light_color && t[3 + c] != 0.f) atomicAdd(a.g_light_color + 3 * n + c, t[3 + c]);
      }
    }
#endif
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
  HFR_CHECK_ARG(a->f.p.blend != HFR_BLEND_SOFTMAX || a->f.image, "shade_backward: the softmax blend needs the forward image");
  HFR_CHECK_ARG(!a->g_verts_ndc || (a->verts_ndc && a->f.faces && a->f.p.F > 0 && a->f.p.V > 0),
                "shade_backward: fused raster backward needs verts_ndc and faces");
  HFR_CHECK_ARG(!a->g_tex_params || a->f.p.tex_pca > 0, "shade_backward: g_tex_params needs a PCA texture (tex_pca > 0)");
  HFR_CHECK_ARG(a->pool_aa <= 1 || (a->pool_aa <= 16 && a->f.p.H % a->pool_aa == 0 && a->f.p.W % a->pool_aa == 0),
                "shade_backward: image size must be a multiple of pool_aa (<= 16)");
  const HfrShadeParams& p = a->f.p;
  dim3 grid((p.W + kBwdTileW - 1) / kBwdTileW, (p.H + kBwdTileH - 1) / kBwdTileH, p.N);
  cudaStream_t st = (cudaStream_t)stream;
#define HFR_LAUNCH_BWD(KM)                                                                              \
  do {                                                                                                  \
    const size_t sm = (size_t)(KM) * 4 * kBwdThreads * sizeof(float);                                   \
    if (p.tex_pca > 0) {                                                                                \
      cudaFuncSetAttribute(shade_bwd_kernel<KM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);   \
      shade_bwd_kernel<KM, true><<<grid, kBwdThreads, sm, st>>>(*a);                                    \
    } else {                                                                                            \
      cudaFuncSetAttribute(shade_bwd_kernel<KM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);  \
      shade_bwd_kernel<KM, false><<<grid, kBwdThreads, sm, st>>>(*a);                                   \
    }                                                                                                   \
  } while (0)
  if (p.K == 1) HFR_LAUNCH_BWD(1);
  else if (p.K == 2) HFR_LAUNCH_BWD(2);
  else if (p.K <= 4) HFR_LAUNCH_BWD(4);
  else if (p.K <= 8) HFR_LAUNCH_BWD(8);
  else HFR_LAUNCH_BWD(16);
#undef HFR_LAUNCH_BWD
  HFR_CHECK_LAUNCH("shade_backward");
  return HFR_OK;
}

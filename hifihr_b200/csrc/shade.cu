// Shading + blending kernels for sm_100a (forward, backward) and the fused rasterize+shade forward.
//
// Replaces HardPhongShader / SoftPhongShader / SoftSilhouetteShader (PyTorch3D; constructed at
// models_res_nimble.py:79-96, lights at :187-190) — texture sampling, attribute interpolation,
// Phong lighting and the three blend modes in ONE pass over the fragments instead of ~45
// full-resolution elementwise kernels.  The backward optionally applies the rasterizer backward
// in the same kernel (fused path), so d(image)->d(vertices) reads Fragments exactly once.
#include "common.cuh"
#include "raster_tile.cuh"
#include "shade_pixel.cuh"

namespace hfr {

template <int KMAX>
__global__ void __launch_bounds__(256) shade_fwd_kernel(HfrShadeFwdArgs a) {
  const int n = blockIdx.y, K = a.p.K;
  const int HW = a.p.H * a.p.W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const size_t pix = (size_t)n * HW + p;
  int64_t id[KMAX];
  float z[KMAX], d[KMAX], b[KMAX * 3];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    id[k] = -1; z[k] = -1.f; d[k] = -1.f; b[3 * k] = b[3 * k + 1] = b[3 * k + 2] = -1.f;
    if (k < K) {
      id[k] = a.pix_to_face[pix * K + k];
      z[k] = a.zbuf[pix * K + k];
      d[k] = a.dists[pix * K + k];
      b[3 * k] = a.bary[(pix * K + k) * 3]; b[3 * k + 1] = a.bary[(pix * K + k) * 3 + 1]; b[3 * k + 2] = a.bary[(pix * K + k) * 3 + 2];
    }
  }
  float rgba[4];
  shade_pixel<KMAX>(a, n, id, z, d, b, rgba);
  *reinterpret_cast<float4*>(a.image + pix * 4) = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
}

// Fused rasterize + shade forward: Fragments and the RGBA image leave the SM in the same pass.
template <int KMAX>
__global__ void __launch_bounds__(kRasterThreads) raster_shade_fwd_kernel(HfrRasterArgs r, HfrShadeFwdArgs s,
                                                                          const uint32_t* __restrict__ ranges) {
  __shared__ RasterSmem sm;
  const PixelCtx c = make_pixel_ctx(r.H, r.W);
  TopK<KMAX> top;
  raster_tile<KMAX>(r, ranges, sm, c.n, c.tx, c.ty, c.xf, c.yf, c.pix_active, c.warp_active, c.wx_lo, c.wx_hi,
                    c.wy_lo, c.wy_hi, top);
  if (c.pix_active) {
    int64_t id[KMAX];
    float z[KMAX], d[KMAX], b[KMAX * 3];
    compute_fragments<KMAX>(r, c.xf, c.yf, top, id, z, d, b);
    const size_t pix = ((size_t)c.n * r.H + c.yi) * r.W + c.xi;
    store_fragments<KMAX>(r, pix, id, z, d, b);
    float rgba[4];
    shade_pixel<KMAX>(s, c.n, id, z, d, b, rgba);
    *reinterpret_cast<float4*>(s.image + pix * 4) = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
  }
}

// Backward.  One thread per pixel; the colours are recomputed (cheaper than storing K*3 floats
// per pixel), the blend is differentiated, then every fragment's shading / interpolation /
// (optionally) rasterization is differentiated and scattered.
template <int KMAX>
__global__ void __launch_bounds__(256) shade_bwd_kernel(HfrShadeBwdArgs a) {
  const HfrShadeFwdArgs& f = a.f;
  const HfrShadeParams& P = f.p;
  const int n = blockIdx.y, K = P.K, HW = P.H * P.W, V = P.V;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < HW;
  const bool phong = P.shade == HFR_SHADE_PHONG_UV;
  float acc_dhat[3] = {0.f, 0.f, 0.f}, acc_lcol[3] = {0.f, 0.f, 0.f};
  float dhat[3] = {0.f, 0.f, 0.f}, dlen = 1.f, lcol[3] = {0.f, 0.f, 0.f};
  if (phong) {
    light_dir_hat(f, n, dhat, &dlen);
    lcol[0] = f.light_color[3 * n]; lcol[1] = f.light_color[3 * n + 1]; lcol[2] = f.light_color[3 * n + 2];
  }
  if (active) {
    const size_t pix = (size_t)n * HW + p;
    int64_t id[KMAX];
    float z[KMAX], d[KMAX], b[KMAX * 3], colors[KMAX * 3];
    bool valid[KMAX];
    bool any = false;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      id[k] = -1; z[k] = -1.f; d[k] = -1.f; b[3 * k] = b[3 * k + 1] = b[3 * k + 2] = -1.f;
      if (k < K) {
        id[k] = f.pix_to_face[pix * K + k];
        if (id[k] >= 0) {
          z[k] = f.zbuf[pix * K + k];
          d[k] = f.dists[pix * K + k];
          b[3 * k] = f.bary[(pix * K + k) * 3]; b[3 * k + 1] = f.bary[(pix * K + k) * 3 + 1]; b[3 * k + 2] = f.bary[(pix * K + k) * 3 + 2];
        }
      }
      valid[k] = k < K && id[k] >= 0;
      any = any || valid[k];
      colors[3 * k] = colors[3 * k + 1] = colors[3 * k + 2] = 1.0f;
    }
    const int kshade = P.blend == HFR_BLEND_SOFTMAX ? K : 1;
    if (any) {
      const float4 g4 = *reinterpret_cast<const float4*>(a.g_image + pix * 4);
      const float g_rgba[4] = {g4.x, g4.y, g4.z, g4.w};
      if (phong) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (valid[k] && k < kshade) {
            FragGeom g;
            gather_frag(f, n, (int)(id[k] - (int64_t)n * P.F), g);
            HfrTexTap tap; HfrPhongCtx ctx; float texel[3];
            shade_fragment(f, n, g, b + 3 * k, dhat, lcol, colors + 3 * k, &tap, &ctx, texel);
          }
        }
      }
      float g_colors[KMAX * 3], g_z[KMAX], g_d[KMAX];
      hfr_blend_bwd<KMAX>(P, K, valid, z, d, colors, g_rgba, g_colors, g_z, g_d);
      const int xi = p % P.W, yi = p / P.W;
      const float xf = hfr_pix_to_ndc(P.W - 1 - xi, P.W, P.H), yf = hfr_pix_to_ndc(P.H - 1 - yi, P.H, P.W);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (!valid[k]) continue;
        float g_bc[3] = {0.f, 0.f, 0.f};
        const int fl = (int)(id[k] - (int64_t)n * P.F);
        int vid[3];
        if (phong && k < kshade) {
          FragGeom g;
          gather_frag(f, n, fl, g);
          vid[0] = g.vid[0]; vid[1] = g.vid[1]; vid[2] = g.vid[2];
          HfrTexTap tap; HfrPhongCtx ctx; float texel[3], col[3];
          shade_fragment(f, n, g, b + 3 * k, dhat, lcol, col, &tap, &ctx, texel);
          float gP[3], gNn[3], gtex[3];
          hfr_phong_bwd(P, dhat, lcol, texel, &ctx, g_colors + 3 * k, gP, gNn, gtex, acc_dhat, acc_lcol);
          // texture: scatter to the 4 taps, and d(texel)/d(uv)
          const size_t tbase = (P.tex_n == 1 ? 0 : (size_t)n * P.tex_h * P.tex_w * 3);
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
          // interpolation: P = sum bc_i X_i, Nn = sum bc_i N_i, uv = sum bc_i uv_i
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            g_bc[i] = gP[0] * g.X[3 * i] + gP[1] * g.X[3 * i + 1] + gP[2] * g.X[3 * i + 2] +
                      gNn[0] * g.Nv[3 * i] + gNn[1] * g.Nv[3 * i + 1] + gNn[2] * g.Nv[3 * i + 2] +
                      gu * g.uv[2 * i] + gv * g.uv[2 * i + 1];
            const float bi = b[3 * k + i];
            if (a.g_verts_view) {
              float* dst = a.g_verts_view + ((size_t)n * V + g.vid[i]) * 3;
              atomicAdd(dst, bi * gP[0]); atomicAdd(dst + 1, bi * gP[1]); atomicAdd(dst + 2, bi * gP[2]);
            }
            if (a.g_vnormals) {
              float* dst = a.g_vnormals + ((size_t)n * V + g.vid[i]) * 3;
              atomicAdd(dst, bi * gNn[0]); atomicAdd(dst + 1, bi * gNn[1]); atomicAdd(dst + 2, bi * gNn[2]);
            }
          }
        } else {
          vid[0] = __ldg(f.faces + 3 * fl); vid[1] = __ldg(f.faces + 3 * fl + 1); vid[2] = __ldg(f.faces + 3 * fl + 2);
        }
        if (a.g_bary) { a.g_bary[(pix * K + k) * 3] = g_bc[0]; a.g_bary[(pix * K + k) * 3 + 1] = g_bc[1]; a.g_bary[(pix * K + k) * 3 + 2] = g_bc[2]; }
        if (a.g_zbuf) a.g_zbuf[pix * K + k] = g_z[k];
        if (a.g_dists) a.g_dists[pix * K + k] = g_d[k];
        if (a.g_verts_ndc) {
          float v[9], gvv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float* __restrict__ src = a.verts_ndc + ((size_t)n * V + vid[i]) * 3;
            v[3 * i] = __ldg(src); v[3 * i + 1] = __ldg(src + 1); v[3 * i + 2] = __ldg(src + 2);
          }
          hfr_raster_eval_bwd(xf, yf, v, a.perspective_correct, a.clip_barycentric, g_bc, g_z[k], g_d[k], gvv);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            float* dst = a.g_verts_ndc + ((size_t)n * V + vid[i]) * 3;
            if (gvv[3 * i] != 0.f) atomicAdd(dst, gvv[3 * i]);
            if (gvv[3 * i + 1] != 0.f) atomicAdd(dst + 1, gvv[3 * i + 1]);
            if (gvv[3 * i + 2] != 0.f) atomicAdd(dst + 2, gvv[3 * i + 2]);
          }
        }
      }
    } else {
      // empty pixel: dense grads (if requested) are zero
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          if (a.g_bary) { a.g_bary[(pix * K + k) * 3] = 0.f; a.g_bary[(pix * K + k) * 3 + 1] = 0.f; a.g_bary[(pix * K + k) * 3 + 2] = 0.f; }
          if (a.g_zbuf) a.g_zbuf[pix * K + k] = 0.f;
          if (a.g_dists) a.g_dists[pix * K + k] = 0.f;
        }
      }
    }
    // dense grads of invalid slots of a non-empty pixel
    if (any) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K && !valid[k]) {
          if (a.g_bary) { a.g_bary[(pix * K + k) * 3] = 0.f; a.g_bary[(pix * K + k) * 3 + 1] = 0.f; a.g_bary[(pix * K + k) * 3 + 2] = 0.f; }
          if (a.g_zbuf) a.g_zbuf[pix * K + k] = 0.f;
          if (a.g_dists) a.g_dists[pix * K + k] = 0.f;
        }
      }
    }
  }
  // per-sample light gradients: warp reduce, one atomic per warp
  if (phong && (a.g_light_dir || a.g_light_color)) {
    float gd[3];
    // d̂ = dir / max(|dir|, eps) is linear-isable after the sum: reduce g_dhat first
#pragma unroll
    for (int c = 0; c < 3; ++c) { acc_dhat[c] = warp_sum(acc_dhat[c]); acc_lcol[c] = warp_sum(acc_lcol[c]); }
    if ((threadIdx.x & 31) == 0) {
      hfr_normalize_eps_bwd(dhat, dlen, acc_dhat, gd);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.g_light_dir && gd[c] != 0.f) atomicAdd(a.g_light_dir + 3 * n + c, gd[c]);
        if (a.g_light_color && acc_lcol[c] != 0.f) atomicAdd(a.g_light_color + 3 * n + c, acc_lcol[c]);
      }
    }
  }
}

static int check_shade(const HfrShadeFwdArgs* a, const char* who) {
  HFR_CHECK_ARG(a, "%s: null args", who);
  const HfrShadeParams& p = a->p;
  HFR_CHECK_ARG(p.N >= 0 && p.H > 0 && p.W > 0 && p.K >= 1 && p.K <= HFR_MAX_K, "%s: bad dims", who);
  HFR_CHECK_ARG(p.blend >= 0 && p.blend <= 2 && p.shade >= 0 && p.shade <= 1, "%s: bad blend/shade mode", who);
  HFR_CHECK_ARG(p.blend == HFR_BLEND_HARD || (p.sigma > 0.f && p.gamma > 0.f), "%s: sigma/gamma must be > 0", who);
  if (p.N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->pix_to_face && a->zbuf && a->bary && a->dists, "%s: null fragments", who);
  if (p.shade == HFR_SHADE_PHONG_UV)
    HFR_CHECK_ARG(a->faces && a->verts_view && a->vnormals && a->faces_uvs && a->verts_uvs && a->texture &&
                      a->light_dir && a->light_color && p.F > 0 && p.V > 0 && p.tex_h > 0 && p.tex_w > 0 &&
                      (p.tex_n == 1 || p.tex_n == p.N),
                  "%s: phong/uv shading needs mesh, uv, texture and light pointers", who);
  return HFR_OK;
}

}  // namespace hfr

#define HFR_DISPATCH_K(K, CALL)            \
  do {                                     \
    if ((K) == 1) { CALL(1); }             \
    else if ((K) == 2) { CALL(2); }        \
    else if ((K) <= 4) { CALL(4); }        \
    else if ((K) <= 8) { CALL(8); }        \
    else { CALL(16); }                     \
  } while (0)

extern "C" int hfr_shade_forward(const HfrShadeFwdArgs* a, void* stream) {
  using namespace hfr;
  if (int rc = check_shade(a, "shade_forward")) return rc;
  HFR_CHECK_ARG(a->p.N == 0 || a->image, "shade_forward: null image");
  if (a->p.N == 0) return HFR_OK;
  dim3 grid((a->p.H * a->p.W + 255) / 256, a->p.N);
#define CALL(KM) shade_fwd_kernel<KM><<<grid, 256, 0, (cudaStream_t)stream>>>(*a)
  HFR_DISPATCH_K(a->p.K, CALL);
#undef CALL
  HFR_CHECK_LAUNCH("shade_forward");
  return HFR_OK;
}

extern "C" int hfr_shade_backward(const HfrShadeBwdArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a, "shade_backward: null args");
  if (int rc = check_shade(&a->f, "shade_backward")) return rc;
  if (a->f.p.N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->g_image, "shade_backward: null g_image");
  HFR_CHECK_ARG(!a->g_verts_ndc || (a->verts_ndc && a->f.faces && a->f.p.F > 0 && a->f.p.V > 0),
                "shade_backward: fused raster backward needs verts_ndc and faces");
  dim3 grid((a->f.p.H * a->f.p.W + 255) / 256, a->f.p.N);
#define CALL(KM) shade_bwd_kernel<KM><<<grid, 256, 0, (cudaStream_t)stream>>>(*a)
  HFR_DISPATCH_K(a->f.p.K, CALL);
#undef CALL
  HFR_CHECK_LAUNCH("shade_backward");
  return HFR_OK;
}

extern "C" int hfr_raster_shade_forward(const HfrRasterShadeArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a, "raster_shade_forward: null args");
  if (int rc = check_raster(&a->r, "raster_shade_forward")) return rc;
  HfrShadeFwdArgs s = a->s;
  s.pix_to_face = a->r.pix_to_face; s.zbuf = a->r.zbuf; s.bary = a->r.bary; s.dists = a->r.dists;
  if (int rc = check_shade(&s, "raster_shade_forward")) return rc;
  HFR_CHECK_ARG(s.p.N == a->r.N && s.p.H == a->r.H && s.p.W == a->r.W && s.p.K == a->r.K,
                "raster_shade_forward: raster / shade dims differ");
  HFR_CHECK_ARG(a->r.N == 0 || s.image, "raster_shade_forward: null image");
  if (a->r.N == 0) return HFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* ranges = reinterpret_cast<uint32_t*>(a->r.workspace);
  if (int rc = launch_raster_setup(a->r, ranges, st)) return rc;
  dim3 grid((a->r.W + kTileW - 1) / kTileW, (a->r.H + kTileH - 1) / kTileH, a->r.N);
#define CALL(KM) raster_shade_fwd_kernel<KM><<<grid, kRasterThreads, 0, st>>>(a->r, s, ranges)
  HFR_DISPATCH_K(a->r.K, CALL);
#undef CALL
  HFR_CHECK_LAUNCH("raster_shade_forward");
  return HFR_OK;
}

// Device-side per-pixel shading: gather the face attributes of the K fragments of one pixel,
// interpolate, sample the UV texture, light, blend.  Shared by the standalone shader kernels
// (shade.cu) and the fused rasterize+shade forward (raster_shade.cu).
#pragma once
#include "shade_math.cuh"

namespace hfr {

struct FragGeom {        // per-fragment gathered attributes
  int vid[3];            // vertex ids (within the mesh)
  float X[9], Nv[9];     // view-space positions / vertex normals of the 3 corners
  float uv[6];           // corner uvs
};

HFR_HD void gather_frag(const HfrShadeFwdArgs& a, int n, int fl, FragGeom& g) {
  const int V = a.p.V;
#if defined(__CUDA_ARCH__)
  if (a.face_attr) {   // one contiguous 112-byte record per (mesh, face): 7 independent 128-bit loads
    const float4* __restrict__ r4 = reinterpret_cast<const float4*>(a.face_attr + ((size_t)n * a.p.F + fl) * HFR_FACE_ATTR_FLOATS);
    float w[28];
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      const float4 q = __ldg(r4 + u);
      w[4 * u] = q.x; w[4 * u + 1] = q.y; w[4 * u + 2] = q.z; w[4 * u + 3] = q.w;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) { g.X[e] = w[e]; g.Nv[e] = w[9 + e]; }
#pragma unroll
    for (int e = 0; e < 6; ++e) g.uv[e] = w[18 + e];
#pragma unroll
    for (int e = 0; e < 3; ++e) g.vid[e] = __float_as_int(w[24 + e]);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int vid = HFR_LDG(a.faces + 3 * fl + i);
    g.vid[i] = vid;
    const float* __restrict__ x = a.verts_view + ((size_t)n * V + vid) * 3;
    const float* __restrict__ nn = a.vnormals + ((size_t)n * V + vid) * 3;
    g.X[3 * i] = HFR_LDG(x); g.X[3 * i + 1] = HFR_LDG(x + 1); g.X[3 * i + 2] = HFR_LDG(x + 2);
    g.Nv[3 * i] = HFR_LDG(nn); g.Nv[3 * i + 1] = HFR_LDG(nn + 1); g.Nv[3 * i + 2] = HFR_LDG(nn + 2);
    const int t = HFR_LDG(a.faces_uvs + 3 * fl + i);
    g.uv[2 * i] = HFR_LDG(a.verts_uvs + 2 * t); g.uv[2 * i + 1] = HFR_LDG(a.verts_uvs + 2 * t + 1);
  }
}

HFR_HD void interp3(const float* bc, const float* A, float* o) {
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = bc[0] * A[c] + bc[1] * A[3 + c] + bc[2] * A[6 + c];
}

// DirectionalLights: the sample's unit light direction.  PointLights (models_res_nimble.py:191-198, taken when
// ifLight=False): the raw light LOCATION - the direction is per fragment, location - P (shade_fragment).
HFR_HD void light_dir_hat(const HfrShadeFwdArgs& a, int n, float* dhat, float* len) {
  const float d[3] = {a.light_dir[3 * n], a.light_dir[3 * n + 1], a.light_dir[3 * n + 2]};
  if (a.p.light_point) { dhat[0] = d[0]; dhat[1] = d[1]; dhat[2] = d[2]; *len = 1.0f; return; }
  hfr_normalize_eps(d, dhat, len);
}

// texel source of sample n: its own / the shared map, or the PCA texture model with the sample's coefficients
// (PCA is a template parameter so the plain-map kernels carry none of the model's state: with it as a runtime
// branch the fused forward lost 6 % and the backward 28 % to extra live registers)
// `sparams`: the sample's coefficients staged by the kernel (shared memory, zero padded to a multiple of 4), or NULL
template <bool PCA>
HFR_HD HfrTexSrc tex_source(const HfrShadeFwdArgs& a, int n, const float* sparams = nullptr) {
  const size_t map_floats = (size_t)a.p.tex_h * a.p.tex_w * 3;
  HfrTexSrc s = hfr_tex_plain(a.texture + (a.p.tex_n == 1 ? 0 : (size_t)n * map_floats));
  if (PCA) {
    s.map_floats = map_floats;
    s.npc = a.p.tex_pca;
    s.stride = a.p.tex_basis_stride;
    s.basis = a.tex_basis;
    s.params = sparams ? sparams : a.tex_params + (size_t)n * a.p.tex_pca;
    s.padded = sparams ? 1 : 0;
  }
  return s;
}

// colour of one fragment (Phong x UV texel).  `tap` and `ctx` are kept for the backward.
template <bool PCA = false>
HFR_HD void shade_fragment(const HfrShadeFwdArgs& a, int n, const FragGeom& g, const float* bc,
                                               const float* dhat, const float* lcol, float* color, HfrTexTap* tap,
                                               HfrPhongCtx* ctx, float* texel, const float* sparams = nullptr,
                                               float* duv = nullptr) {
  float P[3], Nn[3];
  interp3(bc, g.X, P);
  interp3(bc, g.Nv, Nn);
  const float u = bc[0] * g.uv[0] + bc[1] * g.uv[2] + bc[2] * g.uv[4];
  const float v = bc[0] * g.uv[1] + bc[1] * g.uv[3] + bc[2] * g.uv[5];
  hfr_tex_tap(a.p.tex_h, a.p.tex_w, u, v, tap);
  if (PCA && duv) hfr_tex_fetch_d(tex_source<PCA>(a, n, sparams), tap, texel, duv, duv + 3);   // backward: uv-derivative sums too
  else hfr_tex_fetch(tex_source<PCA>(a, n, sparams), tap, texel);
  if (a.p.light_point) {   // PointLights.diffuse / .specular: direction = location - points, normalised per fragment
    const float d[3] = {dhat[0] - P[0], dhat[1] - P[1], dhat[2] - P[2]};
    hfr_normalize_eps(d, ctx->lhat, &ctx->llen);
  } else {
    ctx->lhat[0] = dhat[0]; ctx->lhat[1] = dhat[1]; ctx->lhat[2] = dhat[2]; ctx->llen = 0.0f;
  }
  hfr_phong_fwd(a.p, P, Nn, ctx->lhat, lcol, texel, color, ctx);
}

// Full forward for one pixel given its K fragments.
template <int KMAX, bool PCA = false>
HFR_HD void shade_pixel(const HfrShadeFwdArgs& a, int n, const int64_t* id, const float* z,
                                            const float* d, const float* b, float* rgba, const float* sparams = nullptr) {
  const int K = a.p.K;
  {   // empty pixel: the blends reduce to the background (alpha 0); silhouette colour stays 1
    bool any = false;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) any = any || (k < K && id[k] >= 0);
    if (!any) {
      const bool ones = a.p.blend == HFR_BLEND_SIGMOID_ALPHA;
      rgba[0] = ones ? 1.0f : a.p.background[0]; rgba[1] = ones ? 1.0f : a.p.background[1];
      rgba[2] = ones ? 1.0f : a.p.background[2]; rgba[3] = 0.0f;
      return;
    }
  }
  bool valid[KMAX];
  float colors[KMAX * 3];
  float dhat[3], dlen;
  float lcol[3] = {0.f, 0.f, 0.f};
  const bool phong = a.p.shade == HFR_SHADE_PHONG_UV;
  if (phong) {
    light_dir_hat(a, n, dhat, &dlen);
    lcol[0] = a.light_color[3 * n]; lcol[1] = a.light_color[3 * n + 1]; lcol[2] = a.light_color[3 * n + 2];
  }
  // hard blending only looks at the nearest fragment
  const int kshade = a.p.blend == HFR_BLEND_SOFTMAX ? K : 1;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    valid[k] = k < K && id[k] >= 0;
    colors[3 * k] = colors[3 * k + 1] = colors[3 * k + 2] = 1.0f;
    if (phong && valid[k] && k < kshade) {
      FragGeom g;
      gather_frag(a, n, (int)(id[k] - (int64_t)n * a.p.F), g);
      HfrTexTap tap; HfrPhongCtx ctx; float texel[3];
      shade_fragment<PCA>(a, n, g, b + 3 * k, dhat, lcol, colors + 3 * k, &tap, &ctx, texel, sparams);
    }
  }
  hfr_blend_fwd<KMAX>(a.p, K, valid, z, d, colors, rgba);
}

}  // namespace hfr

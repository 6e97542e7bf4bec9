// Per-fragment shading and per-pixel blending math, host + device, forward and backward.
//
// Semantics (reference construction models_res_nimble.py:79-96,
// 187-190): TexturesUV.sample_textures (bilinear grid_sample, align_corners=True, border
// padding, map flipped vertically), interpolate_face_attributes, phong_shading with
// DirectionalLights (camera at the origin: R=I, T=0), and hard_rgb_blend /
// sigmoid_alpha_blend / softmax_rgb_blend.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------- texture
struct HfrTexTap {
  int idx[4];     // linear texel index (row*Wt + col) in the ORIGINAL (un-flipped) map, -1 = out of range
  float w[4];     // bilinear weights nw, ne, sw, se
  float ix, iy;   // clipped sample position
  float x0, y0;   // floor
  float mx, my;   // d(ix)/du, d(iy)/dv including the border-clip mask
};

HFR_HD void hfr_tex_tap(int Ht, int Wt, float u, float v, HfrTexTap* t) {
  const float gx = u * 2.0f - 1.0f, gy = v * 2.0f - 1.0f;
  float ix = ((gx + 1.0f) / 2.0f) * (float)(Wt - 1);
  float iy = ((gy + 1.0f) / 2.0f) * (float)(Ht - 1);
  t->mx = (ix <= 0.0f || ix >= (float)(Wt - 1)) ? 0.0f : (float)(Wt - 1);
  t->my = (iy <= 0.0f || iy >= (float)(Ht - 1)) ? 0.0f : (float)(Ht - 1);
  ix = fminf((float)(Wt - 1), fmaxf(ix, 0.0f));
  iy = fminf((float)(Ht - 1), fmaxf(iy, 0.0f));
  const float x0 = floorf(ix), y0 = floorf(iy);
  const int xi = (int)x0, yi = (int)y0;
  t->ix = ix; t->iy = iy; t->x0 = x0; t->y0 = y0;
  t->w[0] = (x0 + 1.0f - ix) * (y0 + 1.0f - iy);
  t->w[1] = (ix - x0) * (y0 + 1.0f - iy);
  t->w[2] = (x0 + 1.0f - ix) * (iy - y0);
  t->w[3] = (ix - x0) * (iy - y0);
  const bool xin0 = xi >= 0 && xi < Wt, xin1 = xi + 1 >= 0 && xi + 1 < Wt;
  const bool yin0 = yi >= 0 && yi < Ht, yin1 = yi + 1 >= 0 && yi + 1 < Ht;
  // the sampled map is flipped vertically: row y of the flipped map is row Ht-1-y of the original
  t->idx[0] = (xin0 && yin0) ? (Ht - 1 - yi) * Wt + xi : -1;
  t->idx[1] = (xin1 && yin0) ? (Ht - 1 - yi) * Wt + xi + 1 : -1;
  t->idx[2] = (xin0 && yin1) ? (Ht - 2 - yi) * Wt + xi : -1;
  t->idx[3] = (xin1 && yin1) ? (Ht - 2 - yi) * Wt + xi + 1 : -1;
}

// Where texel values come from: a plain map, or a PCA texture model evaluated on the fly,
//   texel(idx) = mean[idx] + sum_k params[k] * basis[k][idx]
// (NIMBLE-style per-sample texture, SURVEY.md 8(f) row 4: the per-sample 1024^2 map is never materialised).
struct HfrTexSrc {
  const float* tex;      // plain map of this sample, or the PCA mean map
  const float* basis;    // (npc, Ht, Wt, 3), or texel-major (Ht * Wt, stride) when stride > 0, or NULL
  const float* params;   // this sample's npc coefficients or NULL
  int npc;               // 0 = plain map
  int stride;            // > 0: floats per texel record of the texel-major basis (12 * ceil(npc / 4))
  int padded;            // != 0: params holds 4 * ceil(npc / 4) readable entries, zero beyond npc (the kernels' shared copy)
  size_t map_floats;     // Ht * Wt * 3
};

HFR_HD HfrTexSrc hfr_tex_plain(const float* tex) {
  HfrTexSrc s; s.tex = tex; s.basis = nullptr; s.params = nullptr; s.npc = 0; s.stride = 0; s.padded = 0; s.map_floats = 0;
  return s;
}

HFR_HD float4 hfr_ld4(const float4* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

HFR_HD void hfr_texel(const HfrTexSrc& src, int idx, float* v) {
  const float* m = src.tex + (size_t)idx * 3;
  v[0] = m[0]; v[1] = m[1]; v[2] = m[2];
  if (src.stride > 0) {
    // texel-major record: 4 components (12 floats, three 16-byte loads) per round, same summation order as below
    const float4* b = reinterpret_cast<const float4*>(src.basis + (size_t)idx * src.stride);
    for (int k0 = 0; k0 < src.npc; k0 += 4, b += 3) {
      const float4 r0 = hfr_ld4(b), r1 = hfr_ld4(b + 1), r2 = hfr_ld4(b + 2);
      const bool pad = src.padded != 0;
      const float p0 = src.params[k0], p1 = (pad || k0 + 1 < src.npc) ? src.params[k0 + 1] : 0.0f,
                  p2 = (pad || k0 + 2 < src.npc) ? src.params[k0 + 2] : 0.0f, p3 = (pad || k0 + 3 < src.npc) ? src.params[k0 + 3] : 0.0f;
      v[0] += p0 * r0.x; v[1] += p0 * r0.y; v[2] += p0 * r0.z;
      v[0] += p1 * r0.w; v[1] += p1 * r1.x; v[2] += p1 * r1.y;
      v[0] += p2 * r1.z; v[1] += p2 * r1.w; v[2] += p2 * r2.x;
      v[0] += p3 * r2.y; v[1] += p3 * r2.z; v[2] += p3 * r2.w;
    }
    return;
  }
  for (int k = 0; k < src.npc; ++k) {
    const float pk = src.params[k];
    const float* b = src.basis + (size_t)k * src.map_floats + (size_t)idx * 3;
    v[0] += pk * b[0]; v[1] += pk * b[1]; v[2] += pk * b[2];
  }
}

HFR_HD void hfr_tex_fetch(const HfrTexSrc& src, const HfrTexTap* t, float* out) {
  out[0] = out[1] = out[2] = 0.0f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (t->idx[q] >= 0) {
      float s[3];
      hfr_texel(src, t->idx[q], s);
      out[0] += s[0] * t->w[q]; out[1] += s[1] * t->w[q]; out[2] += s[2] * t->w[q];
    }
  }
}
HFR_HD void hfr_tex_fetch(const float* tex, const HfrTexTap* t, float* out) { hfr_tex_fetch(hfr_tex_plain(tex), t, out); }

// The fetch together with the two weight-derivative sums the uv gradient needs,
//   dax[c] = sum_q (d w_q / d ix) * texel_q[c],   day[c] = sum_q (d w_q / d iy) * texel_q[c],
// so the backward visits the taps ONCE (with a PCA texture a second visit is another 4 x npc basis reads):
// d(texel . g)/d(u, v) = ((dax . g) * mx, (day . g) * my)  -  hfr_tex_uv_grad_d below.
HFR_HD void hfr_tex_fetch_d(const HfrTexSrc& src, const HfrTexTap* t, float* out, float* dax, float* day) {
  out[0] = out[1] = out[2] = 0.0f;
  dax[0] = dax[1] = dax[2] = 0.0f;
  day[0] = day[1] = day[2] = 0.0f;
  const float ax = t->x0 + 1.0f - t->ix, bx = t->ix - t->x0, ay = t->y0 + 1.0f - t->iy, by = t->iy - t->y0;
  const float dwx[4] = {-ay, ay, -by, by};   // d w / d ix
  const float dwy[4] = {-ax, -bx, ax, bx};   // d w / d iy
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (t->idx[q] >= 0) {
      float s[3];
      hfr_texel(src, t->idx[q], s);
#pragma unroll
      for (int c = 0; c < 3; ++c) { out[c] += s[c] * t->w[q]; dax[c] += s[c] * dwx[q]; day[c] += s[c] * dwy[q]; }
    }
  }
}
HFR_HD void hfr_tex_uv_grad_d(const HfrTexTap* t, const float* dax, const float* day, const float* g, float* gu, float* gv) {
  *gu = (dax[0] * g[0] + dax[1] * g[1] + dax[2] * g[2]) * t->mx;
  *gv = (day[0] * g[0] + day[1] * g[1] + day[2] * g[2]) * t->my;
}

// d(texel)/d(u,v) contracted with g[3]
HFR_HD void hfr_tex_uv_grad(const HfrTexSrc& src, const HfrTexTap* t, const float* g, float* gu, float* gv) {
  float gix = 0.0f, giy = 0.0f;
  const float ax = t->x0 + 1.0f - t->ix, bx = t->ix - t->x0, ay = t->y0 + 1.0f - t->iy, by = t->iy - t->y0;
  const float dwx[4] = {-ay, ay, -by, by};   // d w / d ix
  const float dwy[4] = {-ax, -bx, ax, bx};   // d w / d iy
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (t->idx[q] >= 0) {
      float s[3];
      hfr_texel(src, t->idx[q], s);
      const float dot = s[0] * g[0] + s[1] * g[1] + s[2] * g[2];
      gix += dot * dwx[q]; giy += dot * dwy[q];
    }
  }
  *gu = gix * t->mx;
  *gv = giy * t->my;
}
HFR_HD void hfr_tex_uv_grad(const float* tex, const HfrTexTap* t, const float* g, float* gu, float* gv) {
  hfr_tex_uv_grad(hfr_tex_plain(tex), t, g, gu, gv);
}

// d(texel)/d(params[k]) contracted with g[3]: sum_q w_q * (g . basis[k][idx_q])
HFR_HD float hfr_tex_param_grad(const HfrTexSrc& src, const HfrTexTap* t, const float* g, int k) {
  float acc = 0.0f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (t->idx[q] >= 0) {
      const float* b = src.stride > 0 ? src.basis + (size_t)t->idx[q] * src.stride + 3 * k
                                      : src.basis + (size_t)k * src.map_floats + (size_t)t->idx[q] * 3;
      acc += t->w[q] * (b[0] * g[0] + b[1] * g[1] + b[2] * g[2]);
    }
  }
  return acc;
}

// the same for components k0 .. k0 + 3 at once (k0 % 4 == 0): with the texel-major basis every tap is three 16-byte
// loads, all twelve of a round independent
HFR_HD void hfr_tex_param_grad4(const HfrTexSrc& src, const HfrTexTap* t, const float* g, int k0, float* out) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  if (src.stride > 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (t->idx[q] >= 0) {
        const float4* b = reinterpret_cast<const float4*>(src.basis + (size_t)t->idx[q] * src.stride + 3 * k0);
        const float4 r0 = hfr_ld4(b), r1 = hfr_ld4(b + 1), r2 = hfr_ld4(b + 2);
        const float w = t->w[q];
        out[0] += w * (r0.x * g[0] + r0.y * g[1] + r0.z * g[2]);
        out[1] += w * (r0.w * g[0] + r1.x * g[1] + r1.y * g[2]);
        out[2] += w * (r1.z * g[0] + r1.w * g[1] + r2.x * g[2]);
        out[3] += w * (r2.y * g[0] + r2.z * g[1] + r2.w * g[2]);
      }
    }
    return;
  }
  for (int i = 0; i < 4; ++i)
    if (k0 + i < src.npc) out[i] = hfr_tex_param_grad(src, t, g, k0 + i);
}

// ---------------------------------------------------------------------------------- lighting
HFR_HD void hfr_normalize_eps(const float* x, float* o, float* len_out) {
  const float len = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const float inv = HFR_RCP(fmaxf(len, 1e-6f));
  o[0] = x[0] * inv; o[1] = x[1] * inv; o[2] = x[2] * inv;
  *len_out = len;
}
// backward of o = x / max(|x|, eps)
HFR_HD void hfr_normalize_eps_bwd(const float* o, float len, const float* g, float* gx) {
  if (len >= 1e-6f) {
    const float d = o[0] * g[0] + o[1] * g[1] + o[2] * g[2];
    const float inv = HFR_RCP(len);
    gx[0] = (g[0] - o[0] * d) * inv; gx[1] = (g[1] - o[1] * d) * inv; gx[2] = (g[2] - o[2] * d) * inv;
  } else {
    gx[0] = g[0] * 1e6f; gx[1] = g[1] * 1e6f; gx[2] = g[2] * 1e6f;
  }
}

struct HfrPhongCtx {   // forward intermediates kept for the backward
  float nh[3], nlen, view[3], vlen, refl[3], cosang, vr, alpha, spec;
  float lhat[3], llen;   // unit light direction of THIS fragment (PointLights: normalize(location - P)) and |location - P|
};

// P: interpolated view-space position, Nn: interpolated (un-normalised) normal, dhat: unit light
// direction, lcol: light diffuse colour, texel: sampled texture.
HFR_HD void hfr_phong_fwd(const HfrShadeParams& p, const float* P, const float* Nn, const float* dhat,
                          const float* lcol, const float* texel, float* color, HfrPhongCtx* c) {
  hfr_normalize_eps(Nn, c->nh, &c->nlen);
  c->cosang = c->nh[0] * dhat[0] + c->nh[1] * dhat[1] + c->nh[2] * dhat[2];
  const float relu_cos = fmaxf(c->cosang, 0.0f);
  const float mask = c->cosang > 0.0f ? 1.0f : 0.0f;
  const float negP[3] = {-P[0], -P[1], -P[2]};   // camera centre is the origin
  hfr_normalize_eps(negP, c->view, &c->vlen);
#pragma unroll
  for (int k = 0; k < 3; ++k) c->refl[k] = -dhat[k] + 2.0f * (c->cosang * c->nh[k]);
  c->vr = c->view[0] * c->refl[0] + c->view[1] * c->refl[1] + c->view[2] * c->refl[2];
  c->alpha = fmaxf(c->vr, 0.0f) * mask;
  c->spec = HFR_POW(c->alpha, p.shininess);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float amb = p.mat_ambient[k] * p.light_ambient[k];
    const float dif = p.mat_diffuse[k] * (lcol[k] * relu_cos);
    const float spe = p.mat_specular[k] * (p.light_specular[k] * c->spec);
    color[k] = (amb + dif) * texel[k] + spe;
  }
}

// Accumulates into g_dhat / g_lcol; writes gP, gNn, gtexel.
HFR_HD void hfr_phong_bwd(const HfrShadeParams& p, const float* dhat, const float* lcol, const float* texel,
                          const HfrPhongCtx* c, const float* gcol, float* gP, float* gNn, float* gtexel,
                          float* g_dhat, float* g_lcol) {
  const float relu_cos = fmaxf(c->cosang, 0.0f);
  const float mask = c->cosang > 0.0f ? 1.0f : 0.0f;
  float g_relu = 0.0f, g_specs = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float amb = p.mat_ambient[k] * p.light_ambient[k];
    const float dif = p.mat_diffuse[k] * (lcol[k] * relu_cos);
    gtexel[k] = gcol[k] * (amb + dif);
    const float g_dif = gcol[k] * texel[k] * p.mat_diffuse[k];
    g_lcol[k] += g_dif * relu_cos;
    g_relu += g_dif * lcol[k];
    g_specs += gcol[k] * p.mat_specular[k] * p.light_specular[k];
  }
  float g_cos = g_relu * mask;
  // spec = alpha^s ; alpha = relu(vr) * mask
  float g_alpha = 0.0f;
  if (c->alpha > 0.0f) g_alpha = g_specs * p.shininess * HFR_POW(c->alpha, p.shininess - 1.0f);
  const float g_vr = (c->vr > 0.0f) ? g_alpha * mask : 0.0f;
  float g_view[3], g_refl[3], g_nh[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { g_view[k] = g_vr * c->refl[k]; g_refl[k] = g_vr * c->view[k]; }
  // refl = -dhat + 2 cos nh
  const float rn = g_refl[0] * c->nh[0] + g_refl[1] * c->nh[1] + g_refl[2] * c->nh[2];
  g_cos += 2.0f * rn;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g_nh[k] = 2.0f * c->cosang * g_refl[k] + g_cos * dhat[k];
    g_dhat[k] += -g_refl[k] + g_cos * c->nh[k];
  }
  hfr_normalize_eps_bwd(c->nh, c->nlen, g_nh, gNn);
  float g_negP[3];
  hfr_normalize_eps_bwd(c->view, c->vlen, g_view, g_negP);
  gP[0] = -g_negP[0]; gP[1] = -g_negP[1]; gP[2] = -g_negP[2];
}

// ---------------------------------------------------------------------------------- blending
HFR_HD float hfr_sigmoid(float x) { return HFR_FDIV(1.0f, 1.0f + HFR_EXP(-x)); }

// colors: K*3 (ignored for entries with valid[k]==0).  Returns rgba[4].
template <int KMAX>
HFR_HD void hfr_blend_fwd(const HfrShadeParams& p, int K, const bool* valid, const float* z, const float* d,
                          const float* colors, float* rgba) {
  if (p.blend == HFR_BLEND_HARD) {
    if (valid[0]) { rgba[0] = colors[0]; rgba[1] = colors[1]; rgba[2] = colors[2]; rgba[3] = 1.0f; }
    else { rgba[0] = p.background[0]; rgba[1] = p.background[1]; rgba[2] = p.background[2]; rgba[3] = 0.0f; }
    return;
  }
  float prod = 1.0f;
  float prob[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    prob[k] = 0.0f;
    if (k < K && valid[k]) prob[k] = hfr_sigmoid(HFR_FDIV(-d[k], p.sigma));
    if (k < K) prod *= (1.0f - prob[k]);
  }
  rgba[3] = 1.0f - prod;
  if (p.blend == HFR_BLEND_SIGMOID_ALPHA) {
    rgba[0] = colors[0]; rgba[1] = colors[1]; rgba[2] = colors[2];
    return;
  }
  const float eps = 1e-10f, zr = p.zfar - p.znear;
  float zmax = 0.0f;   // masked z_inv of empty slots is 0
  float zinv[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    zinv[k] = 0.0f;
    if (k < K && valid[k]) zinv[k] = (p.zfar - z[k]) / zr;   // IEEE divide: 1 ulp of z_inv is amplified by 1/gamma
    if (k < K) zmax = (k == 0) ? zinv[k] : fmaxf(zmax, zinv[k]);
  }
  zmax = fmaxf(zmax, eps);
  float wsum = 0.0f, acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      const float w = prob[k] * HFR_EXP(HFR_FDIV(zinv[k] - zmax, p.gamma));
      wsum += w;
      if (valid[k]) { acc[0] += w * colors[3 * k]; acc[1] += w * colors[3 * k + 1]; acc[2] += w * colors[3 * k + 2]; }
    }
  }
  const float delta = fmaxf(HFR_EXP(HFR_FDIV(eps - zmax, p.gamma)), eps);
  const float den = wsum + delta;
#pragma unroll
  for (int c = 0; c < 3; ++c) rgba[c] = HFR_FDIV(acc[c] + delta * p.background[c], den);
}

// g_rgba[4] -> g_colors (K*3), g_z (K), g_d (K).
template <int KMAX>
HFR_HD void hfr_blend_bwd(const HfrShadeParams& p, int K, const bool* valid, const float* z, const float* d,
                          const float* colors, const float* g_rgba, float* g_colors, float* g_z, float* g_d) {
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { g_z[k] = 0.0f; g_d[k] = 0.0f; g_colors[3 * k] = g_colors[3 * k + 1] = g_colors[3 * k + 2] = 0.0f; }
  if (p.blend == HFR_BLEND_HARD) {
    if (valid[0]) { g_colors[0] = g_rgba[0]; g_colors[1] = g_rgba[1]; g_colors[2] = g_rgba[2]; }
    return;
  }
  float prob[KMAX], gprob[KMAX];
  float prod = 1.0f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    prob[k] = 0.0f; gprob[k] = 0.0f;
    if (k < K && valid[k]) prob[k] = hfr_sigmoid(HFR_FDIV(-d[k], p.sigma));
  }
  // alpha channel = 1 - prod(1 - p_k): d/dp_k = prod_{j != k}(1 - p_j)
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      float others = 1.0f;
#pragma unroll
      for (int j = 0; j < KMAX; ++j)
        if (j < K && j != k) others *= (1.0f - prob[j]);
      gprob[k] = g_rgba[3] * others;
    }
  }
  (void)prod;
  if (p.blend == HFR_BLEND_SIGMOID_ALPHA) {
    if (true) { g_colors[0] = g_rgba[0]; g_colors[1] = g_rgba[1]; g_colors[2] = g_rgba[2]; }
  } else {
    const float eps = 1e-10f, zr = p.zfar - p.znear;
    float zinv[KMAX], e[KMAX];
    float zmax_raw = 0.0f;
    int kmax = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      zinv[k] = 0.0f;
      if (k < K && valid[k]) zinv[k] = (p.zfar - z[k]) / zr;   // IEEE divide: 1 ulp of z_inv is amplified by 1/gamma
      if (k < K && (k == 0 || zinv[k] > zmax_raw)) { zmax_raw = zinv[k]; kmax = k; }
    }
    const float zmax = fmaxf(zmax_raw, eps);
    float wsum = 0.0f, acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      e[k] = 0.0f;
      if (k < K) {
        e[k] = HFR_EXP(HFR_FDIV(zinv[k] - zmax, p.gamma));
        const float w = prob[k] * e[k];
        wsum += w;
        if (valid[k]) { acc[0] += w * colors[3 * k]; acc[1] += w * colors[3 * k + 1]; acc[2] += w * colors[3 * k + 2]; }
      }
    }
    const float dexp = HFR_EXP(HFR_FDIV(eps - zmax, p.gamma));
    const float delta = fmaxf(dexp, eps);
    const float den = wsum + delta;
    float rgb[3], gnum[3];
    float gden = 0.0f, gdelta = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      rgb[c] = HFR_FDIV(acc[c] + delta * p.background[c], den);
      gnum[c] = HFR_FDIV(g_rgba[c], den);
      gden -= gnum[c] * rgb[c];
      gdelta += gnum[c] * p.background[c];
    }
    gdelta += gden;
    float gzmax = 0.0f;
    float gzinv[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      gzinv[k] = 0.0f;
      if (k < K) {
        float gw = gden;
        if (valid[k]) {
          gw += gnum[0] * colors[3 * k] + gnum[1] * colors[3 * k + 1] + gnum[2] * colors[3 * k + 2];
          const float w = prob[k] * e[k];
          g_colors[3 * k] = w * gnum[0]; g_colors[3 * k + 1] = w * gnum[1]; g_colors[3 * k + 2] = w * gnum[2];
        }
        gprob[k] += gw * e[k];
        const float ge = HFR_FDIV(gw * prob[k] * e[k], p.gamma);   // d/d((zinv - zmax))
        gzinv[k] = ge;
        gzmax -= ge;
      }
    }
    if (dexp >= eps) gzmax -= HFR_FDIV(gdelta * delta, p.gamma);
    if (zmax_raw >= eps) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k == kmax) gzinv[k] += gzmax;
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K && valid[k]) g_z[k] = HFR_FDIV(-gzinv[k], zr);
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K && valid[k]) {
      const float s = prob[k];
      g_d[k] = HFR_FDIV(-gprob[k] * s * (1.0f - s), p.sigma);
    }
  }
}

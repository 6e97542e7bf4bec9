// TEST-ONLY host build of the kernels' per-element math (common.cuh / *_math.cuh compile for
// the host with g++ -ffp-contract=off).  It lets the no-GPU CI check the device formulas
// against the oracle; the product never loads this library and has no CPU path.
#include <stdint.h>
#include <math.h>
#include "mano_math.cuh"
#include "raster_math.cuh"
#ifdef HFR_HAVE_SHADE
#include "shade_math.cuh"
#endif
void hfr_set_error(const char*, ...) {}

extern "C" {
void emul_rodrigues_fwd(const float* v, float* R, int n) { for (int i = 0; i < n; ++i) hfr_rodrigues_fwd(v + 3 * i, R + 9 * i); }
void emul_rodrigues_bwd(const float* v, const float* g, float* gv, int n) { for (int i = 0; i < n; ++i) hfr_rodrigues_bwd(v + 3 * i, g + 9 * i, gv + 3 * i); }

// the exact edge-sign rejection of the K = 1 / blur 0 walk against the full evaluation: for every (pixel, face) pair
// returns bit 0 = filter says "certainly outside", bit 1 = the exact math says inside (valid face, pz >= 0)
void emul_sign_filter(const float* fv, const float* pxy, int n, int pc, unsigned char* out) {
  for (int i = 0; i < n; ++i) {
    const float* v = fv + 9 * i;
    const float px = pxy[2 * i], py = pxy[2 * i + 1];
    const float area = XADD(hfr_edge(v[6], v[7], v[0], v[1], v[3], v[4]), HFR_KEPS);
    const bool reject = hfr_edge_sign_outside(px, py, v[0], v[1], v[3], v[4], v[6], v[7], area);
    float pz, bc[3];
    bool inside = false;
    const bool ok = hfr_raster_bary(px, py, v, area, pc, 0, &pz, bc, &inside);
    out[i] = (unsigned char)((reject ? 1 : 0) | ((ok && inside) ? 2 : 0));
  }
}

// depth of every (pixel, face) pair as the fine pass computes it (clip = 1: clamped barycentrics; clip = 0: as given);
// out_pz = pz, flags bit 0 = accepted by hfr_raster_bary (pz >= 0), bit 1 = inside, bit 2 = the face is valid (only valid
// faces reach the fine pass: the setup kernel gives the others an empty tile range)
void emul_pair_depth(const float* fv, const float* pxy, int n, int pc, int clip, float* out_pz, unsigned char* flags) {
  for (int i = 0; i < n; ++i) {
    const float* v = fv + 9 * i;
    const float area = XADD(hfr_edge(v[6], v[7], v[0], v[1], v[3], v[4]), HFR_KEPS);
    float pz = 0.f, bc[3];
    bool inside = false;
    const bool ok = hfr_raster_bary(pxy[2 * i], pxy[2 * i + 1], v, area, pc, clip, &pz, bc, &inside);
    out_pz[i] = pz;
    flags[i] = (unsigned char)((ok ? 1 : 0) | (inside ? 2 : 0) | (hfr_face_valid(v, 0) ? 4 : 0));
  }
}

// naive loop over one mesh using the device evaluation function
void emul_raster(const float* fv, int64_t F, int H, int W, int K, float blur, int pc, int clip, int cull,
                 int64_t* p2f, float* zb, float* ba, float* ds) {
  const float rb = sqrtf(blur);
  for (int yi = 0; yi < H; ++yi) for (int xi = 0; xi < W; ++xi) {
    const float xf = hfr_pix_to_ndc(W - 1 - xi, W, H), yf = hfr_pix_to_ndc(H - 1 - yi, H, W);
    float z[64], d[64], b[192]; int64_t id[64]; int cnt = 0;
    for (int64_t f = 0; f < F; ++f) {
      float pz, bc[3], sd;
      if (!hfr_raster_eval(xf, yf, fv + f * 9, blur, rb, pc, clip, cull, &pz, bc, &sd)) continue;
      if (cnt == K && !(pz < z[K - 1])) continue;
      int pos = cnt < K ? cnt : K - 1;
      while (pos > 0 && pz < z[pos - 1]) { z[pos] = z[pos-1]; d[pos] = d[pos-1]; id[pos] = id[pos-1]; for (int e=0;e<3;++e) b[3*pos+e]=b[3*(pos-1)+e]; --pos; }
      z[pos] = pz; d[pos] = sd; id[pos] = f; b[3*pos]=bc[0]; b[3*pos+1]=bc[1]; b[3*pos+2]=bc[2];
      if (cnt < K) ++cnt;
    }
    const long base = ((long)yi * W + xi) * K;
    for (int k = 0; k < K; ++k) {
      const bool v = k < cnt;
      p2f[base+k] = v ? id[k] : -1; zb[base+k] = v ? z[k] : -1.f; ds[base+k] = v ? d[k] : -1.f;
      for (int e=0;e<3;++e) ba[(base+k)*3+e] = v ? b[3*k+e] : -1.f;
    }
  }
}

void emul_raster_bwd(const float* fv, const int64_t* p2f, const float* gz, const float* gb, const float* gd,
                     int H, int W, int K, int pc, int clip, float* gfv) {
  for (int yi = 0; yi < H; ++yi) for (int xi = 0; xi < W; ++xi) {
    const float xf = hfr_pix_to_ndc(W - 1 - xi, W, H), yf = hfr_pix_to_ndc(H - 1 - yi, H, W);
    for (int k = 0; k < K; ++k) {
      const long i = ((long)yi * W + xi) * K + k;
      if (p2f[i] < 0) continue;
      hfr_raster_eval_bwd(xf, yf, fv + p2f[i] * 9, pc, clip, gb + 3 * i, gz[i], gd[i], gfv + p2f[i] * 9);
    }
  }
}
}

#ifdef HFR_HAVE_SHADE
#include "shade_pixel.cuh"
extern "C" {
// forward shading of a whole (N,H,W,K) fragment set, K <= 8
void emul_shade_fwd(const HfrShadeFwdArgs* a) {
  const HfrShadeParams& p = a->p;
  for (int n = 0; n < p.N; ++n) for (int q = 0; q < p.H * p.W; ++q) {
    const size_t pix = (size_t)n * p.H * p.W + q;
    int64_t id[8]; float z[8], d[8], b[24];
    for (int k = 0; k < 8; ++k) {
      id[k] = -1; z[k] = d[k] = -1.f; b[3*k] = b[3*k+1] = b[3*k+2] = -1.f;
      if (k < p.K) { id[k] = a->pix_to_face[pix*p.K+k]; z[k] = a->zbuf[pix*p.K+k]; d[k] = a->dists[pix*p.K+k];
        for (int e = 0; e < 3; ++e) b[3*k+e] = a->bary[(pix*p.K+k)*3+e]; }
    }
    hfr::shade_pixel<8>(*a, n, id, z, d, b, a->image + pix * 4);
  }
}
// blend backward for P pixels with K<=8 fragments: colors (P,K,3)
void emul_blend_bwd(const HfrShadeParams* p, int P, const int64_t* id, const float* z, const float* d,
                    const float* colors, const float* g_rgba, float* g_colors, float* g_z, float* g_d) {
  const int K = p->K;
  for (int i = 0; i < P; ++i) {
    bool valid[8]; float zz[8], dd[8], cc[24], gc[24], gz[8], gd[8];
    for (int k = 0; k < 8; ++k) { valid[k] = k < K && id[i*K+k] >= 0; zz[k] = k<K? z[i*K+k]:-1.f; dd[k] = k<K? d[i*K+k]:-1.f;
      for (int e=0;e<3;++e) cc[3*k+e] = k<K ? colors[(i*K+k)*3+e] : 1.f; }
    hfr_blend_bwd<8>(*p, K, valid, zz, dd, cc, g_rgba + 4*i, gc, gz, gd);
    for (int k = 0; k < K; ++k) { g_z[i*K+k] = gz[k]; g_d[i*K+k] = gd[k]; for (int e=0;e<3;++e) g_colors[(i*K+k)*3+e] = gc[3*k+e]; }
  }
}
// phong fwd+bwd for M fragments
void emul_phong(const HfrShadeParams* p, int M, const float* P, const float* Nn, const float* dhat, const float* lcol,
                const float* texel, const float* gcol, float* color, float* gP, float* gNn, float* gtexel,
                float* g_dhat, float* g_lcol) {
  for (int i = 0; i < M; ++i) {
    HfrPhongCtx c;
    hfr_phong_fwd(*p, P+3*i, Nn+3*i, dhat, lcol, texel+3*i, color+3*i, &c);
    hfr_phong_bwd(*p, dhat, lcol, texel+3*i, &c, gcol+3*i, gP+3*i, gNn+3*i, gtexel+3*i, g_dhat, g_lcol);
  }
}
// texture sample fwd + uv grad + scatter for M samples
void emul_tex(const float* tex, int Ht, int Wt, int M, const float* uv, const float* g, float* out, float* guv, float* gtex) {
  for (int i = 0; i < M; ++i) {
    HfrTexTap t; hfr_tex_tap(Ht, Wt, uv[2*i], uv[2*i+1], &t);
    hfr_tex_fetch(tex, &t, out + 3*i);
    hfr_tex_uv_grad(tex, &t, g + 3*i, guv + 2*i, guv + 2*i + 1);
    for (int q = 0; q < 4; ++q) if (t.idx[q] >= 0) for (int c = 0; c < 3; ++c) gtex[(size_t)t.idx[q]*3+c] += t.w[q]*g[3*i+c];
  }
}
// PCA texture model (mean + sum_k params[k] * basis[k]) sampled at the taps: fetch, uv grad, d/d(params)
void emul_tex_pca(const float* mean, const float* basis, const float* params, int npc, int Ht, int Wt, int M, const float* uv,
                  const float* g, float* out, float* guv, float* gparams, int stride) {
  HfrTexSrc src;
  src.tex = mean; src.basis = basis; src.params = params; src.npc = npc; src.map_floats = (size_t)Ht * Wt * 3;
  src.stride = stride < 0 ? -stride : stride;   // 0: (npc,Ht,Wt,3); > 0: texel-major records of `stride` floats; < 0: the same
  src.padded = 0;                               // through hfr_tex_fetch_d
  for (int i = 0; i < M; ++i) {
    HfrTexTap t; hfr_tex_tap(Ht, Wt, uv[2*i], uv[2*i+1], &t);
    if (stride < 0) {   // the backward's single visit of the taps: texel + uv-derivative sums
      float dax[3], day[3];
      src.stride = -stride;
      hfr_tex_fetch_d(src, &t, out + 3*i, dax, day);
      hfr_tex_uv_grad_d(&t, dax, day, g + 3*i, guv + 2*i, guv + 2*i + 1);
    } else {
      hfr_tex_fetch(src, &t, out + 3*i);
      hfr_tex_uv_grad(src, &t, g + 3*i, guv + 2*i, guv + 2*i + 1);
    }
    for (int k0 = 0; k0 < npc; k0 += 4) {
      float t4[4];
      hfr_tex_param_grad4(src, &t, g + 3*i, k0, t4);
      for (int j = 0; j < 4 && k0 + j < npc; ++j) gparams[k0 + j] += t4[j];
    }
  }
}
}
#endif

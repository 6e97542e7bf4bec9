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
#ifndef HFR_RASTER_MINB
#define HFR_RASTER_MINB 4
#endif
template <int KMAX>
__global__ void __launch_bounds__(kRasterThreads, (KMAX <= 4 ? HFR_RASTER_MINB : 1)) raster_shade_fwd_kernel(HfrRasterArgs r, HfrShadeFwdArgs s,
                                                                          const uint32_t* __restrict__ ranges,
                                                                          const uint32_t* __restrict__ mesh_box) {
  __shared__ RasterSmem sm;
  const PixelCtx c = make_pixel_ctx(r.H, r.W);
  TopK<KMAX> top;
  raster_tile<KMAX>(r, ranges, mesh_box, sm, c.n, c.tx, c.ty, c.xf, c.yf, c.pix_active, c.warp_active, c.wx_lo, c.wx_hi,
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

int check_shade(const HfrShadeFwdArgs* a, const char* who) {
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
#define CALL(KM) raster_shade_fwd_kernel<KM><<<grid, kRasterThreads, 0, st>>>(a->r, s, ranges, raster_mesh_box(a->r))
  HFR_DISPATCH_K(a->r.K, CALL);
#undef CALL
  HFR_CHECK_LAUNCH("raster_shade_forward");
  return HFR_OK;
}

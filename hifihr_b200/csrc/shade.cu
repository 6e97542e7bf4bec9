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

template <int KMAX, bool PCA>
__global__ void __launch_bounds__(256) shade_fwd_kernel(HfrShadeFwdArgs a) {
  const int n = blockIdx.y, K = a.p.K;
  const int HW = a.p.H * a.p.W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  // texture PCA: this sample's coefficients, staged once per CTA and zero padded to a multiple of 4
  __shared__ float s_tp[PCA ? HFR_MAX_TEX_PCA : 1];
  if (PCA) {
    if (threadIdx.x < HFR_MAX_TEX_PCA) s_tp[threadIdx.x] = (int)threadIdx.x < a.p.tex_pca ? a.tex_params[(size_t)n * a.p.tex_pca + threadIdx.x] : 0.0f;
    __syncthreads();
  }
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
  shade_pixel<KMAX, PCA>(a, n, id, z, d, b, rgba, PCA ? s_tp : nullptr);
  *reinterpret_cast<float4*>(a.image + pix * 4) = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
}

// Epilogue of the fused kernels: the K winners of one pixel -> Fragments (written here) and the blended RGBA
// (returned).  ONE runtime loop over the winners (a single copy of the exact fragment math, the gathers, the
// texture fetch and the Phong code - the kernel stays inside the instruction cache).  The blends need no second
// pass: winners are sorted by depth, so the softmax reference depth z_max belongs to slot 0 and weights
// accumulate on the fly.
// PAY: the winners' barycentrics / signed distances come from the payload cache the tile core filled
// (raster_tile.cuh) instead of being recomputed from the face's vertices.
template <int KMAX, bool PAY>
__device__ __forceinline__ void raster_shade_epilogue(const HfrRasterArgs& r, const HfrShadeFwdArgs& s, const PixelCtx& c,
                                                      const TopK<KMAX>& top, size_t pix, const float4* __restrict__ pay,
                                                      uint32_t perm, float (&rgba)[4]) {
  const HfrShadeParams& P = s.p;
  const int K = r.K;
  const bool ones = P.blend == HFR_BLEND_SIGMOID_ALPHA;
  int64_t id[KMAX];
  float z[KMAX], d[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { id[k] = -1; z[k] = -1.0f; d[k] = -1.0f; }
  float* __restrict__ ba = r.bary + pix * K * 3;
  rgba[0] = ones ? 1.0f : P.background[0]; rgba[1] = ones ? 1.0f : P.background[1]; rgba[2] = ones ? 1.0f : P.background[2]; rgba[3] = 0.0f;
  if (top.f[0] < 0) {
    // empty pixel (winners are sorted, slot 0 empty = all empty): -1 fill, background colour, alpha 0
    if (K == KMAX && (KMAX % 4) == 0) {
#pragma unroll
      for (int e = 0; e < KMAX * 3; e += 4) st_cs_f4(ba + e, -1.0f, -1.0f, -1.0f, -1.0f);
    } else {
      for (int e = 0; e < K * 3; ++e) ba[e] = -1.0f;
    }
  } else {
    const bool phong = P.shade == HFR_SHADE_PHONG_UV;
    const int kshade = phong ? (P.blend == HFR_BLEND_SOFTMAX ? K : 1) : 0;
    float dhat[3] = {0.f, 0.f, 0.f}, dlen, lcol[3] = {0.f, 0.f, 0.f};
    if (phong) {
      light_dir_hat(s, c.n, dhat, &dlen);
      lcol[0] = __ldg(s.light_color + 3 * c.n); lcol[1] = __ldg(s.light_color + 3 * c.n + 1); lcol[2] = __ldg(s.light_color + 3 * c.n + 2);
    }
    const float eps = 1e-10f, zr = P.zfar - P.znear;
    const float zmax = fmaxf((P.zfar - top.z[0]) / zr, eps);   // slot 0 is the nearest winner
    float prod = 1.0f, wsum = 0.0f, acc[3] = {0.f, 0.f, 0.f}, col0[3] = {1.0f, 1.0f, 1.0f};
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      int face = top.f[0];
#pragma unroll
      for (int i = 1; i < KMAX; ++i) face = (k == i) ? top.f[i] : face;
      if (face < 0) {
        ba[3 * k] = -1.0f; ba[3 * k + 1] = -1.0f; ba[3 * k + 2] = -1.0f;
        continue;
      }
      float pz, bc[3], sd;
      if (PAY) {
        const float4 q = pay[((perm >> (4 * k)) & 0xfu) * kRasterThreads];
        bc[0] = q.x; bc[1] = q.y; bc[2] = q.z; sd = q.w;
        pz = top.z[0];
#pragma unroll
        for (int i = 1; i < KMAX; ++i) pz = (k == i) ? top.z[i] : pz;
      } else {
        float v[9];
        const float* __restrict__ src = r.face_verts + (size_t)face * 9;
#pragma unroll
        for (int e = 0; e < 9; ++e) v[e] = __ldg(src + e);
        const float area = XADD(hfr_edge(v[6], v[7], v[0], v[1], v[3], v[4]), HFR_KEPS);
        bool inside;
        hfr_raster_bary(c.xf, c.yf, v, area, r.perspective_correct, r.clip_barycentric, &pz, bc, &inside);
        const float dd = hfr_tri_dist2(c.xf, c.yf, v);
        sd = inside ? -dd : dd;
      }
      ba[3 * k] = bc[0]; ba[3 * k + 1] = bc[1]; ba[3 * k + 2] = bc[2];
#pragma unroll
      for (int i = 0; i < KMAX; ++i)
        if (i == k) { id[i] = face; z[i] = pz; d[i] = sd; }
      float col[3] = {1.0f, 1.0f, 1.0f};
      if (k < kshade) {
        FragGeom g;
        gather_frag(s, c.n, (int)(face - (int64_t)c.n * P.F), g);
        HfrTexTap tap; HfrPhongCtx ctx; float texel[3];
        shade_fragment(s, c.n, g, bc, dhat, lcol, col, &tap, &ctx, texel);
      }
      if (k == 0) { col0[0] = col[0]; col0[1] = col[1]; col0[2] = col[2]; }
      if (P.blend != HFR_BLEND_HARD) {
        const float prob = hfr_sigmoid(HFR_FDIV(-sd, P.sigma));
        prod *= (1.0f - prob);
        if (P.blend == HFR_BLEND_SOFTMAX) {
          const float zinv = (P.zfar - pz) / zr;   // IEEE divide: 1 ulp of z_inv is amplified by 1/gamma
          const float w = prob * HFR_EXP(HFR_FDIV(zinv - zmax, P.gamma));
          wsum += w;
          acc[0] += w * col[0]; acc[1] += w * col[1]; acc[2] += w * col[2];
        }
      }
    }
    if (P.blend == HFR_BLEND_HARD) {
      rgba[0] = col0[0]; rgba[1] = col0[1]; rgba[2] = col0[2]; rgba[3] = 1.0f;
    } else if (P.blend == HFR_BLEND_SIGMOID_ALPHA) {
      rgba[0] = col0[0]; rgba[1] = col0[1]; rgba[2] = col0[2]; rgba[3] = 1.0f - prod;
    } else {
      const float delta = fmaxf(HFR_EXP(HFR_FDIV(eps - zmax, P.gamma)), eps);
      const float den = wsum + delta;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) rgba[ch] = HFR_FDIV(acc[ch] + delta * P.background[ch], den);
      rgba[3] = 1.0f - prod;
    }
  }
  // Fragments ids / depths / distances: 128-bit evict-first stores when K allows
  {
    int64_t* p2f = r.pix_to_face + pix * K;
    float* zb = r.zbuf + pix * K;
    float* ds = r.dists + pix * K;
    if (K == KMAX && (KMAX % 4) == 0) {
#pragma unroll
      for (int k = 0; k < KMAX; k += 2) st_cs_i64x2(p2f + k, id[k], id[k + 1]);
#pragma unroll
      for (int k = 0; k < KMAX; k += 4) {
        st_cs_f4(zb + k, z[k], z[k + 1], z[k + 2], z[k + 3]);
        st_cs_f4(ds + k, d[k], d[k + 1], d[k + 2], d[k + 3]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (k < K) { p2f[k] = id[k]; zb[k] = z[k]; ds[k] = d[k]; }
    }
  }
}

// Fused rasterize + shade forward: Fragments and the RGBA image leave the SM in the same pass.
#ifndef HFR_RASTER_MINB
#define HFR_RASTER_MINB 4
#endif
#ifndef HFR_RASTER_MINB8
#define HFR_RASTER_MINB8 4   // K = 8: measured on B200 at C5 (512^2, B=32): 1 CTA/SM (148 registers) 1943 us, 2: 1142, 3: 957, 4: 937
#endif
#define HFR_RASTER_MINB_FOR(KMAX) ((KMAX) <= 4 ? HFR_RASTER_MINB : ((KMAX) <= 8 ? HFR_RASTER_MINB8 : 2))
// Tiles outside the mesh can stream their -1 Fragments as whole rows (fill_empty_tile).  Measured on B200: with
// the generic unit loop (any K) neutral to slightly slower in the one-tile-per-CTA kernel (473 -> 480 us at 672^2
// K=1) - off; with the fixed-role K=1 / K=4 paths (HFR_FAST_FILL_K1) a win there (K=1 672^2: 472 -> 426 us, C2 K=4:
// 332 -> 320 us) and inside the SSAA-fused kernel, where it also skips the tile's shared-memory round trip and two
// barriers (663 -> 524 us) - on.
#ifndef HFR_FAST_FILL
#define HFR_FAST_FILL 0
#endif
#ifndef HFR_FAST_FILL_K1
#define HFR_FAST_FILL_K1 1   // K = 1 and K = 4, 8, 16: fixed-role row fills
#endif
#ifndef HFR_FAST_FILL_POOL
#define HFR_FAST_FILL_POOL 1
#endif
#ifndef HFR_PAY_MAXK
#define HFR_PAY_MAXK 4   // payload cache for K <= this (0 disables it: the epilogue recomputes the winners); the slot
                          // permutation is one nibble per sorted position of a 32-bit word, so 8 is the limit.  K = 8 with
                          // the cache (32 KB more shared memory, 3 instead of 4 CTAs / SM) measured SLOWER: C5 B=32 forward
                          // 937 -> 1073 us
#endif
#ifndef HFR_FILL_TMA
#define HFR_FILL_TMA 1   // tile-queue path: empty tiles are filled by bulk shared->global copies (0: vector stores)
#endif
template <int KMAX>
__global__ void __launch_bounds__(kRasterThreads, HFR_RASTER_MINB_FOR(KMAX)) raster_shade_fwd_kernel(HfrRasterArgs r, HfrShadeFwdArgs s,
                                                                          const uint32_t* __restrict__ ranges,
                                                                          const uint32_t* __restrict__ mesh_box,
                                                                          const uint32_t* __restrict__ queue) {
  __shared__ RasterSmem sm;
  constexpr bool PAY = KMAX <= HFR_PAY_MAXK;   // 16 B x K x 256 threads of payload cache next to the 30 KB tile state
  extern __shared__ __align__(16) float4 s_pay[];   // dynamic (K = 8: 32 KB, beyond the static limit): PAY ? KMAX * 256 : 0 entries
  PixelCtx c;
  bool empty_tile;
  int run = 1;
  if (queue) {
    // cost-ordered queue (raster.cu): block i takes the next non-empty tile (heaviest class first) or the next group of
    // empty tiles, the two kinds spread evenly so that the ALU-bound tiles and the HBM-bound fills share the SMs for the
    // whole launch
    const uint32_t nc = __ldg(queue), ng = __ldg(queue + 1), i = blockIdx.x, tot = nc + ng;
    if (i >= tot) return;
    const uint32_t ci = (uint32_t)(((uint64_t)i * nc) / tot), ci1 = (uint32_t)(((uint64_t)(i + 1) * nc) / tot);
    empty_tile = ci1 == ci;
    const int TX = (r.W + kTileW - 1) / kTileW, TY = (r.H + kTileH - 1) / kTileH;
    const uint32_t T = (uint32_t)(r.N * TX * TY);
    uint32_t t;
    if (empty_tile) {
      const uint32_t g = __ldg(queue + kQueueHdr + (size_t)(1 + kCostClasses) * T + (i - ci));
      t = g & 0x0fffffffu;
      run = (int)(g >> 28) + 1;
    } else {
      uint32_t left = ci;
      int k = kCostClasses - 1;
      for (; k > 0; --k) {
        const uint32_t ck = __ldg(queue + 2 + k);
        if (left < ck) break;
        left -= ck;
      }
      t = __ldg(queue + kQueueHdr + T + (size_t)k * T + left);
    }
    const int n = (int)(t / (uint32_t)(TX * TY)), rem = (int)(t - (uint32_t)n * (uint32_t)(TX * TY));
    c = make_pixel_ctx_at(n, rem % TX, rem / TX, r.H, r.W);
  } else {
    c = make_pixel_ctx(r.H, r.W);
    empty_tile = (HFR_FAST_FILL || (HFR_FAST_FILL_K1 && (r.K == 1 || (r.K & 3) == 0))) && tile_outside_mesh(mesh_box, c.n, c.tx, c.ty);
  }
  if (empty_tile) {
    // no face touches these tiles: Fragments are streamed out as whole rows, the pixels are the background
    const bool ones = s.p.blend == HFR_BLEND_SIGMOID_ALPHA;
    const float bg[3] = {ones ? 1.0f : s.p.background[0], ones ? 1.0f : s.p.background[1], ones ? 1.0f : s.p.background[2]};
    if (HFR_FILL_TMA && queue && fill_empty_tile_bulk(r, s.image, bg, c.n, c.tx, c.ty, run, reinterpret_cast<uint32_t*>(sm.rec))) return;
    for (int u = 0; u < run; ++u) {
      PixelCtx cu = u == 0 ? c : make_pixel_ctx_at(c.n, c.tx + u, c.ty, r.H, r.W);
      if ((r.K == 1 || (r.K & 3) == 0) && fill_empty_tile(r, cu.n, cu.tx, cu.ty)) {
        const size_t pix = ((size_t)cu.n * r.H + cu.yi) * r.W + cu.xi;
        *reinterpret_cast<float4*>(s.image + pix * 4) = make_float4(bg[0], bg[1], bg[2], 0.0f);
      } else if (cu.pix_active) {   // clipped / unaligned tile: per-pixel fill
        const size_t pix = ((size_t)cu.n * r.H + cu.yi) * r.W + cu.xi;
        TopK<KMAX> none;
        none.init();
        float rgba[4];
        raster_shade_epilogue<KMAX, false>(r, s, cu, none, pix, nullptr, kPermIdentity, rgba);
        *reinterpret_cast<float4*>(s.image + pix * 4) = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
      }
    }
    return;
  }
  TopK<KMAX> top;
  uint32_t perm;
  raster_tile<KMAX, PAY>(r, ranges, mesh_box, sm, c, top, s_pay + threadIdx.x, perm);
  if (!c.pix_active) return;
  const size_t pix = ((size_t)c.n * r.H + c.yi) * r.W + c.xi;
  float rgba[4];
  raster_shade_epilogue<KMAX, PAY>(r, s, c, top, pix, s_pay + threadIdx.x, perm, rgba);
  *reinterpret_cast<float4*>(s.image + pix * 4) = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
}

// Fused rasterize + shade + SSAA pool (the reference's render setting: 672^2, K=1, avg_pool2d(3,3), output split,
// models_res_nimble.py:208-220).  One CTA owns a 16x16 tile of POOLED pixels = aa x aa rasterizer tiles, which it
// walks one after the other with the same tile core; each tile's RGBA goes through 4 KB of shared memory into the
// pooled accumulators (fixed summation order, so the result is deterministic), and only Fragments plus the
// pooled outputs reach HBM - the (N, 672, 672, 4) image never exists.
struct PoolOut {
  int aa, binarize;
  const float* images_in;
  float* pooled; float* re_img; float* re_sil; float* mask_rgbs;
};
template <int KMAX>
__global__ void __launch_bounds__(kRasterThreads, HFR_RASTER_MINB_FOR(KMAX)) raster_shade_pool_fwd_kernel(
    HfrRasterArgs r, HfrShadeFwdArgs s, PoolOut po, const uint32_t* __restrict__ ranges, const uint32_t* __restrict__ mesh_box) {
  __shared__ RasterSmem sm;
  __shared__ float4 s_tile[kTileH * kTileW];
  constexpr bool PAY = KMAX <= (HFR_PAY_MAXK < 2 ? HFR_PAY_MAXK : 2);   // static shared memory stays under 48 KB
  __shared__ float4 s_pay[PAY ? KMAX * kRasterThreads : 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int aa = po.aa, n = blockIdx.z;
  const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);   // this thread's pixel inside a raster tile
  const int px = tid & 15, py = tid >> 4;                                            // this thread's pooled pixel
  const int gx0 = (blockIdx.x * kTileW + px) * aa, gy0 = (blockIdx.y * kTileH + py) * aa;   // its window's corner
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool ones = s.p.blend == HFR_BLEND_SIGMOID_ALPHA;
  const float bg0 = ones ? 1.0f : s.p.background[0], bg1 = ones ? 1.0f : s.p.background[1], bg2 = ones ? 1.0f : s.p.background[2];
  for (int sy = 0; sy < aa; ++sy) {
    for (int sx = 0; sx < aa; ++sx) {
      PixelCtx c;
      c.n = n; c.tx = blockIdx.x * aa + sx; c.ty = blockIdx.y * aa + sy;
      // the part of this thread's aa x aa window that lies in this tile: [xa, xb) x [ya, yb)
      const int tx0 = c.tx * kTileW, ty0 = c.ty * kTileH;
      const int xa = max(gx0, tx0), xb = min(gx0 + aa, tx0 + kTileW), ya = max(gy0, ty0), yb = min(gy0 + aa, ty0 + kTileH);
      c.xi = tx0 + lx; c.yi = ty0 + ly;
      c.pix_active = c.xi < r.W && c.yi < r.H;
      if (HFR_FAST_FILL_POOL && tile_outside_mesh(mesh_box, n, c.tx, c.ty) && fill_empty_tile(r, n, c.tx, c.ty)) {
        // no face touches this tile: rows of -1 Fragments, and the window's share of it is plain background
        if (s.image) *reinterpret_cast<float4*>(s.image + (((size_t)n * r.H + c.yi) * r.W + c.xi) * 4) = make_float4(bg0, bg1, bg2, 0.0f);
        const float cntw = (float)(max(xb - xa, 0) * max(yb - ya, 0));   // window pixels in this tile, all background
        acc.x += cntw * bg0; acc.y += cntw * bg1; acc.z += cntw * bg2;
        continue;
      }
      c.warp_active = tx0 + (warp & 1) * 8 < r.W && ty0 + (warp >> 1) * 4 < r.H;
      c.xf = 0.0f; c.yf = 0.0f;
      TopK<KMAX> top;
      uint32_t perm;
      raster_tile<KMAX, PAY>(r, ranges, mesh_box, sm, c, top, s_pay + tid, perm);
      float rgba[4] = {0.f, 0.f, 0.f, 0.f};
      if (c.pix_active) {
        const size_t pix = ((size_t)n * r.H + c.yi) * r.W + c.xi;
        raster_shade_epilogue<KMAX, PAY>(r, s, c, top, pix, s_pay + tid, perm, rgba);
        if (s.image) *reinterpret_cast<float4*>(s.image + pix * 4) = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
      }
      s_tile[ly * kTileW + lx] = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
      __syncthreads();
      for (int y = ya; y < yb; ++y)
        for (int x = xa; x < xb; ++x) {
          const float4 v = s_tile[(y - ty0) * kTileW + (x - tx0)];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      __syncthreads();   // s_tile and the tile core's shared state are reused by the next tile
    }
  }
  const int Hp = r.H / aa, Wp = r.W / aa;
  const int ox = blockIdx.x * kTileW + px, oy = blockIdx.y * kTileH + py;
  if (ox >= Wp || oy >= Hp) return;
  const float cnt = (float)(aa * aa);
  const float c0 = acc.x / cnt, c1 = acc.y / cnt, c2 = acc.z / cnt, al = acc.w / cnt;
  const float sil = (po.binarize && al > 0.0f) ? 255.0f : al;
  const size_t hw = (size_t)Hp * Wp, p = (size_t)oy * Wp + ox;
  *reinterpret_cast<float4*>(po.pooled + ((size_t)n * hw + p) * 4) = make_float4(c0, c1, c2, sil);
  if (po.re_img) {
    po.re_img[((size_t)n * 3 + 0) * hw + p] = c0; po.re_img[((size_t)n * 3 + 1) * hw + p] = c1; po.re_img[((size_t)n * 3 + 2) * hw + p] = c2;
  }
  if (po.re_sil) po.re_sil[(size_t)n * hw + p] = sil;
  if (po.mask_rgbs) {
    const float m = sil > 0.0f ? 1.0f : 0.0f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) po.mask_rgbs[((size_t)n * 3 + ch) * hw + p] = __ldg(po.images_in + ((size_t)n * 3 + ch) * hw + p) * m;
  }
}

// One thread per (mesh, face, 128-bit unit of the record): 7 coalesced 16-byte stores per face.
__global__ void __launch_bounds__(256) face_attr_kernel(HfrFaceAttrArgs a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)a.N * a.F * 7;
  if (i >= total) return;
  const int u = (int)(i % 7);
  const size_t nf = i / 7;
  const int f = (int)(nf % a.F), n = (int)(nf / a.F);
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int w = 4 * u + e;          // word of the record
    float x = 0.0f;
    if (w < 18) {                     // positions then normals, corner-major xyz
      const int q = w < 9 ? w : w - 9, corner = q / 3, c = q - 3 * corner;
      const int vid = __ldg(a.faces + 3 * f + corner);
      x = __ldg((w < 9 ? a.verts_view : a.vnormals) + ((size_t)n * a.V + vid) * 3 + c);
    } else if (w < 24) {
      const int q = w - 18, corner = q >> 1;
      x = __ldg(a.verts_uvs + 2 * __ldg(a.faces_uvs + 3 * f + corner) + (q & 1));
    } else if (w < 27) {
      x = __int_as_float(__ldg(a.faces + 3 * f + (w - 24)));
    }
    v[e] = x;
  }
  *reinterpret_cast<float4*>(a.face_attr + nf * HFR_FACE_ATTR_FLOATS + 4 * u) = make_float4(v[0], v[1], v[2], v[3]);
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
  HFR_CHECK_ARG(p.tex_pca >= 0 && p.tex_pca <= HFR_MAX_TEX_PCA, "%s: tex_pca must be in [0,%d]", who, HFR_MAX_TEX_PCA);
  HFR_CHECK_ARG(p.tex_basis_stride == 0 || (p.tex_pca > 0 && p.tex_basis_stride == 12 * ((p.tex_pca + 3) / 4) &&
                                            (reinterpret_cast<uintptr_t>(a->tex_basis) & 15) == 0),
                "%s: a texel-major basis has stride 12 * ceil(tex_pca / 4) floats and a 16-byte aligned base", who);
  if (p.shade == HFR_SHADE_PHONG_UV && p.tex_pca > 0)
    HFR_CHECK_ARG(p.tex_n == 1 && a->tex_basis && a->tex_params, "%s: a PCA texture needs the mean map (tex_n = 1), basis and coefficients", who);
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
#define CALL(KM)                                                                          \
  do {                                                                                    \
    if (a->p.tex_pca > 0) shade_fwd_kernel<KM, true><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);   \
    else shade_fwd_kernel<KM, false><<<grid, 256, 0, (cudaStream_t)stream>>>(*a);         \
  } while (0)
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
  HFR_CHECK_ARG(s.p.tex_pca == 0, "raster_shade_forward: PCA textures are shaded by hfr_shade_forward (rasterize with hfr_raster_forward)");
  if (a->r.N == 0) return HFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* ranges = reinterpret_cast<uint32_t*>(a->r.workspace);
  const uint32_t* box = raster_mesh_box(a->r);
  const bool use_queue = a->r.tile_queue != nullptr && box != nullptr;
  if (int rc = launch_raster_setup(a->r, ranges, st, use_queue)) return rc;
  dim3 grid((a->r.W + kTileW - 1) / kTileW, (a->r.H + kTileH - 1) / kTileH, a->r.N);
  const uint32_t* queue = use_queue ? reinterpret_cast<const uint32_t*>(a->r.tile_queue) : nullptr;
  if (use_queue) grid = dim3(grid.x * grid.y * grid.z, 1, 1);
#define CALL(KM)                                                                                              \
  do {                                                                                                        \
    const size_t pay = (KM) <= HFR_PAY_MAXK ? (size_t)(KM) * kRasterThreads * sizeof(float4) : 0;             \
    if (pay + sizeof(RasterSmem) > 48 * 1024)                                                                 \
      cudaFuncSetAttribute(raster_shade_fwd_kernel<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pay); \
    raster_shade_fwd_kernel<KM><<<grid, kRasterThreads, pay, st>>>(a->r, s, ranges, box, queue);             \
  } while (0)
  HFR_DISPATCH_K(a->r.K, CALL);
#undef CALL
  HFR_CHECK_LAUNCH("raster_shade_forward");
  return HFR_OK;
}

extern "C" int hfr_raster_shade_pool_forward(const HfrRasterShadePoolArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a, "raster_shade_pool_forward: null args");
  if (int rc = check_raster(&a->r, "raster_shade_pool_forward")) return rc;
  HfrShadeFwdArgs s = a->s;
  s.pix_to_face = a->r.pix_to_face; s.zbuf = a->r.zbuf; s.bary = a->r.bary; s.dists = a->r.dists;
  if (int rc = check_shade(&s, "raster_shade_pool_forward")) return rc;
  HFR_CHECK_ARG(s.p.N == a->r.N && s.p.H == a->r.H && s.p.W == a->r.W && s.p.K == a->r.K,
                "raster_shade_pool_forward: raster / shade dims differ");
  HFR_CHECK_ARG(a->aa >= 1 && a->aa <= 16 && a->r.H % a->aa == 0 && a->r.W % a->aa == 0,
                "raster_shade_pool_forward: image size must be a multiple of aa (1..16)");
  HFR_CHECK_ARG(a->r.N == 0 || a->pooled, "raster_shade_pool_forward: null pooled image");
  HFR_CHECK_ARG(s.p.tex_pca == 0, "raster_shade_pool_forward: PCA textures are shaded by hfr_shade_forward");
  HFR_CHECK_ARG(!a->mask_rgbs || a->images_in, "raster_shade_pool_forward: mask_rgbs needs images_in");
  if (a->r.N == 0) return HFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* ranges = reinterpret_cast<uint32_t*>(a->r.workspace);
  if (int rc = launch_raster_setup(a->r, ranges, st)) return rc;
  const int Hp = a->r.H / a->aa, Wp = a->r.W / a->aa;
  dim3 grid((Wp + kTileW - 1) / kTileW, (Hp + kTileH - 1) / kTileH, a->r.N);
  PoolOut po{a->aa, a->binarize, a->images_in, a->pooled, a->re_img, a->re_sil, a->mask_rgbs};
#define CALL(KM) raster_shade_pool_fwd_kernel<KM><<<grid, kRasterThreads, 0, st>>>(a->r, s, po, ranges, raster_mesh_box(a->r))
  HFR_DISPATCH_K(a->r.K, CALL);
#undef CALL
  HFR_CHECK_LAUNCH("raster_shade_pool_forward");
  return HFR_OK;
}

extern "C" int hfr_face_attr_forward(const HfrFaceAttrArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->N >= 0 && a->F > 0 && a->V > 0 && a->VT > 0, "face_attr_forward: bad dims");
  if (a->N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->faces && a->verts_view && a->vnormals && a->faces_uvs && a->verts_uvs && a->face_attr,
                "face_attr_forward: null pointer");
  const size_t total = (size_t)a->N * a->F * 7;
  face_attr_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("face_attr_forward");
  return HFR_OK;
}

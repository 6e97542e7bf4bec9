// Atomics-free, deterministic fused backward of shading + blending + rasterization for sm_100a.
//
// Replaces the scatter of upstream rasterize_meshes_backward / interp_face_attrs_backward / grid_sampler_2d_backward
// (9 + 18 + 12 floating-point atomics per covered pixel, reached by loss.backward() at train_hrnet.py:112) with
//
//   1. per 16x16 tile (one CTA): the tile's fragments are counting-sorted BY FACE in shared memory.  All bookkeeping
//      is integer: a hash set of the faces present (atomicCAS), per face a 256-bit mask of the pixels holding it
//      (atomicOr), offsets by popcount prefix.  Inside a face the fragments keep pixel order, whatever order the
//      hardware executed the atomics in.
//   2. the sorted fragments are differentiated 32 at a time with DENSE warps (a boundary tile with 20 % covered pixels
//      fills its lanes as well as an interior one) and face-coherent lanes (the three corners' attributes are
//      broadcast loads), each lane producing the fragment's 18 per-face components: 3 corners x {d/d(view position)
//      with d(ndc)/d(view) folded in, d/d(vertex normal)} plus its 6 light-gradient components.
//   3. a warp owns whole faces: lane c adds component c of the face's fragments ONE AFTER THE OTHER in pixel order
//      (a sequential fp32 chain, so the sum does not depend on how the fragments fell into 32-lane chunks) and writes
//      the (face, tile) record into the slot the rasterizer's setup pass reserved for it.  No atomics, no races.
//   4. sums over all fragments of a sample / of the batch (light gradients, the shared texture's gradient) use 64-bit
//      fixed-point accumulators - integer addition is associative - converted to fp32 by hfr_grad_finish.
//
// hfr_geom_backward gathers the records per vertex (static incidence lists, tiles of a face's range row by row).
#include "common.cuh"
#include "raster_math.cuh"
#include "raster_tile.cuh"
#include "shade_pixel.cuh"

namespace hfr {

constexpr int kBT = 256;         // threads = pixels of one 16x16 tile; thread t <-> pixel (warp = 8x4 block, as the forward)
constexpr int kBW = kBT / 32;
constexpr int kCap = 256;        // distinct faces per pass (a tile with more is processed in several passes)
constexpr int kNC = 24;          // components summed per face: 18 vertex + 3 light direction / location + 3 light colour
constexpr uint16_t kNoFrag = 0xffffu;

template <int KMAX>
struct TCfg {
  static constexpr int HT = KMAX <= 1 ? 512 : (KMAX <= 2 ? 1024 : (KMAX <= 4 ? 2048 : (KMAX <= 8 ? 4096 : 8192)));   // >= 2 x 256 K
};

template <int KMAX>
struct TSmem {
  int keys[TCfg<KMAX>::HT];               // hash set of the faces (local id) present in the tile, -1 = empty
  uint16_t slotmap[TCfg<KMAX>::HT];       // table entry -> dense slot
  uint32_t masks[kCap][8];                // per slot of the pass: which of the 256 pixels hold that face
  int offs[kCap + 1];                     // exclusive prefix of the slots' fragment counts
  uint16_t sorted[KMAX * kBT];            // fragments (pixel | k << 8) grouped by slot, pixel order inside a slot
  uint16_t fragh[KMAX][kBT];              // table entry of every fragment (kNoFrag: carries no gradient)
  float pix[8][kBT];                      // per pixel: gnum[3], gden, g_alpha, gzmax, kmax (int bits), spare
  float frag[KMAX][3][kBT];               // per fragment: sigmoid prob, softmax exponent, prod_{j != k} (1 - p_j)
  float stage[kBW][kNC][33];              // per warp: the chunk's components, transposed for the per-face chains
  float tabx[kTileW], taby[kTileH];       // NDC sample positions of the tile's columns / rows
  int wsum[kBW];
  int wstart[kBW + 1];
  unsigned long long light[6];            // fixed-point light sums of the tile
  int misc[4];
};

__device__ __forceinline__ uint32_t hash_face(int f) { return (uint32_t)f * 2654435761u; }

// block-wide exclusive scan of one int per thread (kBT threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_excl_scan(int v, int* wsum, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();                 // wsum free (previous use consumed)
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kBW; ++w) { const int c = wsum[w]; before += (w < warp) ? c : 0; tot += c; }
  *total = tot;
  return before + incl - v;
}

template <int KMAX>
__global__ void __launch_bounds__(kBT, (KMAX <= 1 ? 4 : (KMAX <= 4 ? 3 : (KMAX <= 8 ? 2 : 1))))
shade_bwd_tiled_kernel(HfrShadeBwdTiledArgs a, WsLayout L) {
  extern __shared__ __align__(16) unsigned char smraw[];
  TSmem<KMAX>& sm = *reinterpret_cast<TSmem<KMAX>*>(smraw);
  constexpr int HT = TCfg<KMAX>::HT;
  const HfrShadeFwdArgs& f = a.f;
  const HfrShadeParams& P = f.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, tx = blockIdx.x, ty = blockIdx.y, K = P.K, V = P.V;
  const uint32_t* __restrict__ ws = reinterpret_cast<const uint32_t*>(a.raster_ws);
  {   // tile outside this mesh's footprint: no fragment, no record
    const uint4 bx = __ldg(reinterpret_cast<const uint4*>(ws + L.box) + n);
    if (tx < (int)bx.x || tx > 255 - (int)bx.y || ty < (int)bx.z || ty > 255 - (int)bx.w) return;
  }
  const int xi = tx * kTileW + (warp & 1) * 8 + (lane & 7);
  const int yi = ty * kTileH + (warp >> 1) * 4 + (lane >> 3);
  const bool active = xi < P.W && yi < P.H;
  const bool phong = P.shade == HFR_SHADE_PHONG_UV;
  const int kshade = phong ? (P.blend == HFR_BLEND_SOFTMAX ? K : 1) : 0;
  const size_t pix = ((size_t)n * P.H + yi) * P.W + xi;

  for (int i = tid; i < HT; i += kBT) sm.keys[i] = -1;
  if (tid < kTileW) sm.tabx[tid] = hfr_pix_to_ndc(P.W - 1 - (tx * kTileW + tid), P.W, P.H);
  else if (tid < kTileW + kTileH) sm.taby[tid - kTileW] = hfr_pix_to_ndc(P.H - 1 - (ty * kTileH + tid - kTileW), P.H, P.W);
  if (tid < 6) sm.light[tid] = 0ull;
  __syncthreads();

  // ================================================================ 1. per pixel: fragments, blend state
  int fl[KMAX];
  unsigned vmask = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) fl[k] = -1;
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
  if (!__syncthreads_or(any)) return;      // no fragment in the whole tile
  {
    float z[KMAX], d[KMAX], prob[KMAX], wexp[KMAX], others[KMAX];
    float gnum[3] = {0.f, 0.f, 0.f}, gden = 0.f, gzmax = 0.f, g_alpha = 0.f;
    int kmax = -1;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { z[k] = -1.f; d[k] = -1.f; prob[k] = 0.f; wexp[k] = 0.f; others[k] = 1.f; }
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
        gnum[0] = g4.x; gnum[1] = g4.y; gnum[2] = g4.z;
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
          gnum[0] = g4.x; gnum[1] = g4.y; gnum[2] = g4.z;
        } else {
          // softmax blend, differentiated from the stored forward pixel (see shade_bwd.cu)
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
          for (int k = 0; k < KMAX; ++k)
            if (k < K) { wexp[k] = HFR_EXP(HFR_FDIV(zinv[k] - zmax, P.gamma)); wsum += prob[k] * wexp[k]; }
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
            gacc += gnum[c] * (rgb[c] * den - delta * P.background[c]);
          }
          gdelta += gden;
          gzmax = -HFR_FDIV(gden * wsum + gacc, P.gamma);
          if (dexp >= eps) gzmax -= HFR_FDIV(gdelta * delta, P.gamma);
          if (!(zmax_raw >= eps)) kmax = -1;
        }
      }
    }
    sm.pix[0][tid] = gnum[0]; sm.pix[1][tid] = gnum[1]; sm.pix[2][tid] = gnum[2]; sm.pix[3][tid] = gden;
    sm.pix[4][tid] = g_alpha; sm.pix[5][tid] = gzmax; sm.pix[6][tid] = __int_as_float(kmax);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      sm.frag[k][0][tid] = prob[k]; sm.frag[k][1][tid] = wexp[k]; sm.frag[k][2][tid] = others[k];
      uint16_t hh = kNoFrag;
      if (((vmask >> k) & 1u) && (P.blend != HFR_BLEND_HARD || k == 0)) {   // hidden slots of a hard blend carry no gradient
        uint32_t h = hash_face(fl[k]) & (HT - 1);
        while (true) {
          const int old = atomicCAS(&sm.keys[h], -1, fl[k]);
          if (old == -1 || old == fl[k]) break;
          h = (h + 1) & (HT - 1);
        }
        hh = (uint16_t)h;
      }
      sm.fragh[k][tid] = hh;
    }
  }
  __syncthreads();
  // dense slot numbers for the occupied table entries (their order is irrelevant: a face's sum is a sequential
  // chain over ITS fragments in pixel order, wherever the face sits in the sorted list)
  int D = 0;
  {
    constexpr int PER = HT / kBT;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) cnt += sm.keys[tid * PER + j] >= 0 ? 1 : 0;
    int q = block_excl_scan(cnt, sm.wsum, &D);
#pragma unroll
    for (int j = 0; j < PER; ++j)
      if (sm.keys[tid * PER + j] >= 0) sm.slotmap[tid * PER + j] = (uint16_t)(q++);
  }
  __syncthreads();

  // per-sample light constants
  float dhat[3] = {0.f, 0.f, 0.f}, dlen = 1.f, lcol[3] = {0.f, 0.f, 0.f};
  if (phong) {
    light_dir_hat(f, n, dhat, &dlen);
    lcol[0] = __ldg(f.light_color + 3 * n); lcol[1] = __ldg(f.light_color + 3 * n + 1); lcol[2] = __ldg(f.light_color + 3 * n + 2);
  }
  const float fxs = __ldg(a.fx_scale);
  const float fcx = __ldg(a.focal + 2 * n), fcy = __ldg(a.focal + 2 * n + 1);
  const size_t tbase = (P.tex_n == 1 ? 0 : (size_t)n * P.tex_h * P.tex_w * 3);
  const float zr = P.zfar - P.znear;

  for (int p0 = 0; p0 < D; p0 += kCap) {
    const int Dp = min(kCap, D - p0);
    // ============================================================== 2. counting sort of the fragments by face
    for (int i = tid; i < Dp * 8; i += kBT) (&sm.masks[0][0])[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const uint16_t hh = sm.fragh[k][tid];
      if (hh != kNoFrag) {
        const int s = (int)sm.slotmap[hh] - p0;
        if ((unsigned)s < (unsigned)Dp) atomicOr(&sm.masks[s][warp], 1u << lane);
      }
    }
    __syncthreads();
    int nfr = 0;
    {
      int c = 0;
      if (tid < Dp) {
#pragma unroll
        for (int w = 0; w < 8; ++w) c += __popc(sm.masks[tid][w]);
      }
      const int ex = block_excl_scan(c, sm.wsum, &nfr);
      if (tid < Dp) sm.offs[tid] = ex;
      if (tid == 0) sm.offs[Dp] = nfr;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const uint16_t hh = sm.fragh[k][tid];
      if (hh != kNoFrag) {
        const int s = (int)sm.slotmap[hh] - p0;
        if ((unsigned)s < (unsigned)Dp) {
          int rank = __popc(sm.masks[s][warp] & ((1u << lane) - 1u));
          for (int w = 0; w < warp; ++w) rank += __popc(sm.masks[s][w]);
          sm.sorted[sm.offs[s] + rank] = (uint16_t)(tid | (k << 8));
        }
      }
    }
    // a warp owns WHOLE faces: those whose first fragment falls into its eighth of the sorted list
    if (tid <= kBW) {
      int st = nfr;
      if (tid < kBW) {
        const int target = (int)(((long long)tid * nfr) / kBW);
        int lo = 0, hi = Dp;                 // first slot with offs >= target (offs is strictly increasing)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sm.offs[mid] >= target) hi = mid; else lo = mid + 1;
        }
        st = sm.offs[lo];
      }
      sm.wstart[tid] = st;
    }
    __syncthreads();

    // ============================================================== 3. differentiate, per-face sequential sums
    const int beg = sm.wstart[warp], end = sm.wstart[warp + 1];
    float tot = 0.0f;                 // lane c < kNC: running sum of component c of the open face
    long long open_rec = -1;          // record slot of the open face (warp-uniform)
    int prev_slot = -1;
    for (int base = beg; base < end; base += 32) {
      const int i = base + lane;
      const bool valid = i < end;
      int slot = -2;
      long long rec = -1;
      float v[kNC];
#pragma unroll
      for (int c = 0; c < kNC; ++c) v[c] = 0.0f;
      if (valid) {
        const int e = sm.sorted[i], p = e & 255, k = e >> 8;
        const uint16_t hh = sm.fragh[k][p];
        slot = sm.slotmap[hh];
        const int face = sm.keys[hh];
        rec = (long long)face_rec_index(ws, L, (int64_t)n * P.F + face, tx, ty);
        const int pw = p >> 5, pl = p & 31;
        const int lx = (pw & 1) * 8 + (pl & 7), ly = (pw >> 1) * 4 + (pl >> 3);
        const size_t fpix = ((size_t)n * P.H + (ty * kTileH + ly)) * P.W + (tx * kTileW + lx);
        const float xf = sm.tabx[lx], yf = sm.taby[ly];
        const float* __restrict__ bp = f.bary + (fpix * K + k) * 3;
        const float bc[3] = {__ldg(bp), __ldg(bp + 1), __ldg(bp + 2)};
        const float pk = sm.frag[k][0][p], ek = sm.frag[k][1][p];
        const float gnum[3] = {sm.pix[0][p], sm.pix[1][p], sm.pix[2][p]};
        const float gden = sm.pix[3][p], gzmax = sm.pix[5][p];
        const int kmax = __float_as_int(sm.pix[6][p]);
        float gprob = sm.pix[4][p] * sm.frag[k][2][p], gz = 0.f;
        float g_bc[3] = {0.f, 0.f, 0.f}, col[3] = {1.0f, 1.0f, 1.0f}, gcol[3];
        FragGeom g;
        HfrTexTap tap; HfrPhongCtx ctx; float texel[3];
        const bool shaded = k < kshade;
        if (shaded) {
          gather_frag(f, n, face, g);
          shade_fragment<false>(f, n, g, bc, dhat, lcol, col, &tap, &ctx, texel);
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            g.vid[c] = __ldg(f.faces + 3 * face + c);
            const float* __restrict__ x = f.verts_view + ((size_t)n * V + g.vid[c]) * 3;
            g.X[3 * c] = __ldg(x); g.X[3 * c + 1] = __ldg(x + 1); g.X[3 * c + 2] = __ldg(x + 2);
          }
        }
        if (P.blend == HFR_BLEND_SOFTMAX) {
          const float wk = pk * ek;
          const float gw = gden + gnum[0] * col[0] + gnum[1] * col[1] + gnum[2] * col[2];
          gcol[0] = wk * gnum[0]; gcol[1] = wk * gnum[1]; gcol[2] = wk * gnum[2];
          gprob += gw * ek;
          const float gzinv = HFR_FDIV(gw * wk, P.gamma) + (k == kmax ? gzmax : 0.f);
          gz = HFR_FDIV(-gzinv, zr);
        } else {
          gcol[0] = gnum[0]; gcol[1] = gnum[1]; gcol[2] = gnum[2];
        }
        if (shaded) {
          float gP[3], gNn[3], gtex[3], gdh[3] = {0.f, 0.f, 0.f}, glc[3] = {0.f, 0.f, 0.f}, ld[3];
          hfr_phong_bwd(P, ctx.lhat, lcol, texel, &ctx, gcol, gP, gNn, gtex, gdh, glc);
          if (P.light_point) {   // direction = location - P
            hfr_normalize_eps_bwd(ctx.lhat, ctx.llen, gdh, ld);
            gP[0] -= ld[0]; gP[1] -= ld[1]; gP[2] -= ld[2];
          } else {
            hfr_normalize_eps_bwd(dhat, dlen, gdh, ld);   // linear in gdh: applied per fragment, summed afterwards
          }
          v[18] = ld[0]; v[19] = ld[1]; v[20] = ld[2]; v[21] = glc[0]; v[22] = glc[1]; v[23] = glc[2];
          float gu = 0.f, gv = 0.f;
          hfr_tex_uv_grad(tex_source<false>(f, n), &tap, gtex, &gu, &gv);
          if (a.tex_acc) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (tap.idx[q] >= 0) {
                unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.tex_acc) + tbase + (size_t)tap.idx[q] * 3;
                const float wq = tap.w[q] * fxs;
                atomicAdd(dst, (unsigned long long)__float2ll_rn(wq * gtex[0]));
                atomicAdd(dst + 1, (unsigned long long)__float2ll_rn(wq * gtex[1]));
                atomicAdd(dst + 2, (unsigned long long)__float2ll_rn(wq * gtex[2]));
              }
            }
          } else if (a.g_texture) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (tap.idx[q] >= 0) {
                float* dst = a.g_texture + tbase + (size_t)tap.idx[q] * 3;
                atomicAdd(dst, tap.w[q] * gtex[0]); atomicAdd(dst + 1, tap.w[q] * gtex[1]); atomicAdd(dst + 2, tap.w[q] * gtex[2]);
              }
            }
          }
#pragma unroll
          for (int c3 = 0; c3 < 3; ++c3) {
            g_bc[c3] = gP[0] * g.X[3 * c3] + gP[1] * g.X[3 * c3 + 1] + gP[2] * g.X[3 * c3 + 2] +
                       gNn[0] * g.Nv[3 * c3] + gNn[1] * g.Nv[3 * c3 + 1] + gNn[2] * g.Nv[3 * c3 + 2] +
                       gu * g.uv[2 * c3] + gv * g.uv[2 * c3 + 1];
#pragma unroll
            for (int c = 0; c < 3; ++c) { v[6 * c3 + c] = bc[c3] * gP[c]; v[6 * c3 + 3 + c] = bc[c3] * gNn[c]; }
          }
        }
        const float gd = P.blend == HFR_BLEND_HARD ? 0.f : HFR_FDIV(-gprob * pk * (1.0f - pk), P.sigma);
        // rasterizer backward on the face's NDC vertices, then d(ndc)/d(view) per corner:
        //   x = fx X / Z + px,  y = fy Y / Z + py,  z = Z
        float vv[9], gvv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          const float* __restrict__ src = a.verts_ndc + ((size_t)n * V + g.vid[c3]) * 3;
          vv[3 * c3] = __ldg(src); vv[3 * c3 + 1] = __ldg(src + 1); vv[3 * c3 + 2] = __ldg(src + 2);
        }
        hfr_raster_eval_bwd(xf, yf, vv, a.perspective_correct, a.clip_barycentric, g_bc, gz, gd, gvv);
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          const float X = g.X[3 * c3], Y = g.X[3 * c3 + 1], Z = g.X[3 * c3 + 2];
          const float iz = HFR_RCP(Z);
          const float ax = gvv[3 * c3] * fcx * iz, ay = gvv[3 * c3 + 1] * fcy * iz;
          v[6 * c3] += ax; v[6 * c3 + 1] += ay;
          v[6 * c3 + 2] += gvv[3 * c3 + 2] - (ax * X + ay * Y) * iz;
        }
      }
      // stage the components, then lane c walks the chunk's members of each face in order
#pragma unroll
      for (int c = 0; c < kNC; ++c) sm.stage[warp][c][lane] = v[c];
      __syncwarp();
      const int up = __shfl_up_sync(0xffffffffu, slot, 1);
      const bool starts = valid && (lane == 0 ? slot != prev_slot : slot != up);
      const unsigned sb = __ballot_sync(0xffffffffu, starts);
      const int cnt = __popc(__ballot_sync(0xffffffffu, valid));
      int m = 0;
      while (m < cnt) {
        if ((sb >> m) & 1u) {      // a new face starts at member m: write out the open one
          if (open_rec >= 0) {
            if (lane < HFR_FACE_REC_FLOATS) {
              if (open_rec < a.rec_cap) a.face_rec[open_rec * HFR_FACE_REC_FLOATS + lane] = tot;
              else if (lane == 0) atomicOr(a.status, 1u);
            } else if (lane < kNC && tot != 0.0f) {
              atomicAdd(&sm.light[lane - HFR_FACE_REC_FLOATS], (unsigned long long)__float2ll_rn(tot * fxs));
            }
          }
          tot = 0.0f;
          open_rec = __shfl_sync(0xffffffffu, rec, m);
        }
        const unsigned nxt = sb & ~((2u << m) - 1u);
        const int m1 = nxt ? __ffs(nxt) - 1 : cnt;
        if (lane < kNC)
          for (int j = m; j < m1; ++j) tot += sm.stage[warp][lane][j];
        m = m1;
      }
      prev_slot = __shfl_sync(0xffffffffu, slot, cnt - 1);
      __syncwarp();
    }
    if (open_rec >= 0) {
      if (lane < HFR_FACE_REC_FLOATS) {
        if (open_rec < a.rec_cap) a.face_rec[open_rec * HFR_FACE_REC_FLOATS + lane] = tot;
        else if (lane == 0) atomicOr(a.status, 1u);
      } else if (lane < kNC && tot != 0.0f) {
        atomicAdd(&sm.light[lane - HFR_FACE_REC_FLOATS], (unsigned long long)__float2ll_rn(tot * fxs));
      }
    }
    __syncthreads();     // the next pass reuses masks / sorted
  }
  if (tid < 6 && a.light_acc && sm.light[tid] != 0ull)
    atomicAdd(reinterpret_cast<unsigned long long*>(a.light_acc) + (size_t)n * 6 + tid, sm.light[tid]);
}

// clears the used part of the record store: [0, min(total, cap)) records, total read from the workspace
__global__ void __launch_bounds__(256) face_rec_zero_kernel(float* __restrict__ rec, const uint32_t* __restrict__ total_ptr, int64_t cap) {
  const int64_t total = min((int64_t)__ldg(total_ptr), cap);
  const int64_t n4 = (total * HFR_FACE_REC_FLOATS + 3) / 4;      // 18 floats per record: 72 B, the store is 16-byte aligned
  float4* r4 = reinterpret_cast<float4*>(rec);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) grad_finish_kernel(HfrGradFinishArgs a) {
  const float inv = 1.0f / __ldg(a.fx_scale);     // a power of two: exact
  const int64_t nl = a.light_acc ? (int64_t)a.N * 6 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_tex + nl; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < a.n_tex) {
      const long long q = a.tex_acc[i];
      a.g_texture[i] = __ll2float_rn(q) * inv;
      if (q != 0) a.tex_acc[i] = 0;
    } else {
      const int64_t j = i - a.n_tex;
      const long long q = a.light_acc[j];
      const int n = (int)(j / 6), c = (int)(j % 6);
      const float g = __ll2float_rn(q) * inv;
      if (c < 3) { if (a.g_light_dir) a.g_light_dir[3 * n + c] = g; }
      else if (a.g_light_color) a.g_light_color[3 * n + c - 3] = g;
      if (q != 0) a.light_acc[j] = 0;
    }
  }
}

int check_shade(const HfrShadeFwdArgs* a, const char* who);

}  // namespace hfr

extern "C" int hfr_shade_backward_tiled(const HfrShadeBwdTiledArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a, "shade_backward_tiled: null args");
  if (int rc = check_shade(&a->f, "shade_backward_tiled")) return rc;
  const HfrShadeParams& p = a->f.p;
  if (p.N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->g_image && a->verts_ndc && a->focal && a->raster_ws && a->face_rec && a->rec_cap > 0 && a->fx_scale && a->status,
                "shade_backward_tiled: null pointer");
  HFR_CHECK_ARG(a->f.faces && a->f.verts_view && p.F > 0 && p.V > 0, "shade_backward_tiled: needs faces and verts_view");
  HFR_CHECK_ARG(p.blend != HFR_BLEND_SOFTMAX || a->f.image, "shade_backward_tiled: the softmax blend needs the forward image");
  HFR_CHECK_ARG(p.tex_pca == 0, "shade_backward_tiled: PCA textures are differentiated by hfr_shade_backward");
  HFR_CHECK_ARG(p.shade != HFR_SHADE_PHONG_UV || a->light_acc, "shade_backward_tiled: Phong shading needs light_acc");
  HFR_CHECK_ARG(a->pool_aa <= 1 || (a->pool_aa <= 16 && p.H % a->pool_aa == 0 && p.W % a->pool_aa == 0),
                "shade_backward_tiled: image size must be a multiple of pool_aa (<= 16)");
  HFR_CHECK_ARG((p.W + kTileW - 1) / kTileW <= 255 && (p.H + kTileH - 1) / kTileH <= 255, "shade_backward_tiled: image too large");
  HFR_CHECK_ARG((reinterpret_cast<uintptr_t>(a->face_rec) & 15) == 0, "shade_backward_tiled: face_rec must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t Ftot = (int64_t)p.N * p.F;
  const WsLayout L = ws_layout(Ftot);
  const uint32_t* ws = reinterpret_cast<const uint32_t*>(a->raster_ws);
  {
    const int64_t n4 = (a->rec_cap * HFR_FACE_REC_FLOATS + 3) / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    face_rec_zero_kernel<<<blocks, 256, 0, st>>>(a->face_rec, ws + L.blk + L.nblk, a->rec_cap);
    HFR_CHECK_LAUNCH("face_rec_zero");
  }
  dim3 grid((p.W + kTileW - 1) / kTileW, (p.H + kTileH - 1) / kTileH, p.N);
#define HFR_LAUNCH_T(KM)                                                                                         \
  do {                                                                                                           \
    const size_t smem = sizeof(TSmem<KM>);                                                                       \
    cudaFuncSetAttribute(shade_bwd_tiled_kernel<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    shade_bwd_tiled_kernel<KM><<<grid, kBT, smem, st>>>(*a, L);                                                  \
  } while (0)
  if (p.K == 1) HFR_LAUNCH_T(1);
  else if (p.K == 2) HFR_LAUNCH_T(2);
  else if (p.K <= 4) HFR_LAUNCH_T(4);
  else if (p.K <= 8) HFR_LAUNCH_T(8);
  else HFR_LAUNCH_T(16);
#undef HFR_LAUNCH_T
  HFR_CHECK_LAUNCH("shade_backward_tiled");
  return HFR_OK;
}

extern "C" int hfr_grad_finish(const HfrGradFinishArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->fx_scale, "grad_finish: null args");
  HFR_CHECK_ARG(a->n_tex == 0 || (a->tex_acc && a->g_texture), "grad_finish: texture accumulator / gradient missing");
  const int64_t total = a->n_tex + (a->light_acc ? (int64_t)a->N * 6 : 0);
  if (total <= 0) return HFR_OK;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  grad_finish_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("grad_finish");
  return HFR_OK;
}

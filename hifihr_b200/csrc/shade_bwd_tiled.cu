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
#ifndef HFR_TILED_CAP
#define HFR_TILED_CAP 128
#endif
constexpr int kCap = HFR_TILED_CAP;   // distinct faces per pass (a tile with more is processed in several passes)
constexpr int kNC = 24;          // components summed per face: 18 vertex + 3 light direction / location + 3 light colour
constexpr int kFR = 36;          // words per staged face record (144 B: LDS.128 rows of different faces spread over the banks)
constexpr uint16_t kNoFrag = 0xffffu;
constexpr int kMaxFacesTiled = 65535;   // local face ids and slots are 16-bit

template <int KMAX>
struct TSmem {
  __align__(16) float facerec[kCap][kFR]; // per face of the pass: view xyz x3, normals x3, uv x3, ndc xy x3, record slot
  float stage[kBW][kNC][33];              // per warp: the chunk's components, transposed for the per-face chains
  float part[KMAX * 8][2][kNC];           // per chunk: partial sums of a face that continues from / into a neighbour chunk
  float pix[7][kBT];                      // per pixel: gnum[3], gden, g_alpha, gzmax, kmax (int bits)
  float frag[KMAX][3][kBT];               // per fragment: sigmoid prob, softmax exponent, prod_{j != k} (1 - p_j)
  uint32_t masks[kCap][8];                // per slot of the pass: which of the 256 pixels hold that face
  int offs[kCap + 1];                     // exclusive prefix of the slots' fragment counts
  uint16_t sorted[KMAX * kBT];            // fragments (pixel | k << 8) grouped by slot, pixel order inside a slot
  uint16_t fragslot[KMAX][kBT];           // canonical slot of every fragment (kNoFrag: carries no gradient)
  uint16_t slotface[KMAX * kBT];          // slot -> face id within the mesh
  uint8_t cflag[KMAX * 8];                // per chunk: bit 0 head partial present, bit 1 tail partial, bit 2 whole chunk is one open run
  float tabx[kTileW], taby[kTileH];       // NDC sample positions of the tile's columns / rows
  int wsum[kBW];
  unsigned long long light[6];            // fixed-point light sums of the tile
  int next_chunk;
};

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

// a finished face: 18 floats into its (face, tile) record, the 6 light components into the tile's fixed-point sums
__device__ __forceinline__ void flush_face(const HfrShadeBwdTiledArgs& a, unsigned long long* light, long long rec, int lane,
                                           float tot, float fxs) {
  if (lane < HFR_FACE_REC_FLOATS) {
    if (rec < a.rec_cap) a.face_rec[rec * HFR_FACE_REC_FLOATS + lane] = tot;
    else if (lane == 0) atomicOr(a.status, 1u);
  } else if (lane < kNC && tot != 0.0f) {
    atomicAdd(&light[lane - HFR_FACE_REC_FLOATS], (unsigned long long)__float2ll_rn(tot * fxs));
  }
}

// fixed-point multiplier from the largest gradient magnitude: 2^34 over the power of two above it
__device__ __forceinline__ float fx_from_gmax(float g) {
  if (!(g > 1e-37f) || !(g < 1e37f)) g = 1.0f;
  const unsigned bits = __float_as_uint(g);
  int e = (int)((bits >> 23) & 255u) - 127 + ((bits & 0x7fffffu) ? 1 : 0);
  e = max(-90, min(90, 34 - e));
  return __uint_as_float((unsigned)(127 + e) << 23);
}

#ifndef HFR_TILED_MINB
#define HFR_TILED_MINB 2
#endif
#ifndef HFR_TILED_MINB8
#define HFR_TILED_MINB8 2   // K = 8: 2 CTAs / SM (128 registers, 102 KB shared memory each): C5 backward 1538 -> 1110 us (B=32)
#endif
template <int KMAX>
__global__ void __launch_bounds__(kBT, (KMAX <= 4 ? HFR_TILED_MINB : (KMAX <= 8 ? HFR_TILED_MINB8 : 1)))
shade_bwd_tiled_kernel(HfrShadeBwdTiledArgs a, WsLayout L, int FW) {
  extern __shared__ __align__(16) unsigned char smraw[];
  TSmem<KMAX>& sm = *reinterpret_cast<TSmem<KMAX>*>(smraw);
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(smraw + ((sizeof(TSmem<KMAX>) + 15) & ~(size_t)15));   // FW words: faces present in the tile
  uint32_t* wprefix = bitmap + FW;                                                                     // FW words: set bits before each word
  const HfrShadeFwdArgs& f = a.f;
  const HfrShadeParams& P = f.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = P.K, V = P.V;
  const uint32_t* __restrict__ ws = reinterpret_cast<const uint32_t*>(a.raster_ws);
  int n, tx, ty;
  if (a.tile_queue) {
    // the forward's tile queue: block i takes the i-th tile that holds faces, heaviest cost class first
    const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(a.tile_queue);
    uint32_t left = blockIdx.x;
    if (left >= __ldg(q)) return;
    const int TX = (P.W + kTileW - 1) / kTileW, TY = (P.H + kTileH - 1) / kTileH;
    const uint32_t T = (uint32_t)(P.N * TX * TY);
    int k = kCostClasses - 1;
    for (; k > 0; --k) {
      const uint32_t ck = __ldg(q + 2 + k);
      if (left < ck) break;
      left -= ck;
    }
    const uint32_t t = __ldg(q + kQueueHdr + T + (size_t)k * T + left);
    n = (int)(t / (uint32_t)(TX * TY));
    const int rem = (int)(t - (uint32_t)n * (uint32_t)(TX * TY));
    tx = rem % TX; ty = rem / TX;
  } else {
    n = blockIdx.z; tx = blockIdx.x; ty = blockIdx.y;
    // tile outside this mesh's footprint: no fragment, no record
    const uint4 bx = __ldg(reinterpret_cast<const uint4*>(ws + L.box) + n);
    if (tx < (int)bx.x || tx > 255 - (int)bx.y || ty < (int)bx.z || ty > 255 - (int)bx.w) return;
  }
  const int xi = tx * kTileW + (warp & 1) * 8 + (lane & 7);
  const int yi = ty * kTileH + (warp >> 1) * 4 + (lane >> 3);
  const bool active = xi < P.W && yi < P.H;
  const bool phong = P.shade == HFR_SHADE_PHONG_UV;
  const int kshade = phong ? (P.blend == HFR_BLEND_SOFTMAX ? K : 1) : 0;
  const size_t pix = ((size_t)n * P.H + yi) * P.W + xi;

  float k_mrgb = 0.0f;
  if (a.fix_sums) {
    const float icnt = 1.0f / (float)a.fix_count;
    k_mrgb = __ldg(a.fix_w + 1) * 2.0f * (__ldg(a.fix_sums + HFR_LOSS_SUM_T) - __ldg(a.fix_sums + HFR_LOSS_SUM_R)) * icnt * (-icnt);
  }
  for (int i = tid; i < FW; i += kBT) bitmap[i] = 0u;
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
      if (a.fix_sums) {
        // mean-RGB term (losses.py:369): d/d(rim) = k_mrgb for every pixel and channel with rim = rgb * alpha / scale,
        // k_mrgb from the (all-reduced) sums - added here so that the loss backward never waits for the collective
        const float4 im = __ldg(reinterpret_cast<const float4*>(a.fix_image + (a.pool_aa > 1
            ? (((size_t)n * (P.H / a.pool_aa) + yi / a.pool_aa) * (P.W / a.pool_aa) + xi / a.pool_aa) : pix) * 4));
        const float sc = a.pool_aa > 1 ? 1.0f / (float)(a.pool_aa * a.pool_aa) : 1.0f;
        const float so = im.w * a.fix_inv_scale;
        g4.x += k_mrgb * so * sc; g4.y += k_mrgb * so * sc; g4.z += k_mrgb * so * sc;
        if (!(a.pool_aa > 1 && a.pool_binarize)) g4.w += k_mrgb * (im.x + im.y + im.z) * a.fix_inv_scale * sc;
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
      // hidden slots of a hard blend carry no gradient
      if (!(((vmask >> k) & 1u) && (P.blend != HFR_BLEND_HARD || k == 0))) fl[k] = -1;
      if (fl[k] >= 0) atomicOr(&bitmap[fl[k] >> 5], 1u << (fl[k] & 31));
    }
  }
  __syncthreads();
  // canonical slot of a face = its rank among the faces present (bitmap order = face-id order): whatever order the
  // atomics ran in, the sorted fragment list below comes out the same - so do the 32-fragment chunks
  int D = 0;
  {
    const int per = (FW + kBT - 1) / kBT, w0 = tid * per, w1 = min(w0 + per, FW);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(bitmap[w]);
    int q = block_excl_scan(cnt, sm.wsum, &D);
    for (int w = w0; w < w1; ++w) {
      wprefix[w] = (uint32_t)q;
      uint32_t bits = bitmap[w];
      while (bits) {
        const int bpos = __ffs(bits) - 1;
        bits &= bits - 1;
        sm.slotface[q++] = (uint16_t)(w * 32 + bpos);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    uint16_t sl = kNoFrag;
    if (fl[k] >= 0) sl = (uint16_t)(wprefix[fl[k] >> 5] + __popc(bitmap[fl[k] >> 5] & ((1u << (fl[k] & 31)) - 1u)));
    sm.fragslot[k][tid] = sl;
  }

  // per-sample light constants
  float dhat[3] = {0.f, 0.f, 0.f}, dlen = 1.f, lcol[3] = {0.f, 0.f, 0.f};
  if (phong) {
    light_dir_hat(f, n, dhat, &dlen);
    lcol[0] = __ldg(f.light_color + 3 * n); lcol[1] = __ldg(f.light_color + 3 * n + 1); lcol[2] = __ldg(f.light_color + 3 * n + 2);
  }
  float fxs;
  if (a.gmax_bits) {       // every CTA derives the same multiplier; the first thread of each leaves it for hfr_grad_finish
    fxs = fx_from_gmax(__uint_as_float(__ldg(a.gmax_bits)) + 4.0f * fabsf(k_mrgb));
    if (tid == 0) *a.fx_scale = fxs;
  } else {
    fxs = *a.fx_scale;
  }
  const float fcx = __ldg(a.focal + 2 * n), fcy = __ldg(a.focal + 2 * n + 1);
  const size_t tbase = (P.tex_n == 1 ? 0 : (size_t)n * P.tex_h * P.tex_w * 3);
  const float zr = P.zfar - P.znear;

  for (int p0 = 0; p0 < D; p0 += kCap) {
    const int Dp = min(kCap, D - p0);
    // ============================================================== 2. counting sort of the fragments by face
    for (int i = tid; i < Dp * 8; i += kBT) (&sm.masks[0][0])[i] = 0u;
    if (tid == 0) sm.next_chunk = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int s = (int)sm.fragslot[k][tid] - p0;
      if ((unsigned)s < (unsigned)Dp) atomicOr(&sm.masks[s][warp], 1u << lane);
    }
    // stage the pass's faces: corner attributes and the record slot, one face per thread
    if (tid < Dp) {
      const int face = sm.slotface[p0 + tid];
      float* r = sm.facerec[tid];
      int vid[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) vid[c] = __ldg(f.faces + 3 * face + c);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* __restrict__ x = f.verts_view + ((size_t)n * V + vid[c]) * 3;
        const float* __restrict__ q = a.verts_ndc + ((size_t)n * V + vid[c]) * 3;
        r[3 * c] = __ldg(x); r[3 * c + 1] = __ldg(x + 1); r[3 * c + 2] = __ldg(x + 2);
        r[24 + 2 * c] = __ldg(q); r[25 + 2 * c] = __ldg(q + 1);
        if (phong) {
          const float* __restrict__ nn = f.vnormals + ((size_t)n * V + vid[c]) * 3;
          const int t = __ldg(f.faces_uvs + 3 * face + c);
          r[9 + 3 * c] = __ldg(nn); r[10 + 3 * c] = __ldg(nn + 1); r[11 + 3 * c] = __ldg(nn + 2);
          r[18 + 2 * c] = __ldg(f.verts_uvs + 2 * t); r[19 + 2 * c] = __ldg(f.verts_uvs + 2 * t + 1);
        }
      }
      r[30] = __uint_as_float(face_rec_index(ws, L, (int64_t)n * P.F + face, tx, ty));
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
      const int s = (int)sm.fragslot[k][tid] - p0;
      if ((unsigned)s < (unsigned)Dp) {
        int rank = __popc(sm.masks[s][warp] & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) rank += __popc(sm.masks[s][w]);
        sm.sorted[sm.offs[s] + rank] = (uint16_t)(tid | (k << 8));
      }
    }
    const int nch = (nfr + 31) >> 5;
    if (tid < nch) sm.cflag[tid] = 0;
    __syncthreads();

    // ============================================================== 3. differentiate 32 sorted fragments at a time
    // Chunks are canonical (the sorted list is), so any warp may take any chunk: they are handed out dynamically.
    // Inside a chunk lane c adds component c of each face's fragments one after the other; a face that lies wholly
    // inside the chunk is finished there, a face cut by a chunk boundary leaves partial sums that the fix-up below
    // joins in chunk order.
    while (true) {
      int c = 0;
      if (lane == 0) c = atomicAdd(&sm.next_chunk, 1);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= nch) break;
      const int i = c * 32 + lane;
      const bool valid = i < nfr;
      int slot = -2;
      float v[kNC];
#pragma unroll
      for (int q = 0; q < kNC; ++q) v[q] = 0.0f;
      if (valid) {
        const int e = sm.sorted[i], p = e & 255, k = e >> 8;
        slot = (int)sm.fragslot[k][p] - p0;
        const float4* __restrict__ r4 = reinterpret_cast<const float4*>(sm.facerec[slot]);
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
        {
          const float4 q0 = r4[0], q1 = r4[1];
          g.X[0] = q0.x; g.X[1] = q0.y; g.X[2] = q0.z; g.X[3] = q0.w; g.X[4] = q1.x; g.X[5] = q1.y; g.X[6] = q1.z; g.X[7] = q1.w;
          g.X[8] = sm.facerec[slot][8];
        }
        HfrTexTap tap; HfrPhongCtx ctx; float texel[3];
        const bool shaded = k < kshade;
        if (shaded) {
          const float4 q2 = r4[2], q3 = r4[3], q4 = r4[4], q5 = r4[5];
          g.Nv[0] = q2.y; g.Nv[1] = q2.z; g.Nv[2] = q2.w; g.Nv[3] = q3.x; g.Nv[4] = q3.y; g.Nv[5] = q3.z; g.Nv[6] = q3.w;
          g.Nv[7] = q4.x; g.Nv[8] = q4.y;
          g.uv[0] = q4.z; g.uv[1] = q4.w; g.uv[2] = q5.x; g.uv[3] = q5.y; g.uv[4] = q5.z; g.uv[5] = q5.w;
          shade_fragment<false>(f, n, g, bc, dhat, lcol, col, &tap, &ctx, texel);
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
            for (int q = 0; q < 3; ++q) { v[6 * c3 + q] = bc[c3] * gP[q]; v[6 * c3 + 3 + q] = bc[c3] * gNn[q]; }
          }
        }
        const float gd = P.blend == HFR_BLEND_HARD ? 0.f : HFR_FDIV(-gprob * pk * (1.0f - pk), P.sigma);
        // rasterizer backward on the face's NDC vertices (z_ndc = view z), then d(ndc)/d(view) per corner:
        //   x = fx X / Z + px,  y = fy Y / Z + py,  z = Z
        float vv[9], gvv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        {
          const float4 q6 = r4[6];
          const float2 q7 = *reinterpret_cast<const float2*>(sm.facerec[slot] + 28);
          vv[0] = q6.x; vv[1] = q6.y; vv[2] = g.X[2]; vv[3] = q6.z; vv[4] = q6.w; vv[5] = g.X[5]; vv[6] = q7.x; vv[7] = q7.y; vv[8] = g.X[8];
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
      // stage the components, then lane q walks the chunk's members of each face in order
#pragma unroll
      for (int q = 0; q < kNC; ++q) sm.stage[warp][q][lane] = v[q];
      __syncwarp();
      const int cnt = min(32, nfr - c * 32);
      // does the first face continue from the previous chunk / the last one into the next?
      int prev_slot = -1, next_slot = -1;
      if (c > 0) { const int e = sm.sorted[c * 32 - 1]; prev_slot = (int)sm.fragslot[e >> 8][e & 255] - p0; }
      if (c * 32 + 32 < nfr) { const int e = sm.sorted[c * 32 + 32]; next_slot = (int)sm.fragslot[e >> 8][e & 255] - p0; }
      const int up = __shfl_up_sync(0xffffffffu, slot, 1);
      const bool starts = valid && (lane == 0 || slot != up);
      const unsigned sb = __ballot_sync(0xffffffffu, starts);
      const int first_slot = __shfl_sync(0xffffffffu, slot, 0), last_slot = __shfl_sync(0xffffffffu, slot, cnt - 1);
      const bool head_cont = first_slot == prev_slot, tail_cont = last_slot == next_slot;
      int m = 0;
      unsigned flags = 0;
      while (m < cnt) {
        const unsigned nxt = sb & ~((2u << m) - 1u);
        const int m1 = nxt ? __ffs(nxt) - 1 : cnt;
        float tot = 0.0f;
        if (lane < kNC)
          for (int j = m; j < m1; ++j) tot += sm.stage[warp][lane][j];
        const bool is_head = (m == 0) && head_cont, is_tail = (m1 == cnt) && tail_cont;
        if (is_head) {                         // (a run that is head AND tail fills the whole chunk: flag bit 2)
          if (lane < kNC) sm.part[c][0][lane] = tot;
          flags |= 1u | (is_tail ? 4u : 0u);
        } else if (is_tail) {
          if (lane < kNC) sm.part[c][1][lane] = tot;
          flags |= 2u;
        } else {
          const int sl = __shfl_sync(0xffffffffu, slot, m);
          flush_face(a, sm.light, (long long)__float_as_uint(sm.facerec[sl][30]), lane, tot, fxs);
        }
        m = m1;
      }
      if (lane == 0) sm.cflag[c] = (uint8_t)flags;
      __syncwarp();
    }
    __syncthreads();
    // fix-up: a face cut by chunk boundaries = tail partial of its first chunk + the whole-chunk partials in between +
    // head partial of its last chunk, added in chunk order (canonical, so the rounding is too)
    for (int t = tid; t < nch * 32; t += kBT) {
      const int c = t >> 5, q = t & 31;
      if (q < kNC && (sm.cflag[c] & 2u)) {
        float tot = sm.part[c][1][q];
        int cc = c + 1;
        while (sm.cflag[cc] & 4u) { tot += sm.part[cc][0][q]; ++cc; }
        tot += sm.part[cc][0][q];
        const int e = sm.sorted[min(c * 32 + 31, nfr - 1)];
        const int sl = (int)sm.fragslot[e >> 8][e & 255] - p0;
        flush_face(a, sm.light, (long long)__float_as_uint(sm.facerec[sl][30]), q, tot, fxs);
      }
    }
    __syncthreads();     // the next pass reuses masks / sorted / facerec
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
  const float inv = 1.0f / *a.fx_scale;     // a power of two: exact
  if (a.gmax_bits && blockIdx.x == 0 && threadIdx.x == 0) *a.gmax_bits = 0u;
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
  HFR_CHECK_ARG(!a->fix_sums || (a->fix_w && a->fix_image && a->fix_count > 0 && a->fix_inv_scale > 0.f),
                "shade_backward_tiled: the mean-RGB fix-up needs sums, weights, the loss image and its element count");
  HFR_CHECK_ARG(a->pool_aa <= 1 || (a->pool_aa <= 16 && p.H % a->pool_aa == 0 && p.W % a->pool_aa == 0),
                "shade_backward_tiled: image size must be a multiple of pool_aa (<= 16)");
  HFR_CHECK_ARG((p.W + kTileW - 1) / kTileW <= 255 && (p.H + kTileH - 1) / kTileH <= 255, "shade_backward_tiled: image too large");
  HFR_CHECK_ARG((reinterpret_cast<uintptr_t>(a->face_rec) & 15) == 0, "shade_backward_tiled: face_rec must be 16-byte aligned");
  HFR_CHECK_ARG(p.F <= kMaxFacesTiled, "shade_backward_tiled: at most %d faces per mesh", kMaxFacesTiled);
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
  if (a->tile_queue) grid = dim3(grid.x * grid.y * grid.z, 1, 1);
  const int FW = (p.F + 31) / 32;
#define HFR_LAUNCH_T(KM)                                                                                         \
  do {                                                                                                           \
    const size_t smem = ((sizeof(TSmem<KM>) + 15) & ~(size_t)15) + (size_t)FW * 8;                               \
    HFR_CHECK_ARG(smem <= 227 * 1024, "shade_backward_tiled: mesh too large for shared memory (%zu B)", smem);   \
    cudaFuncSetAttribute(shade_bwd_tiled_kernel<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    shade_bwd_tiled_kernel<KM><<<grid, kBT, smem, st>>>(*a, L, FW);                                              \
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

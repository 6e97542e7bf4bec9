// Tile-binned rasterizer core (device only): coarse face culling per 16x16 pixel tile with warp
// ballots, face records staged in shared memory, fine per-pixel coverage test with top-K depth
// insertion in registers.  The epilogue (what to do with the K winners of a pixel) is a functor,
// so the same core serves the standalone rasterizer (writes Fragments) and the fused
// rasterize+shade kernel.
//
// Work decomposition
//   grid  = (tiles_x, tiles_y, N), block = 256 threads = one 16x16 tile, one pixel per thread;
//           a warp owns an 8x4 pixel block, so every Fragments row segment a warp writes is
//           contiguous (8 pixels * K * {8,4,12,4} B) and sector-aligned.
//   coarse: every thread looks at one face's packed tile range (4 B, coalesced, from the setup
//           kernel) per step; hits are compacted with a popc prefix into the tile list (any order:
//           the list is depth-sorted afterwards and the top-K resolves z-ties by the packed face
//           index explicitly - the CPU reference's (z, face) ordering).
//   stage:  the listed faces' 9 floats are gathered once per tile into 64-byte shared records
//           together with the dilated bbox and the barycentric denominator.
//   fine:   per 32 list entries, lane i builds the 32-bit mask of the warp's 8x4 pixels that lie in
//           face i's dilated bbox; a 5-step shuffle transpose turns the 32 masks into one mask per
//           PIXEL of the faces it has to look at.  Every lane then walks ITS OWN faces (front to
//           back), so the expensive exact coverage / depth / distance math runs on dense warps
//           instead of idling the lanes a face does not touch.
#pragma once
#include "raster_math.cuh"

namespace hfr {

constexpr int kTileW = 16, kTileH = 16, kRasterThreads = 256;
constexpr int kListCap = 256;   // faces per staged batch
constexpr int kRecFloats = 20;  // 80-byte record (9 vertex floats, dilated bbox, area, zmin, face id, pad): the 20-word
                                // stride spreads the per-lane LDS.128 of the fine pass over all bank groups
constexpr int kDepthBuckets = 32;   // one per lane: every warp scans the histogram in registers
constexpr int kChunk = 32 * kRasterThreads;  // faces handled per coarse pass (one hit bit per face per thread)
constexpr int kCostClasses = 8;     // tile queue (raster.cu): cost classes, header words, empty tiles per fill group
constexpr int kQueueHdr = 16;
constexpr int kFillRun = 8;

struct RasterSmem {
  __align__(16) float rec[kListCap * kRecFloats];
  uint32_t wmask[kRasterThreads / 32][kListCap / 32][32];   // per warp, per 32 list entries: one face mask per lane (pixel)
  float tabx[kTileW], taby[kTileH];                          // NDC sample positions of the tile's columns / rows
  int wsum[kRasterThreads / 32];
  uint16_t order[kListCap];   // record indices ordered front to back (depth buckets of the faces' nearest vertex)
  int fidx[kListCap];         // face index (within the mesh) of every list entry of the batch being staged
  float zred[2][kRasterThreads / 32];   // per-warp min / max of the batch's nearest-vertex depths
  int hist[kDepthBuckets];              // faces per depth bucket
  int bmin[kDepthBuckets];              // smallest nearest-vertex depth in the bucket (float bits; depths are > 0)
};

// 32x32 bit-matrix transpose across the warp: on return bit i of lane L = bit L of lane i's input
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const uint32_t m = d == 16 ? 0x0000ffffu : d == 8 ? 0x00ff00ffu : d == 4 ? 0x0f0f0f0fu : d == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, d);
    x = (lane & d) ? ((x & ~m) | ((y >> d) & m)) : ((x & m) | ((y << d) & ~m));
  }
  return x;
}

__device__ __forceinline__ uint32_t pack_tile_range(int txmin, int txmax, int tymin, int tymax) {
  return (uint32_t)txmin | ((uint32_t)txmax << 8) | ((uint32_t)tymin << 16) | ((uint32_t)tymax << 24);
}
constexpr uint32_t kEmptyRange = 0x00ff00ffu;  // txmin=255 > txmax=0

template <int KMAX>
struct TopK {
  float z[KMAX];
  int f[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KMAX; ++i) { z[i] = INFINITY; f[i] = -1; }
  }
  __device__ __forceinline__ float worst() const { return z[KMAX - 1]; }
  // does (pz, face) beat the current K-th entry in the reference's (z, face index) order?
  __device__ __forceinline__ bool beats_worst(float pz, int face) const {
    return pz < z[KMAX - 1] || (pz == z[KMAX - 1] && face < f[KMAX - 1]);
  }
  // replace the worst slot and bubble towards the front, ordered by (z, face index): faces may arrive in
  // any order (the tile list is depth-sorted), ties still resolve to the smaller packed face index
  // as insert(), for the payload cache: `perm` holds one nibble per sorted position = the physical payload slot of
  // that entry.  The new entry takes over the slot of the entry it evicts (returned), and the nibbles follow the
  // bubble, so payloads are written once and never moved.
  __device__ __forceinline__ int insert_slot(float pz, int face, uint32_t& perm) {
    const int phys = (int)((perm >> ((4 * (KMAX - 1)) & 31)) & 0xfu);   // (payload slots exist for KMAX <= 8 only)
    z[KMAX - 1] = pz; f[KMAX - 1] = face;
#pragma unroll
    for (int i = KMAX - 1; i > 0; --i) {
      if (z[i] < z[i - 1] || (z[i] == z[i - 1] && f[i] < f[i - 1])) {
        const float tz = z[i]; z[i] = z[i - 1]; z[i - 1] = tz;
        const int tf = f[i]; f[i] = f[i - 1]; f[i - 1] = tf;
        const uint32_t x = ((perm >> ((4 * i) & 31)) ^ (perm >> ((4 * (i - 1)) & 31))) & 0xfu;
        perm ^= (x << ((4 * i) & 31)) | (x << ((4 * (i - 1)) & 31));
      }
    }
    return phys;
  }
  __device__ __forceinline__ void insert(float pz, int face) {
    z[KMAX - 1] = pz; f[KMAX - 1] = face;
#pragma unroll
    for (int i = KMAX - 1; i > 0; --i) {
      if (z[i] < z[i - 1] || (z[i] == z[i - 1] && f[i] < f[i - 1])) {
        const float tz = z[i]; z[i] = z[i - 1]; z[i - 1] = tz;
        const int tf = f[i]; f[i] = f[i - 1]; f[i - 1] = tf;
      }
    }
  }
};

struct PixelCtx {
  int n, tx, ty, xi, yi;
  float xf, yf;          // NDC sample position; filled by raster_tile (tiles outside the mesh never need it)
  bool pix_active, warp_active;
};

__device__ __forceinline__ PixelCtx make_pixel_ctx(int H, int W) {
  PixelCtx c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  c.n = blockIdx.z; c.tx = blockIdx.x; c.ty = blockIdx.y;
  const int wx0 = c.tx * kTileW + (warp & 1) * 8, wy0 = c.ty * kTileH + (warp >> 1) * 4;
  c.xi = wx0 + (lane & 7);
  c.yi = wy0 + (lane >> 3);
  c.pix_active = c.xi < W && c.yi < H;
  c.warp_active = wx0 < W && wy0 < H;
  c.xf = 0.0f; c.yf = 0.0f;
  return c;
}

// the same for an explicit tile (tile-queue order: the CTA index no longer names the tile)
__device__ __forceinline__ PixelCtx make_pixel_ctx_at(int n, int tx, int ty, int H, int W) {
  PixelCtx c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  c.n = n; c.tx = tx; c.ty = ty;
  const int wx0 = tx * kTileW + (warp & 1) * 8, wy0 = ty * kTileH + (warp >> 1) * 4;
  c.xi = wx0 + (lane & 7);
  c.yi = wy0 + (lane >> 3);
  c.pix_active = c.xi < W && c.yi < H;
  c.warp_active = wx0 < W && wy0 < H;
  c.xf = 0.0f; c.yf = 0.0f;
  return c;
}

// hfr_pix_to_ndc with the per-axis range / offset hoisted (same X* sequence, same bits)
__device__ __forceinline__ float pix_to_ndc_pre(int i, float range, float offset, int S1) {
  return XADD(-offset, XDIV(XADD(XMUL(range, (float)i), offset), (float)S1));
}

// Coarse + stage + fine for one tile.  On return `top` holds, per thread (= pixel), the KMAX
// nearest valid faces as packed face ids (sorted by (z, id)).
//
//   coarse: thread t tests faces t, t + 256, ... (coalesced loads) and keeps one hit bit per face from a
//           single pass over the packed tile ranges; one block-wide scan of the hit counts then
//           gives every thread its slots in the tile list.  A tile no face touches leaves after
//           that scan.
//   stage:  listed faces are gathered once per tile into 80-byte shared records and ordered front
//           to back by their nearest vertex (counting sort over depth buckets in shared memory).
//   fine:   per 32 list entries lane i computes which of the warp's 8x4 pixels lie in face i's dilated
//           bbox; the 32 masks are transposed (5 shuffles) into one face mask per pixel.  Each lane then
//           walks its own faces front to back (per-lane LDS.128 of the record), skips faces whose nearest
//           vertex is not in front of its current K-th depth and stops once a whole depth bucket is, so the exact coverage / depth /
//           distance math runs on densely populated warps.  The top-K is ordered by (z, packed face index),
//           the CPU reference's order, ties included.
//
// PAY: the fused kernels keep, per thread and per top-K slot, the (barycentrics, signed distance) of the entry in
// shared memory (`pay`, [slot][thread] float4; `perm` maps sorted position -> slot), written when a candidate enters
// the top K.  The epilogue then reads the winners' Fragments values instead of recomputing the exact math.
constexpr uint32_t kPermIdentity = 0x76543210u;
template <int KMAX, bool PAY>
__device__ __forceinline__ void raster_tile(const HfrRasterArgs& a, const uint32_t* __restrict__ tile_ranges,
                                            const uint32_t* __restrict__ mesh_box, RasterSmem& sm, PixelCtx& c, TopK<KMAX>& top,
                                            float4* __restrict__ pay, uint32_t& perm) {
  perm = kPermIdentity;
  const int n = c.n, tx = c.tx, ty = c.ty;
  const bool pix_active = c.pix_active, warp_active = c.warp_active;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f0 = a.mesh_first[n];
  const int nf = (int)a.mesh_nfaces[n];
  const float blur = a.blur_radius, rblur = sqrtf(a.blur_radius);
  const int pc = a.perspective_correct, clip = a.clip_barycentric;
  // pz is a convex combination of the vertex depths unless an outside pixel keeps unclipped
  // barycentrics (blur > 0 without clipping): only then the zmin early-out is not exact
  const bool zcull = clip || !(blur > 0.0f);
  const bool hard = KMAX == 1 && !(blur > 0.0f);   // see the walk below (compile-time false for K > 1: those kernels are unchanged)
  top.init();
  if (mesh_box) {   // tile outside the mesh's footprint (union of its faces' tile ranges): nothing to do
    const uint4 bx = __ldg(reinterpret_cast<const uint4*>(mesh_box) + n);
    if (tx < (int)bx.x || tx > 255 - (int)bx.y || ty < (int)bx.z || ty > 255 - (int)bx.w) return;
  }
  // NDC sample positions of this tile's columns / rows (the same X* sequence as the oracle's pixel centres)
  if (tid < kTileW + kTileH) {
    const int H = a.H, W = a.W;
    if (tid < kTileW) {
      const float rx = W > H ? XDIV(XMUL(2.0f, (float)W), (float)H) : 2.0f, ox = XDIV(rx, 2.0f);
      sm.tabx[tid] = pix_to_ndc_pre(W - 1 - (tx * kTileW + tid), rx, ox, W);
    } else {
      const float ry = H > W ? XDIV(XMUL(2.0f, (float)H), (float)W) : 2.0f, oy = XDIV(ry, 2.0f);
      sm.taby[tid - kTileW] = pix_to_ndc_pre(H - 1 - (ty * kTileH + tid - kTileW), ry, oy, H);
    }
  }
  __syncthreads();
  const int lx0 = (warp & 1) * 8, ly0 = (warp >> 1) * 4;
  const float xf = sm.tabx[lx0 + (lane & 7)], yf = sm.taby[ly0 + (lane >> 3)];
  c.xf = xf; c.yf = yf;
  const uint32_t active_mask = __ballot_sync(0xffffffffu, pix_active);

  for (int cbase = 0; cbase < nf; cbase += kChunk) {
    const int cn = min(nf - cbase, kChunk);
    const int per = (cn + kRasterThreads - 1) / kRasterThreads;   // <= 32
    // thread t looks at faces cbase + j * 256 + t: every load of the warp is one coalesced 128-byte line and the
    // iterations are independent (4 in flight).  Hit bit j <-> that face; the order of the tile list is irrelevant
    // (it is depth-sorted below and z-ties are resolved by the packed face index explicitly).
    uint32_t hits = 0;
    for (int j0 = 0; j0 < per; j0 += 4) {
      uint32_t w4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int fi = (j0 + u) * kRasterThreads + tid;
        w4[u] = (j0 + u < per && fi < cn) ? __ldg(tile_ranges + f0 + cbase + fi) : kEmptyRange;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w = w4[u];
        const int txmin = w & 255, txmax = (w >> 8) & 255, tymin = (w >> 16) & 255, tymax = w >> 24;
        if (tx >= txmin && tx <= txmax && ty >= tymin && ty <= tymax) hits |= 1u << (j0 + u);
      }
    }
    // block-wide exclusive scan of the per-thread hit counts
    const int cnt = __popc(hits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();   // previous chunk's records consumed, wsum free
    if (lane == 31) sm.wsum[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kRasterThreads / 32; ++w) {
      const int c = sm.wsum[w];
      before += (w < warp) ? c : 0;
      total += c;
    }
    if (total == 0) continue;
    const int my0 = before + incl - cnt;   // list position of this thread's first hit

    for (int lbase = 0; lbase < total; lbase += kListCap) {
      const int bcnt = min(total - lbase, kListCap);
      if (lbase > 0) __syncthreads();   // previous batch consumed
      // stage, balanced: the hit face ids of this batch go to a shared index list first (a thread may own many
      // hits, a coherent mesh puts up to 32 consecutive faces of one thread into the same tile), then thread p
      // gathers and precomputes record p - one record per thread, 9 independent loads each
      {
        uint32_t m = hits;
        int pos = my0;
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          if (pos >= lbase && pos < lbase + bcnt) sm.fidx[pos - lbase] = cbase + j * kRasterThreads + tid;
          ++pos;
        }
      }
      __syncthreads();
      if (tid < bcnt) {
        const int face = (int)(f0 + sm.fidx[tid]);
        const float* __restrict__ v = a.face_verts + (size_t)face * 9;
        float r[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) r[e] = __ldg(v + e);
        float4* dst = reinterpret_cast<float4*>(sm.rec + tid * kRecFloats);
        const float xmin = XSUB(hfr_min3(r[0], r[3], r[6]), rblur), xmax = XADD(hfr_max3(r[0], r[3], r[6]), rblur);
        const float ymin = XSUB(hfr_min3(r[1], r[4], r[7]), rblur), ymax = XADD(hfr_max3(r[1], r[4], r[7]), rblur);
        const float area = XADD(hfr_edge(r[6], r[7], r[0], r[1], r[3], r[4]), HFR_KEPS);
        dst[0] = make_float4(r[0], r[1], r[2], r[3]);
        dst[1] = make_float4(r[4], r[5], r[6], r[7]);
        dst[2] = make_float4(r[8], xmin, xmax, ymin);
        // zmin shrunk by 1e-5: the rounded pz of a convex combination can undershoot the smallest z by a few ulp
        // (slot 12 = the walk's exit depth, filled by the depth-bucket pass below)
        dst[3] = make_float4(0.0f, area, hfr_min3(r[2], r[5], r[8]) * 0.99999f, __int_as_float(face));
        sm.rec[tid * kRecFloats + 16] = ymax;
      }
      __syncthreads();
      // depth order: the records are bucketed by their (shrunk) nearest-vertex depth into kDepthBuckets equal
      // slices of the batch's depth range (a counting sort: O(1) per face instead of a rank by pairwise compares).
      // Faces nearest to the camera are visited first, so a pixel's K-th depth drops quickly and most faces behind
      // it fail the zmin test.  Inside a bucket the order is arbitrary, therefore the walk may only STOP at a face
      // when the smallest depth of that face's bucket (slot 12 of the record) is not in front of the K-th depth:
      // later faces sit in the same or a farther bucket (the bucket index is monotone in the depth).
      if (zcull) {
        const float zi = tid < bcnt ? sm.rec[tid * kRecFloats + 14] : INFINITY;
        float lo = zi, hi = tid < bcnt ? zi : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
          hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { sm.zred[0][warp] = lo; sm.zred[1][warp] = hi; }
        if (tid < kDepthBuckets) { sm.hist[tid] = 0; sm.bmin[tid] = 0x7f800000; }
        __syncthreads();
        float zlo = sm.zred[0][0], zhi = sm.zred[1][0];
#pragma unroll
        for (int w = 1; w < kRasterThreads / 32; ++w) { zlo = fminf(zlo, sm.zred[0][w]); zhi = fmaxf(zhi, sm.zred[1][w]); }
        const float scale = zhi > zlo ? (float)kDepthBuckets / (zhi - zlo) : 0.0f;
        int b = 0, pos = 0;
        if (tid < bcnt) {
          b = min(kDepthBuckets - 1, (int)((zi - zlo) * scale));
          pos = atomicAdd(&sm.hist[b], 1);
          atomicMin(&sm.bmin[b], __float_as_int(zi));
        }
        __syncthreads();
        int incl = sm.hist[lane];   // every warp scans the 32 bucket counts in registers
        const int own = incl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int base = __shfl_sync(0xffffffffu, incl - own, b);
        if (tid < bcnt) {
          sm.order[base + pos] = (uint16_t)tid;
          sm.rec[tid * kRecFloats + 12] = __int_as_float(sm.bmin[b]);
        }
      } else if (tid < bcnt) {
        sm.order[tid] = (uint16_t)tid;
      }
      __syncthreads();
      if (warp_active) {
        const int nch = (bcnt + 31) >> 5;
        // A. which faces does each pixel of this warp have to look at?
        for (int ch = 0; ch < nch; ++ch) {
          const int i = ch * 32 + lane;
          uint32_t M = 0;
          if (i < bcnt) {
            const float* rp = sm.rec + (int)sm.order[i] * kRecFloats;
            const float4 q2 = *reinterpret_cast<const float4*>(rp + 8);   // (z2, xmin, xmax, ymin)
            const float ymax = rp[16];
            uint32_t xm = 0, ym = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float x = sm.tabx[lx0 + e]; xm |= (x < q2.y || x > q2.z) ? 0u : (1u << e); }
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float y = sm.taby[ly0 + e]; ym |= (y < q2.w || y > ymax) ? 0u : (1u << (8 * e)); }
            M = (xm * ym) & active_mask;   // lane = 8 * row + column; the four row copies of xm cannot carry into each other
          }
          sm.wmask[warp][ch][lane] = warp_transpose32(M, lane);
        }
        __syncwarp();
        // B. every lane walks its own faces, front to back; a lane is finished at the end of its list or at the
        //    first face whose nearest vertex is not in front of its K-th depth (the list is depth-ordered)
        int ch = 0;
        uint32_t Wc = sm.wmask[warp][0][lane];
        bool done = false;
        if (hard) {
          // K = 1, blur_radius == 0 (HardPhong settings): only a pixel INSIDE the face can win, and a pixel is outside as
          // soon as one edge function is zero or disagrees in sign with the face area (the barycentric quotient, and
          // its perspective-corrected form with z > 0, is then <= 0 - exact for IEEE division, underflow included).
          // Every lane therefore skips through its candidates with that sign test (21 flops) until one survives; the
          // exact coverage / depth math then runs once per SURVIVOR on lanes that all carry one, and the
          // point-triangle distance is only evaluated where it is an output.
          while (true) {
            int ri = -1;
            float4 q3 = make_float4(0.f, 0.f, 0.f, 0.f);
            while (!done) {
              while (Wc == 0 && ++ch < nch) Wc = sm.wmask[warp][ch][lane];
              if (Wc == 0) { done = true; break; }
              const int j = __ffs(Wc) - 1;
              Wc &= Wc - 1;
              const int r = sm.order[ch * 32 + j];
              q3 = *reinterpret_cast<const float4*>(sm.rec + r * kRecFloats + 12);   // (bucket zmin, area, zmin, face)
              if (zcull && !(q3.z < top.worst())) {
                if (!(q3.x < top.worst())) done = true;
                continue;
              }
              const float4* r4 = reinterpret_cast<const float4*>(sm.rec + r * kRecFloats);
              const float4 q0 = r4[0], q1 = r4[1];   // x0 y0 z0 x1 | y1 z1 x2 y2
              if (hfr_edge_sign_outside(xf, yf, q0.x, q0.y, q0.w, q1.x, q1.z, q1.w, q3.y)) continue;
              ri = r;
              break;
            }
            if (!__any_sync(0xffffffffu, ri >= 0)) break;   // every lane left the search finished
            if (ri >= 0) {
              const float4* r4 = reinterpret_cast<const float4*>(sm.rec + ri * kRecFloats);
              const float4 q0 = r4[0], q1 = r4[1];
              const float v[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, sm.rec[ri * kRecFloats + 8]};
              const int face = __float_as_int(q3.w);
              float pz, bc[3];
              bool inside;
              if (hfr_raster_bary(xf, yf, v, q3.y, pc, clip, &pz, bc, &inside) && inside && top.beats_worst(pz, face)) {
                if (PAY) {
                  const float dd = hfr_tri_dist2(xf, yf, v);
                  const int slot = top.insert_slot(pz, face, perm);
                  pay[slot * kRasterThreads] = make_float4(bc[0], bc[1], bc[2], -dd);
                } else {
                  top.insert(pz, face);
                }
              }
            }
          }
        } else {
          while (true) {
            int ri = -1;
            if (!done) {
              while (Wc == 0 && ++ch < nch) Wc = sm.wmask[warp][ch][lane];
              if (Wc == 0) {
                done = true;
              } else {
                const int j = __ffs(Wc) - 1;
                Wc &= Wc - 1;
                ri = sm.order[ch * 32 + j];
              }
            }
            float4 q3 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ri >= 0) {
              q3 = *reinterpret_cast<const float4*>(sm.rec + ri * kRecFloats + 12);   // (bucket zmin, area, zmin, face)
              if (zcull && !(q3.z < top.worst())) {   // this face cannot enter the top K ...
                if (!(q3.x < top.worst())) done = true;   // ... and neither can any later one
                ri = -1;
              }
            }
            if (!__any_sync(0xffffffffu, ri >= 0)) {
              if (__all_sync(0xffffffffu, done)) break;
              continue;
            }
            if (ri >= 0) {
              const float4* r4 = reinterpret_cast<const float4*>(sm.rec + ri * kRecFloats);
              const float4 q0 = r4[0], q1 = r4[1];
              const float v[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, sm.rec[ri * kRecFloats + 8]};
              const int face = __float_as_int(q3.w);
              float pz, bc[3];
              bool inside;
              if (hfr_raster_bary(xf, yf, v, q3.y, pc, clip, &pz, bc, &inside)) {
                if (top.beats_worst(pz, face)) {
                  const float dd = (PAY || !inside) ? hfr_tri_dist2(xf, yf, v) : 0.0f;
                  if (inside || dd < blur) {
                    if (PAY) {
                      const int slot = top.insert_slot(pz, face, perm);
                      pay[slot * kRasterThreads] = make_float4(bc[0], bc[1], bc[2], inside ? -dd : dd);
                    } else {
                      top.insert(pz, face);
                    }
                  }
                }
              }
            }
          }
        }
      }
    }
  }
}


// Recompute the winners' barycentrics / depth / distance from the packed face floats (the winner
// already passed the validity, bbox and blur tests in the fine pass; the formulas are the same
// X* sequences, so z is reproduced bit for bit).
template <int KMAX>
__device__ __forceinline__ void compute_fragments(const HfrRasterArgs& a, float xf, float yf, const TopK<KMAX>& top,
                                                  int64_t* id, float* z, float* d, float* b) {
  const int K = a.K;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    id[k] = -1; z[k] = -1.0f; d[k] = -1.0f; b[3 * k] = b[3 * k + 1] = b[3 * k + 2] = -1.0f;
    if (k < K && top.f[k] >= 0) {
      float v[9];
      const float* __restrict__ src = a.face_verts + (size_t)top.f[k] * 9;
#pragma unroll
      for (int e = 0; e < 9; ++e) v[e] = __ldg(src + e);
      const float area = XADD(hfr_edge(v[6], v[7], v[0], v[1], v[3], v[4]), HFR_KEPS);
      float pz, bc[3];
      bool inside;
      hfr_raster_bary(xf, yf, v, area, a.perspective_correct, a.clip_barycentric, &pz, bc, &inside);
      const float dd = hfr_tri_dist2(xf, yf, v);
      id[k] = top.f[k]; z[k] = pz; d[k] = inside ? -dd : dd;
      b[3 * k] = bc[0]; b[3 * k + 1] = bc[1]; b[3 * k + 2] = bc[2];
    }
  }
}

// Stream the four Fragments tensors of one pixel out (128-bit evict-first stores when K allows).
template <int KMAX>
__device__ __forceinline__ void store_fragments(const HfrRasterArgs& a, size_t pix, const int64_t* id, const float* z,
                                                const float* d, const float* b) {
  const int K = a.K;
  int64_t* p2f = a.pix_to_face + pix * K;
  float* zb = a.zbuf + pix * K;
  float* ds = a.dists + pix * K;
  float* ba = a.bary + pix * K * 3;
  if (K == KMAX && (KMAX % 4) == 0) {
#pragma unroll
    for (int k = 0; k < KMAX; k += 2) st_cs_i64x2(p2f + k, id[k], id[k + 1]);
#pragma unroll
    for (int k = 0; k < KMAX; k += 4) {
      st_cs_f4(zb + k, z[k], z[k + 1], z[k + 2], z[k + 3]);
      st_cs_f4(ds + k, d[k], d[k + 1], d[k + 2], d[k + 3]);
    }
#pragma unroll
    for (int e = 0; e < KMAX * 3; e += 4) st_cs_f4(ba + e, b[e], b[e + 1], b[e + 2], b[e + 3]);
  } else {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        p2f[k] = id[k]; zb[k] = z[k]; ds[k] = d[k];
        ba[3 * k] = b[3 * k]; ba[3 * k + 1] = b[3 * k + 1]; ba[3 * k + 2] = b[3 * k + 2];
      }
    }
  }
}

// Is tile (tx, ty) of mesh n outside the mesh's footprint (union of its faces' tile ranges)?  CTA-uniform.
__device__ __forceinline__ bool tile_outside_mesh(const uint32_t* __restrict__ mesh_box, int n, int tx, int ty) {
  if (!mesh_box) return false;
  const uint4 bx = __ldg(reinterpret_cast<const uint4*>(mesh_box) + n);
  return tx < (int)bx.x || tx > 255 - (int)bx.y || ty < (int)bx.z || ty > 255 - (int)bx.w;
}

// -1 fill of the four Fragments tensors over one whole tile that no face touches: the tile's 16 row segments are
// contiguous runs (16 K x {8, 4, 4, 12} B), so the CTA streams them as 448 K 128-bit evict-first stores instead of
// 6-13 narrow stores per pixel.  Returns false (nothing written) when the tile is clipped by the image border or the
// rows are not 16-byte aligned; the per-pixel epilogue then does the fill.
#ifndef HFR_FILL_PLAIN
#define HFR_FILL_PLAIN 0
#endif
__device__ __forceinline__ bool fill_empty_tile(const HfrRasterArgs& a, int n, int tx, int ty) {
  const int K = a.K, W = a.W, H = a.H;
  if ((tx + 1) * kTileW > W || (ty + 1) * kTileH > H || ((W * K) & 3)) return false;
  const int per_row = 28 * K;                         // 128-bit units per tile row: 8K ids, 4K z, 4K dists, 12K bary
  const int row = threadIdx.x >> 4, t16 = threadIdx.x & 15;   // 16 threads stream one tile row
  const size_t pix = ((size_t)n * H + (size_t)ty * kTileH + row) * W + (size_t)tx * kTileW;
  float* const p0 = reinterpret_cast<float*>(a.pix_to_face + pix * K);
  float* const p1 = a.zbuf + pix * K;
  float* const p2 = a.dists + pix * K;
  float* const p3 = a.bary + pix * K * 3;
  const float m1 = -1.0f, mi = __int_as_float(-1);
  if (K == 1) {
    // 28 units per row, fixed roles: lanes 0-7 the ids, 8-11 z, 12-15 dists, then lanes 0-11 the barycentrics
    float* const d1 = t16 < 8 ? p0 + 4 * t16 : (t16 < 12 ? p1 + 4 * (t16 - 8) : p2 + 4 * (t16 - 12));
    const float v1 = t16 < 8 ? mi : m1;
    st_cs_f4(d1, v1, v1, v1, v1);
    if (t16 < 12) st_cs_f4(p3 + 4 * t16, m1, m1, m1, m1);
    return true;
  }
  if ((K & 3) == 0) {
    // K = 4, 8, 16: every group of 16 lanes stays inside one tensor (K/2 rounds of ids, K/4 of z, K/4 of dists,
    // 3K/4 of barycentrics per row) - no per-unit address selection
    for (int i = 0; i < K / 2; ++i) st_cs_f4(p0 + 4 * (t16 + 16 * i), mi, mi, mi, mi);
    for (int i = 0; i < K / 4; ++i) {
      st_cs_f4(p1 + 4 * (t16 + 16 * i), m1, m1, m1, m1);
      st_cs_f4(p2 + 4 * (t16 + 16 * i), m1, m1, m1, m1);
    }
    for (int i = 0; i < 3 * K / 4; ++i) st_cs_f4(p3 + 4 * (t16 + 16 * i), m1, m1, m1, m1);
    return true;
  }
  for (int u = t16; u < per_row; u += 16) {
    float* dst = u < 8 * K ? p0 + 4 * u : (u < 12 * K ? p1 + 4 * (u - 8 * K) : (u < 16 * K ? p2 + 4 * (u - 12 * K) : p3 + 4 * (u - 16 * K)));
    const float v = u < 8 * K ? mi : m1;   // two int64 -1 per 128 bits of pix_to_face
#if HFR_FILL_PLAIN
    *reinterpret_cast<float4*>(dst) = make_float4(v, v, v, v);
#else
    st_cs_f4(dst, v, v, v, v);
#endif
  }
  return true;
}

// The same fill with the TMA engine: the tile's 16 row segments of the four Fragments tensors and of the RGBA image are
// bulk copies (cp.async.bulk shared -> global) out of constant patterns in shared memory - 80 copy instructions per
// tile instead of 256 threads x 14 vector stores, issued with an evict-first L2 policy (the data is read once, by the
// backward).  A group of `len` horizontally adjacent empty tiles is one set of copies (len x longer rows).  `pat` needs
// ((2 + 3) * 16 K + 64) * len words.  Returns false when the run is clipped by the image border or rows are unaligned.
__device__ __forceinline__ bool fill_empty_tile_bulk(const HfrRasterArgs& a, float* __restrict__ image, const float* bg, int n,
                                                     int tx, int ty, int len, uint32_t* pat) {
  const int K = a.K, W = a.W, H = a.H;
  if ((tx + len) * kTileW > W || (ty + 1) * kTileH > H || ((W * K) & 3)) return false;
  const int tid = threadIdx.x;
  const int nid = 32 * K * len, nfl = 48 * K * len;   // words: ids row (16 px * K * 2 per tile), float row (16 px * K * 3)
  uint32_t* pid = pat;
  float* pfl = reinterpret_cast<float*>(pat + nid);
  float* pim = pfl + nfl;
  for (int i = tid; i < nid + nfl + 64 * len; i += kRasterThreads) {
    if (i < nid) pid[i] = 0xffffffffu;
    else if (i < nid + nfl) pfl[i - nid] = -1.0f;
    else { const int e = (i - nid - nfl) & 3; pim[i - nid - nfl] = e < 3 ? bg[e] : 0.0f; }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk-copy engine
  __syncthreads();
  if (tid < 5 * kTileH) {
    const int row = tid / 5, which = tid - 5 * row;
    const size_t pix = ((size_t)n * H + (size_t)ty * kTileH + row) * W + (size_t)tx * kTileW;
    void* dst;
    const void* src;
    uint32_t bytes;
    if (which == 0) { dst = a.pix_to_face + pix * K; src = pid; bytes = 128u * K * len; }
    else if (which == 1) { dst = a.zbuf + pix * K; src = pfl; bytes = 64u * K * len; }
    else if (which == 2) { dst = a.dists + pix * K; src = pfl; bytes = 64u * K * len; }
    else if (which == 3) { dst = a.bary + pix * K * 3; src = pfl; bytes = 192u * K * len; }
    else { dst = image + pix * 4; src = pim; bytes = 256u * len; }
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
                 "r"((uint32_t)__cvta_generic_to_shared(src)), "r"(bytes), "l"(policy)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the patterns must outlive the reads
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Rasterizer workspace layout (32-bit words), filled by the setup pass (raster.cu) for Ftot packed faces / N meshes:
//   [0, Ftot + 64)                     packed tile range per face (pack_tile_range)
//   [box, box + 4 N)                   per-mesh tile box {txmin, 255 - txmax, tymin, 255 - tymax}   (only when N <= Ftot)
//   [loc, loc + Ftot)                  per face: exclusive prefix of the tile-range AREAS inside its 256-face block
//   [blk, blk + nblk + 1)              per 256-face block: exclusive prefix of the block totals; [blk + nblk] = total
// (face, tile) record index of the atomics-free backward = blk[f >> 8] + loc[f] + (ty - tymin) * (txmax - txmin + 1)
// + (tx - txmin): every (face, tile-in-its-range) pair owns one slot, so gradients are scattered without atomics and
// gathered per vertex in a fixed order (shade_bwd_tiled.cu, geom.cu).
struct WsLayout { int64_t box, loc, blk, nblk, words; };
__host__ __device__ __forceinline__ WsLayout ws_layout(int64_t Ftot) {
  WsLayout w;
  w.box = (Ftot + 64 + 3) & ~(int64_t)3;
  w.loc = w.box + 4 * Ftot;
  w.nblk = (Ftot + 255) >> 8;
  w.blk = w.loc + Ftot;
  w.words = w.blk + w.nblk + 4;
  return w;
}
__device__ __forceinline__ uint32_t face_rec_index(const uint32_t* __restrict__ ws, const WsLayout& L, int64_t fp, int tx, int ty) {
  const uint32_t r = __ldg(ws + fp);
  const int txmin = r & 255, txmax = (r >> 8) & 255, tymin = (r >> 16) & 255;
  return __ldg(ws + L.blk + (fp >> 8)) + __ldg(ws + L.loc + fp) + (uint32_t)((ty - tymin) * (txmax - txmin + 1) + (tx - txmin));
}

// host helpers defined in raster.cu
int launch_raster_setup(const HfrRasterArgs& a, uint32_t* ranges, cudaStream_t s, bool use_queue = false);
const uint32_t* raster_mesh_box(const HfrRasterArgs& a);
int check_raster(const HfrRasterArgs* a, const char* who);

}  // namespace hfr

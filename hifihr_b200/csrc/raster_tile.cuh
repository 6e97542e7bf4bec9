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
//           kernel) per step; hits are compacted IN FACE ORDER with a ballot/popc prefix, so the
//           tile list is sorted by face index and z-ties resolve to the smaller index with a
//           strict `<` (the CPU reference's (z, face) ordering).
//   stage:  the listed faces' 9 floats are gathered once per tile into 64-byte shared records
//           together with the dilated bbox and the barycentric denominator.
//   fine:   each warp first culls the tile list against its own 8x4 block (one ballot per 32
//           faces), then all lanes walk the surviving faces together; records are read with
//           broadcast LDS.128.
#pragma once
#include "raster_math.cuh"

namespace hfr {

constexpr int kTileW = 16, kTileH = 16, kRasterThreads = 256;
constexpr int kListCap = 512;   // faces per staged batch
constexpr int kRecFloats = 16;  // 64-byte record

struct RasterSmem {
  int list[kListCap];
  __align__(16) float rec[kListCap * kRecFloats];
  int wcount[2][kRasterThreads / 32];
};

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
  // replace the worst slot and bubble towards the front; strict < keeps earlier faces first on ties
  __device__ __forceinline__ void insert(float pz, int face) {
    z[KMAX - 1] = pz; f[KMAX - 1] = face;
#pragma unroll
    for (int i = KMAX - 1; i > 0; --i) {
      if (z[i] < z[i - 1]) {
        const float tz = z[i]; z[i] = z[i - 1]; z[i - 1] = tz;
        const int tf = f[i]; f[i] = f[i - 1]; f[i - 1] = tf;
      }
    }
  }
};

// Coarse + stage + fine for one tile.  On return `top` holds, per thread (= pixel), the KMAX
// nearest valid faces as packed face ids (sorted by (z, id)).
template <int KMAX>
__device__ __forceinline__ void raster_tile(const HfrRasterArgs& a, const uint32_t* __restrict__ tile_ranges,
                                            RasterSmem& sm, int n, int tx, int ty, float xf, float yf,
                                            bool pix_active, bool warp_active, float wx_lo, float wx_hi, float wy_lo, float wy_hi,
                                            TopK<KMAX>& top) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t f0 = a.mesh_first[n];
  const int nf = (int)a.mesh_nfaces[n];
  const float blur = a.blur_radius, rblur = sqrtf(a.blur_radius);
  const int pc = a.perspective_correct, clip = a.clip_barycentric;
  top.init();
  int count = 0;

  auto process = [&](int cnt) {
    __syncthreads();  // list complete
    for (int i = tid; i < cnt; i += kRasterThreads) {
      const float* __restrict__ v = a.face_verts + (size_t)sm.list[i] * 9;
      float r[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) r[e] = __ldg(v + e);
      float4* dst = reinterpret_cast<float4*>(sm.rec + i * kRecFloats);
      const float xmin = XSUB(hfr_min3(r[0], r[3], r[6]), rblur), xmax = XADD(hfr_max3(r[0], r[3], r[6]), rblur);
      const float ymin = XSUB(hfr_min3(r[1], r[4], r[7]), rblur), ymax = XADD(hfr_max3(r[1], r[4], r[7]), rblur);
      const float area = XADD(hfr_edge(r[6], r[7], r[0], r[1], r[3], r[4]), HFR_KEPS);
      dst[0] = make_float4(r[0], r[1], r[2], r[3]);
      dst[1] = make_float4(r[4], r[5], r[6], r[7]);
      dst[2] = make_float4(r[8], xmin, xmax, ymin);
      dst[3] = make_float4(ymax, area, 0.f, 0.f);
    }
    __syncthreads();
    if (warp_active) {
      for (int b0 = 0; b0 < cnt; b0 += 32) {
        const int i = b0 + lane;
        bool ok = false;
        if (i < cnt) {
          const float4 q2 = *reinterpret_cast<const float4*>(sm.rec + i * kRecFloats + 8);
          const float ymax = sm.rec[i * kRecFloats + 12];
          ok = !(q2.z < wx_lo || q2.y > wx_hi || ymax < wy_lo || q2.w > wy_hi);
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        while (m) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          const float4* r4 = reinterpret_cast<const float4*>(sm.rec + (b0 + j) * kRecFloats);
          const float4 q0 = r4[0], q1 = r4[1], q2 = r4[2], q3 = r4[3];
          if (pix_active && !(xf < q2.y || xf > q2.z || yf < q2.w || yf > q3.x)) {
            const float v[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
            float pz, bc[3];
            bool inside;
            if (hfr_raster_bary(xf, yf, v, q3.y, pc, clip, &pz, bc, &inside)) {
              if (pz < top.worst()) {
                bool keep = inside;
                if (!keep) keep = hfr_tri_dist2(xf, yf, v) < blur;
                if (keep) top.insert(pz, sm.list[b0 + j]);
              }
            }
          }
        }
      }
    }
    __syncthreads();  // records consumed before the list is rebuilt
  };

  for (int base = 0; base < nf; base += kRasterThreads) {
    const int fi = base + tid;
    bool hit = false;
    if (fi < nf) {
      const uint32_t w = __ldg(tile_ranges + f0 + fi);
      const int txmin = w & 255, txmax = (w >> 8) & 255, tymin = (w >> 16) & 255, tymax = w >> 24;
      hit = tx >= txmin && tx <= txmax && ty >= tymin && ty <= tymax;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int par = (base / kRasterThreads) & 1;
    if (lane == 0) sm.wcount[par][warp] = __popc(bal);
    __syncthreads();
    int prefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kRasterThreads / 32; ++w) {
      const int c = sm.wcount[par][w];
      prefix += (w < warp) ? c : 0;
      total += c;
    }
    if (hit) sm.list[count + prefix + __popc(bal & ((1u << lane) - 1))] = (int)(f0 + fi);
    count += total;
    if (count > kListCap - kRasterThreads) {
      process(count);
      count = 0;
    }
  }
  if (count > 0) process(count);
}


// Recompute the winners' barycentrics / depth / distance from the packed face floats.
template <int KMAX>
__device__ __forceinline__ void compute_fragments(const HfrRasterArgs& a, float xf, float yf, const TopK<KMAX>& top,
                                                  int64_t* id, float* z, float* d, float* b) {
  const int K = a.K;
  const float rblur = sqrtf(a.blur_radius);
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    id[k] = -1; z[k] = -1.0f; d[k] = -1.0f; b[3 * k] = b[3 * k + 1] = b[3 * k + 2] = -1.0f;
    if (k < K && top.f[k] >= 0) {
      float v[9];
      const float* __restrict__ src = a.face_verts + (size_t)top.f[k] * 9;
#pragma unroll
      for (int e = 0; e < 9; ++e) v[e] = __ldg(src + e);
      float pz, bc[3], sd;
      if (hfr_raster_eval(xf, yf, v, a.blur_radius, rblur, a.perspective_correct, a.clip_barycentric,
                          a.cull_backfaces, &pz, bc, &sd)) {
        id[k] = top.f[k]; z[k] = pz; d[k] = sd; b[3 * k] = bc[0]; b[3 * k + 1] = bc[1]; b[3 * k + 2] = bc[2];
      }
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

struct PixelCtx {
  int n, tx, ty, xi, yi;
  float xf, yf, wx_lo, wx_hi, wy_lo, wy_hi;
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
  c.xf = hfr_pix_to_ndc(W - 1 - c.xi, W, H);
  c.yf = hfr_pix_to_ndc(H - 1 - c.yi, H, W);
  const int wx1 = min(wx0 + 7, W - 1), wy1 = min(wy0 + 3, H - 1);
  c.wx_hi = hfr_pix_to_ndc(W - 1 - wx0, W, H);
  c.wx_lo = hfr_pix_to_ndc(W - 1 - wx1, W, H);
  c.wy_hi = hfr_pix_to_ndc(H - 1 - wy0, H, W);
  c.wy_lo = hfr_pix_to_ndc(H - 1 - wy1, H, W);
  return c;
}

// host helpers defined in raster.cu
int launch_raster_setup(const HfrRasterArgs& a, uint32_t* ranges, cudaStream_t s);
int check_raster(const HfrRasterArgs* a, const char* who);

}  // namespace hfr

// Mesh rasterizer for sm_100a: face setup, tile-binned forward (Fragments), backward.
//
// Replaces pytorch3d._C.rasterize_meshes / rasterize_meshes_backward as reached from
// MeshRasterizer.forward (models_res_nimble.py:208) — semantics in SURVEY.md Appendix A.2-A.5.
// Outputs are bit-identical to the scalar CPU oracle (oracle/raster_naive.c): same fp32
// operation order, no FMA contraction in the coverage / depth / distance math.
#include "common.cuh"
#include "raster_tile.cuh"

namespace hfr {

// One thread per packed face: conservative range of 16x16 tiles its dilated bbox can touch, and the
// union of those ranges per mesh (4 min-reduced words per mesh, the two upper bounds stored as 255 - x,
// initialised to 0xff by a memset), so that tiles outside a mesh's footprint skip the coarse scan.
__global__ void __launch_bounds__(256) raster_setup_kernel(HfrRasterArgs a, uint32_t* __restrict__ ranges,
                                                           uint32_t* __restrict__ mesh_box, uint32_t* __restrict__ rec_loc,
                                                           uint32_t* __restrict__ blk_tot, uint32_t* __restrict__ tile_cnt) {
  __shared__ uint32_t s_wsum[8];
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = f < a.Ftot;
  uint32_t out = kEmptyRange;
  int tx0 = 255, tx1 = 0, ty0 = 255, ty1 = 0;
  if (live) {
    float v[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) v[e] = __ldg(a.face_verts + f * 9 + e);
    if (hfr_face_valid(v, a.cull_backfaces)) {
      const float r = sqrtf(a.blur_radius);
      const float xmin = hfr_min3(v[0], v[3], v[6]) - r, xmax = hfr_max3(v[0], v[3], v[6]) + r;
      const float ymin = hfr_min3(v[1], v[4], v[7]) - r, ymax = hfr_max3(v[1], v[4], v[7]) + r;
      // invert hfr_pix_to_ndc: i = ((x + off) * S1 - off) / range ; pixel column = S1 - 1 - i
      const float rx = a.W > a.H ? 2.0f * a.W / a.H : 2.0f, ry = a.H > a.W ? 2.0f * a.H / a.W : 2.0f;
      const float ox = 0.5f * rx, oy = 0.5f * ry;
      const float ix_hi = ((xmax + ox) * a.W - ox) / rx, ix_lo = ((xmin + ox) * a.W - ox) / rx;
      const float iy_hi = ((ymax + oy) * a.H - oy) / ry, iy_lo = ((ymin + oy) * a.H - oy) / ry;
      // one extra pixel of slack on each side absorbs the rounding of this inverse map
      float cx0 = floorf((float)(a.W - 1) - ix_hi) - 1.0f, cx1 = ceilf((float)(a.W - 1) - ix_lo) + 1.0f;
      float cy0 = floorf((float)(a.H - 1) - iy_hi) - 1.0f, cy1 = ceilf((float)(a.H - 1) - iy_lo) + 1.0f;
      if (cx1 >= 0.0f && cy1 >= 0.0f && cx0 <= (float)(a.W - 1) && cy0 <= (float)(a.H - 1) && cx0 == cx0 &&
          cx1 == cx1 && cy0 == cy0 && cy1 == cy1) {
        const int x0 = (int)fmaxf(cx0, 0.0f), x1 = (int)fminf(cx1, (float)(a.W - 1));
        const int y0 = (int)fmaxf(cy0, 0.0f), y1 = (int)fminf(cy1, (float)(a.H - 1));
        tx0 = x0 / kTileW; tx1 = x1 / kTileW; ty0 = y0 / kTileH; ty1 = y1 / kTileH;
        out = pack_tile_range(tx0, tx1, ty0, ty1);
      }
    }
    ranges[f] = out;
  }
  {   // exclusive prefix of the range areas inside this 256-face block (record slots of the atomics-free backward)
    const uint32_t area = (live && out != kEmptyRange) ? (uint32_t)((tx1 - tx0 + 1) * (ty1 - ty0 + 1)) : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = area;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = s_wsum[w]; before += (w < warp) ? c : 0u; total += c; }
    if (live) rec_loc[f] = before + incl - area;
    if (threadIdx.x == 0) blk_tot[blockIdx.x] = total;
  }
  if (mesh_box == nullptr) return;
  // mesh of this face: last n with mesh_first[n] <= f (meshes are packed in order)
  int n = -1;
  if (live && out != kEmptyRange) {
    int lo = 0, hi = a.N - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(a.mesh_first + mid) <= f) lo = mid; else hi = mid - 1;
    }
    n = lo;
    if (f >= __ldg(a.mesh_first + n) + __ldg(a.mesh_nfaces + n)) n = -1;   // a gap between meshes
  }
  const int n0 = __shfl_sync(0xffffffffu, n, 0);
  if (__all_sync(0xffffffffu, n == n0 || n < 0)) {   // the common case: one mesh per warp -> 4 atomics per warp
    const unsigned m0 = __reduce_min_sync(0xffffffffu, (unsigned)(n < 0 ? 255 : tx0));
    const unsigned m1 = __reduce_min_sync(0xffffffffu, (unsigned)(n < 0 ? 255 : 255 - tx1));
    const unsigned m2 = __reduce_min_sync(0xffffffffu, (unsigned)(n < 0 ? 255 : ty0));
    const unsigned m3 = __reduce_min_sync(0xffffffffu, (unsigned)(n < 0 ? 255 : 255 - ty1));
    const int nn = __reduce_max_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && nn >= 0) {
      atomicMin(mesh_box + 4 * nn, m0); atomicMin(mesh_box + 4 * nn + 1, m1);
      atomicMin(mesh_box + 4 * nn + 2, m2); atomicMin(mesh_box + 4 * nn + 3, m3);
    }
  } else if (n >= 0) {
    atomicMin(mesh_box + 4 * n, (unsigned)tx0); atomicMin(mesh_box + 4 * n + 1, (unsigned)(255 - tx1));
    atomicMin(mesh_box + 4 * n + 2, (unsigned)ty0); atomicMin(mesh_box + 4 * n + 3, (unsigned)(255 - ty1));
  }
  // tile queue: how many faces' ranges cover each tile (the cost proxy the tiles are ordered by; integer atomics)
  if (tile_cnt && n >= 0) {
    const int TX = (a.W + kTileW - 1) / kTileW, TY = (a.H + kTileH - 1) / kTileH;
    uint32_t* c = tile_cnt + (size_t)n * TX * TY;
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(c + ty * TX + tx, 1u);
  }
}

// Tile queue layout (32-bit words), T = N * TX * TY tiles:
//   [0] nc = tiles with at least one face, [1] ng = groups of empty tiles, [2, 2 + 8) = tiles per cost class
//   [16, 16 + T)            face count per tile (n-major, row-major), zeroed with the header every launch
//   [16 + T, 16 + 9 T)      8 class lists of tile ids (class c = floor(log2(face count)), capped at 7), capacity T each
//   [16 + 9 T, 16 + 10 T)   groups of up to kFillRun (8) horizontally adjacent empty tiles: first tile id | (len - 1) << 28
// Orders inside a list depend on atomic timing; no result depends on the order tiles are processed in.
__device__ __forceinline__ int cost_class(uint32_t cnt) { return min(kCostClasses - 1, 31 - __clz(cnt)); }   // cnt > 0
__global__ void __launch_bounds__(256) raster_order_kernel(uint32_t* __restrict__ q, int N, int TX, int TY, int max_run) {
  // one thread per tile row.  Pass 1 counts the row's entries per list, the block reserves its share of every list with
  // ONE global atomic per list, pass 2 writes the entries.
  __shared__ uint32_t s_cnt[kCostClasses + 1], s_base[kCostClasses + 1];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < N * TY;
  const int T = N * TX * TY;
  const uint32_t* cnt = q + kQueueHdr + (size_t)(live ? r : 0) * TX;
  uint32_t* lists = q + kQueueHdr + T;
  uint32_t* groups = q + kQueueHdr + (size_t)(1 + kCostClasses) * T;
  if (threadIdx.x <= kCostClasses) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t mine[kCostClasses + 1];
#pragma unroll
  for (int k = 0; k <= kCostClasses; ++k) mine[k] = 0;
  if (live) {
    int run = 0;
    for (int tx = 0; tx <= TX; ++tx) {
      const uint32_t c = tx < TX ? cnt[tx] : 1u;
      if (c == 0) {
        if (++run == max_run) { ++mine[kCostClasses]; run = 0; }
      } else {
        if (run > 0) { ++mine[kCostClasses]; run = 0; }
        if (tx < TX) {
          const int k = cost_class(c);
#pragma unroll
          for (int j = 0; j < kCostClasses; ++j) mine[j] += (j == k) ? 1u : 0u;
        }
      }
    }
  }
  uint32_t off[kCostClasses + 1];
#pragma unroll
  for (int k = 0; k <= kCostClasses; ++k) off[k] = mine[k] ? atomicAdd(&s_cnt[k], mine[k]) : 0u;
  __syncthreads();
  if (threadIdx.x <= kCostClasses) {
    const int k = threadIdx.x;
    const uint32_t c = s_cnt[k];
    s_base[k] = c ? atomicAdd(k == kCostClasses ? q + 1 : q + 2 + k, c) : 0u;
    if (k < kCostClasses && c) atomicAdd(q, c);
  }
  __syncthreads();
  if (!live) return;
  int run0 = -1;
  for (int tx = 0; tx <= TX; ++tx) {
    const uint32_t c = tx < TX ? cnt[tx] : 1u;
    if (c == 0) {
      if (run0 < 0) run0 = tx;
      if (tx - run0 + 1 == max_run) {
        groups[s_base[kCostClasses] + off[kCostClasses]++] = (uint32_t)(r * TX + run0) | ((uint32_t)(max_run - 1) << 28);
        run0 = -1;
      }
    } else {
      if (run0 >= 0) {
        groups[s_base[kCostClasses] + off[kCostClasses]++] = (uint32_t)(r * TX + run0) | ((uint32_t)(tx - run0 - 1) << 28);
        run0 = -1;
      }
      if (tx < TX) {
        const int k = cost_class(c);
        uint32_t pos = 0;
#pragma unroll
        for (int j = 0; j < kCostClasses; ++j)
          if (j == k) pos = s_base[j] + off[j]++;
        lists[(size_t)k * T + pos] = (uint32_t)(r * TX + tx);
      }
    }
  }
}

// exclusive scan of the per-block record totals, in place (one CTA; nblk + 1 entries, the last one receives the total)
__global__ void __launch_bounds__(1024) raster_scan_kernel(uint32_t* __restrict__ blk, int nblk) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + tid;
    const uint32_t v = i < nblk ? blk[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const uint32_t c = s_w[w]; before += (w < warp) ? c : 0u; total += c; }
    const uint32_t carry = s_carry;
    if (i < nblk) blk[i] = carry + before + incl - v;
    __syncthreads();
    if (tid == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (tid == 0) blk[nblk] = s_carry;
}

#ifndef HFR_RASTERONLY_MINB
#define HFR_RASTERONLY_MINB 5   // C3 standalone rasterizer (B=128, 256^2, 11968 faces): 3 CTAs/SM 2.83 ms, 4: 2.47, 5: 2.34
#endif
template <int KMAX>
__global__ void __launch_bounds__(kRasterThreads, (KMAX <= 8 ? HFR_RASTERONLY_MINB : 2)) raster_fwd_kernel(HfrRasterArgs a, const uint32_t* __restrict__ ranges,
                                                                    const uint32_t* __restrict__ mesh_box) {
  __shared__ RasterSmem sm;
  PixelCtx c = make_pixel_ctx(a.H, a.W);
  // K = 1 / 4: a tile no face touches streams its -1 Fragments as whole rows (fixed-role fills, see shade.cu)
  if ((a.K == 1 || (a.K & 3) == 0) && tile_outside_mesh(mesh_box, c.n, c.tx, c.ty) && fill_empty_tile(a, c.n, c.tx, c.ty)) return;
  TopK<KMAX> top;
  uint32_t perm;
  raster_tile<KMAX, false>(a, ranges, mesh_box, sm, c, top, nullptr, perm);
  if (c.pix_active) {
    int64_t id[KMAX];
    float z[KMAX], d[KMAX], b[KMAX * 3];
    compute_fragments<KMAX>(a, c.xf, c.yf, top, id, z, d, b);
    store_fragments<KMAX>(a, ((size_t)c.n * a.H + c.yi) * a.W + c.xi, id, z, d, b);
  }
}

// Backward: one thread per pixel, K fragments each; 9 reductions per fragment into the packed
// per-face gradient (the upstream contract).  The fused path (shade.cu) scatters per vertex.
__global__ void __launch_bounds__(256) raster_bwd_kernel(HfrRasterBwdArgs a) {
  const size_t P = (size_t)a.N * a.H * a.W;
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += (size_t)gridDim.x * blockDim.x) {
    const int xi = (int)(pix % a.W), yi = (int)((pix / a.W) % a.H);
    const float xf = hfr_pix_to_ndc(a.W - 1 - xi, a.W, a.H), yf = hfr_pix_to_ndc(a.H - 1 - yi, a.H, a.W);
    for (int k = 0; k < a.K; ++k) {
      const int64_t f = a.pix_to_face[pix * a.K + k];
      if (f < 0) continue;
      float v[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) v[e] = __ldg(a.face_verts + f * 9 + e);
      float gb[3] = {0.f, 0.f, 0.f};
      if (a.g_bary) { gb[0] = a.g_bary[(pix * a.K + k) * 3]; gb[1] = a.g_bary[(pix * a.K + k) * 3 + 1]; gb[2] = a.g_bary[(pix * a.K + k) * 3 + 2]; }
      const float gz = a.g_zbuf ? a.g_zbuf[pix * a.K + k] : 0.0f;
      const float gd = a.g_dists ? a.g_dists[pix * a.K + k] : 0.0f;
      float gv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      hfr_raster_eval_bwd(xf, yf, v, a.perspective_correct, a.clip_barycentric, gb, gz, gd, gv);
#pragma unroll
      for (int e = 0; e < 9; ++e)
        if (gv[e] != 0.0f) atomicAdd(a.g_face_verts + f * 9 + e, gv[e]);
    }
  }
}

int check_raster(const HfrRasterArgs* a, const char* who) {
  HFR_CHECK_ARG(a && a->N >= 0 && a->H > 0 && a->W > 0, "%s: bad image size", who);
  HFR_CHECK_ARG(a->K >= 1 && a->K <= HFR_MAX_K, "%s: faces_per_pixel must be in [1,%d], got %d", who, HFR_MAX_K, a->K);
  HFR_CHECK_ARG((a->W + kTileW - 1) / kTileW <= 255 && (a->H + kTileH - 1) / kTileH <= 255,
                "%s: image side above %d px unsupported", who, 255 * kTileW);
  HFR_CHECK_ARG(a->Ftot >= 0 && a->Ftot < (1ll << 31), "%s: bad face count", who);
  HFR_CHECK_ARG(a->blur_radius >= 0.0f, "%s: blur_radius must be >= 0", who);
  HFR_CHECK_ARG(a->N == 0 || (a->face_verts && a->mesh_first && a->mesh_nfaces && a->pix_to_face && a->zbuf &&
                               a->bary && a->dists && a->workspace),
                "%s: null pointer", who);
  return HFR_OK;
}

// workspace layout: ws_layout() in raster_tile.cuh; the tile box only when N <= Ftot (hfr_raster_workspace_bytes sizes
// the buffer for that)
const uint32_t* raster_mesh_box(const HfrRasterArgs& a) {
  return (a.N <= a.Ftot) ? reinterpret_cast<const uint32_t*>(a.workspace) + ws_layout(a.Ftot).box : nullptr;   // 16-byte aligned
}

int launch_raster_setup(const HfrRasterArgs& a, uint32_t* ranges, cudaStream_t s, bool use_queue) {
  if (a.Ftot > 0) {
    uint32_t* box = const_cast<uint32_t*>(raster_mesh_box(a));
    if (box) cudaMemsetAsync(box, 0xff, (size_t)a.N * 4 * sizeof(uint32_t), s);
    const WsLayout L = ws_layout(a.Ftot);
    uint32_t* ws = reinterpret_cast<uint32_t*>(a.workspace);
    uint32_t* queue = (box && use_queue) ? reinterpret_cast<uint32_t*>(a.tile_queue) : nullptr;
    const int TX = (a.W + kTileW - 1) / kTileW, TY = (a.H + kTileH - 1) / kTileH, T = a.N * TX * TY;
    if (queue) cudaMemsetAsync(queue, 0, (size_t)(kQueueHdr + T) * sizeof(uint32_t), s);
    raster_setup_kernel<<<(unsigned)L.nblk, 256, 0, s>>>(a, ranges, box, ws + L.loc, ws + L.blk, queue ? queue + kQueueHdr : nullptr);
    HFR_CHECK_LAUNCH("raster_setup");
    raster_scan_kernel<<<1, 1024, 0, s>>>(ws + L.blk, (int)L.nblk);
    HFR_CHECK_LAUNCH("raster_scan");
    if (queue) {
      // empty tiles per fill group: bounded by the 16 K-word pattern buffer of the bulk-copy fill ((80 K + 64) words per tile)
      const int max_run = a.K <= 4 ? kFillRun : (a.K <= 8 ? kFillRun / 2 : kFillRun / 4);
      raster_order_kernel<<<(a.N * TY + 255) / 256, 256, 0, s>>>(queue, a.N, TX, TY, max_run);
      HFR_CHECK_LAUNCH("raster_order");
    }
  }
  return HFR_OK;
}

template <int KMAX>
static void launch_fwd(const HfrRasterArgs& a, const uint32_t* ranges, cudaStream_t s) {
  dim3 grid((a.W + kTileW - 1) / kTileW, (a.H + kTileH - 1) / kTileH, a.N);
  raster_fwd_kernel<KMAX><<<grid, kRasterThreads, 0, s>>>(a, ranges, raster_mesh_box(a));
}

}  // namespace hfr

extern "C" int64_t hfr_raster_queue_bytes(int32_t N, int32_t H, int32_t W) {
  if (N < 1 || H < 1 || W < 1) return 64;
  const int64_t T = (int64_t)N * ((W + hfr::kTileW - 1) / hfr::kTileW) * ((H + hfr::kTileH - 1) / hfr::kTileH);
  return (hfr::kQueueHdr + (2 + hfr::kCostClasses) * T) * 4 + 64;
}

extern "C" int64_t hfr_raster_workspace_bytes(int64_t Ftot) { return hfr::ws_layout(Ftot < 1 ? 1 : Ftot).words * 4 + 64; }

extern "C" const uint32_t* hfr_raster_tile_box(const void* workspace, int64_t Ftot, int32_t N) {
  if (!workspace || N <= 0 || N > Ftot) return nullptr;
  return reinterpret_cast<const uint32_t*>(workspace) + hfr::ws_layout(Ftot).box;
}

extern "C" int hfr_raster_forward(const HfrRasterArgs* a, void* stream) {
  using namespace hfr;
  if (int rc = check_raster(a, "raster_forward")) return rc;
  if (a->N == 0) return HFR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* ranges = reinterpret_cast<uint32_t*>(a->workspace);
  if (int rc = launch_raster_setup(*a, ranges, s)) return rc;
  if (a->K == 1) launch_fwd<1>(*a, ranges, s);
  else if (a->K == 2) launch_fwd<2>(*a, ranges, s);
  else if (a->K <= 4) launch_fwd<4>(*a, ranges, s);
  else if (a->K <= 8) launch_fwd<8>(*a, ranges, s);
  else launch_fwd<16>(*a, ranges, s);
  HFR_CHECK_LAUNCH("raster_forward");
  return HFR_OK;
}

extern "C" int hfr_raster_backward(const HfrRasterBwdArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->N >= 0 && a->H > 0 && a->W > 0 && a->K >= 1 && a->K <= HFR_MAX_K, "raster_backward: bad dims");
  if (a->N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->face_verts && a->pix_to_face && a->g_face_verts, "raster_backward: null pointer");
  const size_t P = (size_t)a->N * a->H * a->W;
  const unsigned blocks = (unsigned)((P + 255) / 256 < 148u * 32u ? (P + 255) / 256 : 148u * 32u);
  raster_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("raster_backward");
  return HFR_OK;
}

// Batched hand layer for sm_100a: per-sample pose stage, ONE tensor-core blend contraction per batch, vertex-parallel
// skinning - forward and backward.  Replaces the same reference code as mano.cu (ManoLayer.forward,
// utils/my_mano.py:315-483; the two blend products are :386-393) and produces the same numbers to fp32 rounding.
//
//   forward : hand_prep_kernel   (1 CTA / sample)   pose PCA, Rodrigues, joints, kinematic chain -> coef, A, joints
//             blend_fwd_kernel   (1 CTA / 32 basis columns x 128 samples)  v_posed = coef . basis  on tcgen05
//             hand_skin_kernel   (vertex-parallel)  verts = sum_j w_vj A_j [v_posed; 1] + offset
//   backward: hand_skin_bwd_kernel (1 CTA / sample) g_v_posed = T^T g_v, d/d(A_j) joint-major, joint-gradient routing
//             blend_bwd_kernel   (split over 3V)     g_coef partials = g_v_posed . basis^T  on tcgen05
//             hand_chain_bwd_kernel (1 CTA / sample) fixed-order sum of the partials, chain', Rodrigues', PCA', shape'
//
// Tensor-core numerics: kind::tf32 keeps 11 significant bits per operand, so each fp32 operand x is split into
// hi = x with the low 13 mantissa bits cleared and lo = (x - hi) likewise; a.b ~ hi_a hi_b + lo_a hi_b + hi_a lo_b
// (three MMAs, fp32 accumulation in TMEM) leaves a relative error of ~2^-21 per product - the blend offsets are
// millimetres, so vertices move by < 1e-9 m against the fp32 FMA path (tests pin 1e-6 m).
//
// Shared-memory operand layout (UMMA canonical K-major, no swizzle): a tile of R rows x K columns is stored as
// K/4 chunks of R x 16 bytes; rows of a chunk are 16 B apart (8 rows = one 128-byte core matrix, SBO = 128 B), chunks
// are R * 16 B apart (LBO).  One MMA (K = 8) reads two neighbouring chunks.  The constant basis is stored in global
// memory already in this layout, hi and lo tiles back to back (hfr_mano_pack_basis), so one cp.async.bulk per tile
// brings it in; the per-batch operand (coefficients / g_v_posed) is split and written in the same layout by the
// kernel that produces it (pose stage / skinning backward), so it arrives by cp.async.bulk as well.
#include "mano_batched.cuh"
#include "mano_math.cuh"

namespace hfr {

__device__ unsigned int g_batched_timeout = 0;   // set when an mbarrier wait gave up (never expected; guards against hangs)

#ifdef HFR_MANO_TIMING   // tuning builds only: clock64 of block 0 at phase boundaries of the small per-sample kernels
__device__ long long g_bt[3][16];
#define BT(kern, i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_bt[kern][i] = clock64(); } while (0)
#else
#define BT(kern, i) do { } while (0)
#endif

namespace {

constexpr int kPrepThreads = 256;
constexpr int kGemmThreads = 128;     // 4 warps: one TMEM lane quadrant each
constexpr int kSkinBwdThreads = 512;
constexpr int kChainThreads = 256;
constexpr int kMTile = 128;           // samples per GEMM CTA (UMMA M)
constexpr int kFwdN = 32;             // basis columns per forward CTA (UMMA N)
constexpr int kBwdK = 64;             // reduction columns per backward chunk

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol error must not hang the GPU (the flag is checked by the tests)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 400000000LL) { atomicExch(&g_batched_timeout, 1u); break; }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {   // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, M x N x 8 tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor: K-major, no swizzle (layout in the file header)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: tf32 x tf32 -> fp32, both operands K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
  hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
  lo = make_float4(tf32_hi(x.x - hi.x), tf32_hi(x.y - hi.y), tf32_hi(x.z - hi.z), tf32_hi(x.w - hi.w));
}

// sum of term(k) for k = start, start + stride, ... < n with U loads in flight per round: the loop is unrolled over a FIXED
// count with predicated terms (a runtime trip count would leave nvcc's unrolled body unused and fall into a remainder
// loop that waits for one load at a time)
template <int U, typename F>
__device__ __forceinline__ float dot_pred(int n, int start, int stride, F term) {
  float acc = 0.0f;
  for (int k0 = start; k0 < n; k0 += U * stride) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * stride;
      v[u] = k < n ? term(k) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------- workspace
struct WsPtrs {
  float* state; float* coefp; float* off; float* vp; float* gvpp; float* gA; float* gGt; float* gcoefp;
  int ksplit;
  size_t floats;
};
inline int bwd_ksplit(const BatchedDims& d, int B) {
  const int mt = (B + kMTile - 1) / kMTile;
  int ks = (2 * 148 + mt - 1) / mt;        // enough CTAs for two waves of the chip, at most one chunk each
  ks = ks < 1 ? 1 : ks;
  return ks > d.NCH64 ? d.NCH64 : ks;
}
inline WsPtrs carve_ws(const HfrHandModel& m, const BatchedDims& d, int B, void* ws) {
  WsPtrs p;
  float* f = reinterpret_cast<float*>(ws);
  size_t o = 0;
  auto take = [&](size_t n) { float* r = f ? f + o : nullptr; o += (n + 31) & ~(size_t)31; return r; };
  p.state = take((size_t)B * d.ST);
  const size_t mt = (size_t)(B + kMTile - 1) / kMTile;
  p.coefp = take(mt * 2 * d.KP * kMTile);          // per 128-sample tile: hi | lo, each [KP/4][128][4] (MMA operand layout)
  p.off = take((size_t)B * 4);
  p.vp = take((size_t)B * d.C3P);
  p.gvpp = take(mt * d.NCH64 * 2 * kBwdK * kMTile);  // per (tile, 64-column chunk): hi | lo, each [16][128][4]
  p.gA = take((size_t)B * 12 * m.NJ);
  p.gGt = take((size_t)B * 4 * m.NJ);
  p.ksplit = bwd_ksplit(d, B);
  p.gcoefp = take((size_t)p.ksplit * B * d.NKP16);
  p.floats = o;
  return p;
}

// offsets inside a sample's pose state
struct StateOff { int full, R, J, G, A; };
__host__ __device__ inline StateOff state_off(int NJ) { return StateOff{0, 3 * NJ, 12 * NJ, 15 * NJ, 27 * NJ}; }

// ---------------------------------------------------------------------------------------------- pose stage
struct PoseSmem {
  float full[3 * HFR_MAX_JOINTS], R[9 * HFR_MAX_JOINTS], J[3 * HFR_MAX_JOINTS], G[12 * HFR_MAX_JOINTS], A[12 * HFR_MAX_JOINTS];
  float coef[64 + 9 * HFR_MAX_JOINTS];
  int depth[HFR_MAX_JOINTS], parent[HFR_MAX_JOINTS];
};

// pose -> full axis-angle, R, pose map, J, chain G, A  (the math of mano.cu's mano_setup for any CTA size)
template <int T>
__device__ void pose_setup(const HfrHandModel& m, PoseSmem& s, const float* __restrict__ pose, const float* __restrict__ betas,
                           const float* __restrict__ rots, int n_rot, int poff, const int* __restrict__ depth_tab) {
  const int tid = threadIdx.x, NJ = m.NJ, NPOSE = 3 * (NJ - 1);
  // full axis-angle pose: 4 lanes per output share the PCA dot product (12 independent loads each for MANO), joined
  // by two shuffles in a fixed order
  for (int base = 0; base < 12 * NJ; base += T) {
    const int idx = base + tid, i = idx >> 2, part = idx & 3;
    float h = 0.0f;
    const bool live = i >= 3 && i < 3 * NJ && n_rot < NJ;
    if (live && m.NPC > 0) {
      const int o = i - 3;
      h = dot_pred<12>(m.NPC, part, 4, [&](int k) { return pose[poff + k] * __ldg(m.pca_comps + k * NPOSE + o); });
    }
    h += __shfl_xor_sync(0xffffffffu, h, 1);
    h += __shfl_xor_sync(0xffffffffu, h, 2);
    if (part == 0 && i < 3 * NJ) {
      float v = 0.0f;
      if (n_rot < NJ) {
        if (i < 3) {
          v = n_rot > 0 ? 0.0f : pose[i];
        } else {
          v = (m.pose_mean ? m.pose_mean[i - 3] : 0.0f) + (m.NPC > 0 ? h : pose[poff + i - 3]);
        }
      }
      s.full[i] = v;
    }
  }
  for (int i = tid; i < m.NS; i += T) s.coef[i] = betas ? betas[i] : 0.0f;
  for (int j = tid; j < NJ; j += T) {
    s.depth[j] = depth_tab[j];
    s.parent[j] = m.parents[j];
  }
  __syncthreads();
  BT(0, 4);
  for (int j = tid; j < NJ; j += T) {
    if (j < n_rot) {
      for (int e = 0; e < 9; ++e) s.R[9 * j + e] = rots[9 * j + e];
    } else {
      hfr_rodrigues_fwd(s.full + 3 * j, s.R + 9 * j);
    }
  }
  for (int i = tid; i < 3 * NJ; i += T) {
    s.J[i] = m.J_template[i] + dot_pred<10>(m.NS, 0, 1, [&](int k) { return __ldg(m.J_shapedirs + i * m.NS + k) * s.coef[k]; });
  }
  __syncthreads();
  BT(0, 5);
  for (int i = tid; i < 9 * (NJ - 1); i += T) {
    const int e = i % 9;
    s.coef[m.NS + i] = s.R[9 + i] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
  }
  // kinematic chain one tree level at a time, 12 threads per joint
  int maxd = 0;
  for (int j = 0; j < NJ; ++j) maxd = max(maxd, s.depth[j]);
  for (int d = 0; d <= maxd; ++d) {
    for (int idx = tid; idx < 12 * NJ; idx += T) {
      const int j = idx / 12, e = idx - 12 * j, r = e >> 2, c = e & 3;
      if (s.depth[j] != d) continue;
      const int p = s.parent[j];
      const float* Rj = s.R + 9 * j;
      float val;
      if (p < 0) {
        val = c < 3 ? Rj[r * 3 + c] : s.J[3 * j + r];
      } else {
        const float* P = s.G + 12 * p;
        if (c < 3) {
          val = P[r * 4 + 0] * Rj[0 * 3 + c] + P[r * 4 + 1] * Rj[1 * 3 + c] + P[r * 4 + 2] * Rj[2 * 3 + c];
        } else {
          const float t0 = s.J[3 * j + 0] - s.J[3 * p + 0], t1 = s.J[3 * j + 1] - s.J[3 * p + 1],
                      t2 = s.J[3 * j + 2] - s.J[3 * p + 2];
          val = P[r * 4 + 0] * t0 + P[r * 4 + 1] * t1 + P[r * 4 + 2] * t2 + P[r * 4 + 3];
        }
      }
      s.G[idx] = val;
    }
    __syncthreads();
  }
  BT(0, 6);
  for (int i = tid; i < 3 * NJ; i += T) {   // A_j = G_j with the rest joint removed
    const int j = i / 3, r = i % 3;
    const float* G = s.G + 12 * j;
    const float* Jj = s.J + 3 * j;
    float* A = s.A + 12 * j;
    A[r * 4 + 0] = G[r * 4 + 0];
    A[r * 4 + 1] = G[r * 4 + 1];
    A[r * 4 + 2] = G[r * 4 + 2];
    A[r * 4 + 3] = G[r * 4 + 3] - (G[r * 4 + 0] * Jj[0] + G[r * 4 + 1] * Jj[1] + G[r * 4 + 2] * Jj[2]);
  }
  __syncthreads();
}

__device__ __forceinline__ void skin_matrix_b(const HfrHandModel& m, const float* A, int v, float* T) {
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.0f;
  float w[8];
  int ji[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = i < m.NW ? __ldg(m.skin_w + i * m.V + v) : 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ji[i] = w[i] != 0.0f ? __ldg(m.skin_idx + i * m.V + v) : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (w[i] != 0.0f) {
      const float* Aj = A + 12 * ji[i];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] += w[i] * Aj[e];
    }
  }
}

// forward pose stage: writes the pose state, the padded coefficient row, the joint outputs and the centring offset
__global__ void __launch_bounds__(kPrepThreads) hand_prep_kernel(HfrHandModel m, HfrManoFwdArgs a, int pose_dim, BatchedDims d,
                                                                  WsPtrs w, int want_outputs, const int* __restrict__ depth_tab) {
  __shared__ PoseSmem s;
  __shared__ float s_tipvp[3 * 24], s_tip[3 * 24], s_off[4];
  const int b = blockIdx.x, tid = threadIdx.x, NJ = m.NJ, lane = tid & 31, warp = tid >> 5;
  const int poff = a.pose_off > 0 ? a.pose_off : 3;
  BT(0, 0);
  pose_setup<kPrepThreads>(m, s, a.pose ? a.pose + (size_t)b * pose_dim : nullptr, a.betas ? a.betas + (size_t)b * m.NS : nullptr,
                           a.rots ? a.rots + (size_t)b * a.n_rot_in * 9 : nullptr, a.rots ? a.n_rot_in : 0, poff, depth_tab);
  BT(0, 1);
  const StateOff so = state_off(NJ);
  float* st = w.state + (size_t)b * d.ST;
  for (int i = tid; i < 3 * NJ; i += kPrepThreads) { st[so.full + i] = s.full[i]; st[so.J + i] = s.J[i]; }
  for (int i = tid; i < 9 * NJ; i += kPrepThreads) st[so.R + i] = s.R[i];
  for (int i = tid; i < 12 * NJ; i += kPrepThreads) { st[so.G + i] = s.G[i]; st[so.A + i] = s.A[i]; }
  {   // the sample's coefficient row, split into tf32 hi / lo parts, straight into the MMA operand layout
    float* tile = w.coefp + (size_t)(b >> 7) * 2 * d.KP * kMTile;
    const int row = b & (kMTile - 1);
    for (int k = tid; k < d.KP; k += kPrepThreads) {
      const float c = k < d.NK ? s.coef[k] : 0.0f, hi = tf32_hi(c);
      const size_t o = ((size_t)(k >> 2) * kMTile + row) * 4 + (k & 3);
      tile[o] = hi;
      tile[(size_t)d.KP * kMTile + o] = tf32_hi(c - hi);
    }
  }
  BT(0, 2);
  // the joint outputs and a centring on a tip / the palm need skinned tip vertices before the batched product has run
  const bool center_tip = !a.trans && m.center_joint >= 0 && (m.joint_order[m.center_joint] >= NJ || (a.root_palm && m.joint_order[m.center_joint] == 0));
  if (!want_outputs) return;
  if (!a.joints && !center_tip) {
    if (tid < 3) {
      float o = 0.0f;
      if (a.trans) o = a.trans[(size_t)b * 3 + tid];
      else if (m.center_joint >= 0) o = -s.G[12 * m.joint_order[m.center_joint] + tid * 4 + 3];
      w.off[(size_t)b * 4 + tid] = o;
    }
    BT(0, 3);
    return;
  }
  // v_posed of the tip vertices and the two palm vertices (the joint outputs and the centring need them before the
  // batched product has run): one warp per (vertex, coordinate), lanes over the coefficients
  const int nsp = m.NT + (a.root_palm ? 2 : 0);
  for (int q = warp; q < 3 * nsp; q += kPrepThreads / 32) {
    const int vi = q / 3, c = q - 3 * vi;
    const int v = vi < m.NT ? m.tip_verts[vi] : m.palm_verts[vi - m.NT];
    float acc = dot_pred<5>(d.NK, lane, 32, [&](int k) { return s.coef[k] * __ldg(m.dirs + (size_t)k * m.C3 + 3 * v + c); });
    acc = warp_sum(acc);
    if (lane == 0) s_tipvp[q] = acc + __ldg(m.v_template + 3 * v + c);
  }
  __syncthreads();
  if (tid < nsp) {
    const int v = tid < m.NT ? m.tip_verts[tid] : m.palm_verts[tid - m.NT];
    float T[12];
    skin_matrix_b(m, s.A, v, T);
    const float x = s_tipvp[3 * tid], y = s_tipvp[3 * tid + 1], z = s_tipvp[3 * tid + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) s_tip[3 * tid + r] = T[r * 4 + 0] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3];
  }
  __syncthreads();
  const float* palm = s_tip + 3 * m.NT;
  auto joint_src = [&](int src, int c) -> float {
    if (src >= NJ) return s_tip[3 * (src - NJ) + c];
    if (src == 0 && a.root_palm) return (palm[c] + palm[3 + c]) / 2.0f;
    return s.G[12 * src + c * 4 + 3];
  };
  if (tid < 3) {
    float o = 0.0f;
    if (a.trans) o = a.trans[(size_t)b * 3 + tid];
    else if (m.center_joint >= 0) o = -joint_src(m.joint_order[m.center_joint], tid);
    s_off[tid] = o;
    w.off[(size_t)b * 4 + tid] = o;
  }
  __syncthreads();
  if (a.joints) {
    float* jout = a.joints + (size_t)b * (NJ + m.NT) * 3;
    for (int i = tid; i < (NJ + m.NT) * 3; i += kPrepThreads) {
      const int k = i / 3, c = i % 3;
      jout[i] = joint_src(m.joint_order[k], c) + s_off[c];
    }
  }
}

// ---------------------------------------------------------------------------------------------- forward product
// v_posed[b][c] = v_template[c] + sum_k coef[b][k] basis[k][c] for 128 samples x 32 columns per CTA.
__global__ void __launch_bounds__(kGemmThreads, 1) blend_fwd_kernel(HfrHandModel m, BatchedDims d, WsPtrs w, int B) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int KP = d.KP;
  float* sAhi = reinterpret_cast<float*>(smem_raw);
  float* sAlo = sAhi + (size_t)kMTile * KP;
  float* sB = sAlo + (size_t)kMTile * KP;                  // hi tile then lo tile, kFwdN x KP each
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)2 * kFwdN * KP);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = blockIdx.x, b0 = blockIdx.y * kMTile;
  const int rows = min(kMTile, B - b0);
  const uint32_t tile_bytes = (uint32_t)(2 * kFwdN * KP * sizeof(float));
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    // both operands are already in MMA layout in global memory (basis: hfr_mano_pack_basis; coefficients: the pose
    // stage): two bulk copies, one barrier
    const uint32_t a_bytes = (uint32_t)(2 * kMTile * KP * sizeof(float));
    mbar_expect_tx(&bars[0], tile_bytes + a_bytes);
    bulk_g2s(sB, reinterpret_cast<const float*>(m.basis_packed) + (size_t)tile * 2 * kFwdN * KP, tile_bytes, &bars[0]);
    bulk_g2s(sAhi, w.coefp + (size_t)blockIdx.y * 2 * KP * kMTile, a_bytes, &bars[0]);
  }
  if (warp == 0) tmem_alloc(tslot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    mbar_wait(&bars[0], 0);
    const uint32_t idesc = umma_idesc_tf32(kMTile, kFwdN);
    const uint32_t aH = smem_u32(sAhi), aL = smem_u32(sAlo), bH = smem_u32(sB), bL = bH + (uint32_t)(kFwdN * KP * 4);
    for (int ks = 0; ks < KP / 8; ++ks) {
      const uint32_t ao = (uint32_t)ks * 2u * kMTile * 16u, bo = (uint32_t)ks * 2u * kFwdN * 16u;
      const uint64_t dAh = umma_desc(aH + ao, kMTile * 16, 128), dAl = umma_desc(aL + ao, kMTile * 16, 128);
      const uint64_t dBh = umma_desc(bH + bo, kFwdN * 16, 128), dBl = umma_desc(bL + bo, kFwdN * 16, 128);
      umma_tf32(tmem, dAl, dBh, idesc, ks > 0);   // small terms first
      umma_tf32(tmem, dAh, dBl, idesc, 1);
      umma_tf32(tmem, dAh, dBh, idesc, 1);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    const int row = tid, c0 = tile * kFwdN;
    if (row < rows) {
      float4* out = reinterpret_cast<float4*>(w.vp + (size_t)(b0 + row) * d.C3P + c0);
      const float4* vt = reinterpret_cast<const float4*>(m.v_template + c0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + 4 * q < m.C3) t = __ldg(vt + q);
        out[q] = make_float4(v[4 * q] + t.x, v[4 * q + 1] + t.y, v[4 * q + 2] + t.z, v[4 * q + 3] + t.w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// ---------------------------------------------------------------------------------------------- skinning
__global__ void __launch_bounds__(256) hand_skin_kernel(HfrHandModel m, BatchedDims d, WsPtrs w, float* __restrict__ verts) {
  __shared__ float sA[12 * HFR_MAX_JOINTS], s_off[4];
  const int b = blockIdx.y, tid = threadIdx.x;
  const StateOff so = state_off(m.NJ);
  for (int i = tid; i < 12 * m.NJ; i += 256) sA[i] = w.state[(size_t)b * d.ST + so.A + i];
  if (tid < 3) s_off[tid] = w.off[(size_t)b * 4 + tid];
  __syncthreads();
  const int v = blockIdx.x * 256 + tid;
  if (v >= m.V) return;
  float T[12];
  skin_matrix_b(m, sA, v, T);
  const float* vp = w.vp + (size_t)b * d.C3P + 3 * v;
  const float x = vp[0], y = vp[1], z = vp[2];
  float* o = verts + ((size_t)b * m.V + v) * 3;
#pragma unroll
  for (int r = 0; r < 3; ++r) o[r] = T[r * 4 + 0] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3] + s_off[r];
}

// ---------------------------------------------------------------------------------------------- backward, stage 1
// per sample: gradient routing of the joint outputs / centring, d/d(A_j) (joint-major, no atomics) and
// g_v_posed = T^T g_v, written split into tf32 hi / lo parts in the MMA operand layout of the transposed product
__global__ void __launch_bounds__(kSkinBwdThreads) hand_skin_bwd_kernel(HfrHandModel m, HfrManoBwdArgs a, BatchedDims d, WsPtrs w,
                                                                        int jv_cap) {
  extern __shared__ __align__(16) float sm[];
  float* gv = sm;                        // C3P
  float* vp = gv + d.C3P;                // C3P
  float* gvp = vp + d.C3P;               // C3P: g_v_posed
  float* sA = gvp + d.C3P;               // 12 NJ
  float* gGt = sA + 12 * m.NJ;           // 3 NJ (padded to 4 NJ)
  float* red = gGt + 4 * m.NJ;           // 3 * warps + 8
  int* jvv = reinterpret_cast<int*>(red + 3 * (kSkinBwdThreads / 32) + 8);   // jv_cap vertex ids, then jv_cap weights
  float* jvw = reinterpret_cast<float*>(jvv + jv_cap);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSkinBwdThreads / 32;
  const int NJ = m.NJ, V = m.V, NJO = NJ + m.NT;
  const StateOff so = state_off(NJ);
  const float* gv_in = a.g_verts + (size_t)b * V * 3;
  const float* vp_in = w.vp + (size_t)b * d.C3P;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  BT(1, 0);
  for (int i0 = tid; i0 < d.C3P; i0 += 5 * kSkinBwdThreads) {   // 10 loads in flight per round (fixed unroll, predicated)
    float g[5], x[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int i = i0 + u * kSkinBwdThreads;
      g[u] = i < 3 * V ? __ldg(gv_in + i) : 0.0f;
      x[u] = i < d.C3P ? __ldg(vp_in + i) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int i = i0 + u * kSkinBwdThreads;
      if (i < d.C3P) {
        gv[i] = g[u];
        vp[i] = x[u];
        const int c = i % 3;
        if (c == 0) sx += g[u]; else if (c == 1) sy += g[u]; else sz += g[u];
      }
    }
  }
  const float* gj_in = a.g_joints ? a.g_joints + (size_t)b * NJO * 3 : nullptr;
  if (gj_in) {
    for (int i = tid; i < NJO * 3; i += kSkinBwdThreads) {
      const float g = gj_in[i];
      const int c = i % 3;
      if (c == 0) sx += g; else if (c == 1) sy += g; else sz += g;
    }
  }
  // the joint-major weight lists go to shared memory once (coalesced), so the reduction loops below never wait on
  // global memory
  const int nnz = __ldg(m.jv_ptr + NJ);
  const bool staged = nnz <= jv_cap;
  if (staged) {
    for (int i0 = tid; i0 < nnz; i0 += 6 * kSkinBwdThreads) {
      int vi[6];
      float wi[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * kSkinBwdThreads;
        vi[u] = i < nnz ? __ldg(m.jv_vert + i) : 0;
        wi[u] = i < nnz ? __ldg(m.jv_w + i) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * kSkinBwdThreads;
        if (i < nnz) { jvv[i] = vi[u]; jvw[i] = wi[u]; }
      }
    }
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if (lane == 0) { red[warp * 3] = sx; red[warp * 3 + 1] = sy; red[warp * 3 + 2] = sz; }
  for (int i = tid; i < 12 * NJ; i += kSkinBwdThreads) sA[i] = w.state[(size_t)b * d.ST + so.A + i];
  for (int i = tid; i < 4 * NJ; i += kSkinBwdThreads) gGt[i] = 0.0f;
  __syncthreads();
  BT(1, 1);
  if (tid < 3) {
    float t = 0.f;
    for (int q = 0; q < nwarps; ++q) t += red[q * 3 + tid];
    red[3 * nwarps + tid] = t;
  }
  __syncthreads();
  BT(1, 2);
  if (tid == 0) {
    auto route = [&](int src, int c, float g) {
      if (src >= NJ) {
        gv[3 * m.tip_verts[src - NJ] + c] += g;
      } else if (src == 0 && a.root_palm) {
        gv[3 * m.palm_verts[0] + c] += 0.5f * g;
        gv[3 * m.palm_verts[1] + c] += 0.5f * g;
      } else {
        gGt[3 * src + c] += g;
      }
    };
    if (gj_in) {
      for (int k = 0; k < NJO; ++k)
        for (int c = 0; c < 3; ++c) route(m.joint_order[k], c, gj_in[3 * k + c]);
    }
    if (a.trans) {
      if (a.g_trans) for (int c = 0; c < 3; ++c) a.g_trans[(size_t)b * 3 + c] = red[3 * nwarps + c];
    } else if (m.center_joint >= 0) {
      const int src = m.joint_order[m.center_joint];
      for (int c = 0; c < 3; ++c) route(src, c, -red[3 * nwarps + c]);
    }
  }
  __syncthreads();
  BT(1, 3);
  for (int i = tid; i < 3 * NJ; i += kSkinBwdThreads) w.gGt[(size_t)b * 4 * NJ + i] = gGt[i];
  if (warp < nwarps / 2) {
    // d/d(A_j) = sum_v w_vj g_v (x) [v_posed; 1]: 16 lanes per joint walk its weight list, each lane keeps all 12 entries
    // of the 3x4 block in registers (8 shared-memory reads per list element), joined by a fixed shuffle tree
    for (int j0 = 0; j0 < NJ; j0 += kSkinBwdThreads / 32) {
      const int j = j0 + (tid >> 4), q = tid & 15;
      float acc[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) acc[e] = 0.0f;
      if (j < NJ) {
        const int p0 = __ldg(m.jv_ptr + j), p1 = __ldg(m.jv_ptr + j + 1);
#pragma unroll 2
        for (int p = p0 + q; p < p1; p += 16) {
          const int v = staged ? jvv[p] : __ldg(m.jv_vert + p);
          const float wt = staged ? jvw[p] : __ldg(m.jv_w + p);
          const float x = vp[3 * v], y = vp[3 * v + 1], z = vp[3 * v + 2];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float wg = wt * gv[3 * v + r];
            acc[4 * r + 0] += wg * x; acc[4 * r + 1] += wg * y; acc[4 * r + 2] += wg * z; acc[4 * r + 3] += wg;
          }
        }
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
#pragma unroll
        for (int e = 0; e < 12; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
      }
      if (j < NJ && q == 0) {
#pragma unroll
        for (int e = 0; e < 12; ++e) w.gA[(size_t)b * 12 * NJ + 12 * j + e] = acc[e];
      }
    }
  } else {
    // g_v_posed = T_rot^T g_v (zero padded up to the pitch)
    const int t2 = tid - kSkinBwdThreads / 2, T2 = kSkinBwdThreads / 2;
    for (int v = t2; v < V; v += T2) {
      float T[12];
      skin_matrix_b(m, sA, v, T);
      const float g0 = gv[3 * v], g1 = gv[3 * v + 1], g2 = gv[3 * v + 2];
      gvp[3 * v + 0] = T[0] * g0 + T[4] * g1 + T[8] * g2;
      gvp[3 * v + 1] = T[1] * g0 + T[5] * g1 + T[9] * g2;
      gvp[3 * v + 2] = T[2] * g0 + T[6] * g1 + T[10] * g2;
    }
    for (int i = 3 * V + t2; i < d.C3P; i += T2) gvp[i] = 0.0f;
  }
  __syncthreads();
  BT(1, 4);
  {   // operand layout of the transposed product: per (128-sample tile, 64-column chunk) hi | lo, each [16][128][4]
    float4* tile = reinterpret_cast<float4*>(w.gvpp) + (size_t)(b >> 7) * d.NCH64 * 2 * (kBwdK / 4) * kMTile;
    const int row = b & (kMTile - 1);
    for (int c4 = tid; c4 < d.C3P / 4; c4 += kSkinBwdThreads) {
      float4 hi, lo;
      split4(reinterpret_cast<const float4*>(gvp)[c4], hi, lo);
      const int q = c4 >> 4, kc = c4 & 15;
      float4* dst = tile + ((size_t)q * 2 * (kBwdK / 4) + kc) * kMTile + row;
      dst[0] = hi;
      dst[(size_t)(kBwdK / 4) * kMTile] = lo;
    }
  }
  BT(1, 5);
}

// ---------------------------------------------------------------------------------------------- backward product
// g_coef partial[ks][b][k] = sum over this CTA's 64-column chunks of g_v_posed[b][c] basis[k][c]
__global__ void __launch_bounds__(kGemmThreads, 1) blend_bwd_kernel(HfrHandModel m, BatchedDims d, WsPtrs w, int B, size_t packed_bwd_off) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int N = d.NKP16;
  float* sAhi = reinterpret_cast<float*>(smem_raw);       // 128 x 64
  float* sAlo = sAhi + kMTile * kBwdK;
  float* sB = sAlo + kMTile * kBwdK;                       // hi then lo, N x 64 each
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)2 * N * kBwdK);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ks = blockIdx.x, b0 = blockIdx.y * kMTile;
  const int rows = min(kMTile, B - b0);
  const int per = (d.NCH64 + w.ksplit - 1) / w.ksplit;
  const int q0 = ks * per, q1 = min(q0 + per, d.NCH64);
  const uint32_t tile_bytes = (uint32_t)(2 * N * kBwdK * sizeof(float));
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tslot, 256);
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t idesc = umma_idesc_tf32(kMTile, N);
  uint32_t phase = 0;
  for (int q = q0; q < q1; ++q) {
    if (tid == 0) {   // both operand tiles of this chunk are in MMA layout in global memory: two bulk copies
      const uint32_t a_bytes = (uint32_t)(2 * kMTile * kBwdK * sizeof(float));
      mbar_expect_tx(&bars[0], tile_bytes + a_bytes);
      bulk_g2s(sB, reinterpret_cast<const float*>(m.basis_packed) + packed_bwd_off + (size_t)q * 2 * N * kBwdK, tile_bytes, &bars[0]);
      bulk_g2s(sAhi, w.gvpp + ((size_t)blockIdx.y * d.NCH64 + q) * 2 * kBwdK * kMTile, a_bytes, &bars[0]);
    }
    if (tid == 0) {
      mbar_wait(&bars[0], phase);
      const uint32_t aH = smem_u32(sAhi), aL = smem_u32(sAlo), bH = smem_u32(sB), bL = bH + (uint32_t)(N * kBwdK * 4);
      for (int k8 = 0; k8 < kBwdK / 8; ++k8) {
        const uint32_t ao = (uint32_t)k8 * 2u * kMTile * 16u, bo = (uint32_t)k8 * 2u * (uint32_t)N * 16u;
        const uint64_t dAh = umma_desc(aH + ao, kMTile * 16, 128), dAl = umma_desc(aL + ao, kMTile * 16, 128);
        const uint64_t dBh = umma_desc(bH + bo, N * 16, 128), dBl = umma_desc(bL + bo, N * 16, 128);
        umma_tf32(tmem, dAl, dBh, idesc, (q > q0 || k8 > 0) ? 1u : 0u);
        umma_tf32(tmem, dAh, dBl, idesc, 1);
        umma_tf32(tmem, dAh, dBh, idesc, 1);
      }
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], phase);   // the operands may be overwritten once the MMAs have retired
    tc_fence_after();
    phase ^= 1;
  }
  {
    const int row = tid;
    float* out = w.gcoefp + ((size_t)ks * B + (b0 + row)) * N;
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      if (q1 > q0) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.0f;
      }
      if (row < rows) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + 4 * i < N) reinterpret_cast<float4*>(out + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------- backward, stage 3
// per sample: fixed-order sum of the split-K partials, kinematic-chain backward one tree level at a time with a thread
// per output element (children write private slots, parents add them in index order: no atomics, reproducible),
// Rodrigues', shape' and pose-PCA'
__global__ void __launch_bounds__(kChainThreads) hand_chain_bwd_kernel(HfrHandModel m, HfrManoBwdArgs a, int pose_dim, BatchedDims d,
                                                                       WsPtrs w, const int* __restrict__ depth_tab) {
  __shared__ float s_full[3 * HFR_MAX_JOINTS], s_R[9 * HFR_MAX_JOINTS], s_J[3 * HFR_MAX_JOINTS], s_G[12 * HFR_MAX_JOINTS];
  __shared__ float gG[12 * HFR_MAX_JOINTS], gJ[3 * HFR_MAX_JOINTS], gR[9 * HFR_MAX_JOINTS], gGt[3 * HFR_MAX_JOINTS], gAt[3 * HFR_MAX_JOINTS];
  __shared__ float gcoef[64 + 9 * HFR_MAX_JOINTS], gfull[3 * HFR_MAX_JOINTS], contrib[15 * HFR_MAX_JOINTS];
  __shared__ int s_depth[HFR_MAX_JOINTS], s_parent[HFR_MAX_JOINTS];
  const int b = blockIdx.x, tid = threadIdx.x, NJ = m.NJ, NK = d.NK, NPOSE = 3 * (NJ - 1);
  const int B = a.B, poff = a.pose_off > 0 ? a.pose_off : 3, n_rot = a.rots ? a.n_rot_in : 0;
  const StateOff so = state_off(NJ);
  const float* st = w.state + (size_t)b * d.ST;
  BT(2, 0);
  for (int i = tid; i < 3 * NJ; i += kChainThreads) {
    s_full[i] = st[so.full + i]; s_J[i] = st[so.J + i]; gGt[i] = w.gGt[(size_t)b * 4 * NJ + i];
    gAt[i] = w.gA[(size_t)b * 12 * NJ + 12 * (i / 3) + 4 * (i % 3) + 3];
  }
  for (int i = tid; i < 9 * NJ; i += kChainThreads) { s_R[i] = st[so.R + i]; gR[i] = 0.0f; }
  for (int i = tid; i < 12 * NJ; i += kChainThreads) { s_G[i] = st[so.G + i]; gG[i] = w.gA[(size_t)b * 12 * NJ + i]; }
  for (int j = tid; j < NJ; j += kChainThreads) { s_depth[j] = depth_tab[j]; s_parent[j] = m.parents[j]; }
  // the split-K partials of the transposed product, summed in chunk order (reproducible); 16 loads in flight per thread
  for (int k = tid; k < NK; k += kChainThreads) {
    float acc = 0.0f;
    for (int ks0 = 0; ks0 < w.ksplit; ks0 += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = ks0 + u < w.ksplit ? __ldg(w.gcoefp + ((size_t)(ks0 + u) * B + b) * d.NKP16 + k) : 0.0f;
#pragma unroll
      for (int u = 0; u < 16; ++u) acc += v[u];
    }
    gcoef[k] = acc;
  }
  __syncthreads();
  BT(2, 1);
  // A_j = [G.R | G.t - G.R J]  =>  gG.R = gA.R - gA.t (x) J,  gG.t = gA.t + direct joint gradients,  gJ = -G.R^T gA.t
  for (int idx = tid; idx < 15 * NJ; idx += kChainThreads) {
    if (idx < 3 * NJ) {
      const int j = idx / 3, k = idx - 3 * j;
      const float* G = s_G + 12 * j;
      gJ[idx] = -(G[k] * gAt[3 * j] + G[4 + k] * gAt[3 * j + 1] + G[8 + k] * gAt[3 * j + 2]);
    } else {
      const int e = idx - 3 * NJ, j = e / 12, q = e - 12 * j, r = q >> 2, k = q & 3;
      if (k < 3) gG[12 * j + q] -= gAt[3 * j + r] * s_J[3 * j + k];
      else gG[12 * j + q] = gAt[3 * j + r] + gGt[3 * j + r];
    }
  }
  int maxd = 0;
  for (int i = 0; i < NJ; ++i) maxd = max(maxd, s_depth[i]);
  __syncthreads();
  BT(2, 2);
  for (int dl = maxd; dl >= 1; --dl) {
    // children of this level: differentiate G_j = G_p o [R_j | J_j - J_p], 24 outputs per joint
    for (int idx = tid; idx < 24 * NJ; idx += kChainThreads) {
      const int j = idx / 24, o = idx - 24 * j;
      if (s_depth[j] != dl) continue;
      const int p = s_parent[j];
      const float* P = s_G + 12 * p;
      const float* gGj = gG + 12 * j;
      if (o < 9) {
        const int k = o / 3, c = o - 3 * k;
        gR[9 * j + o] = P[k] * gGj[c] + P[4 + k] * gGj[4 + c] + P[8 + k] * gGj[8 + c];
      } else if (o < 12) {
        const int k = o - 9;
        const float gtl = P[k] * gGj[3] + P[4 + k] * gGj[7] + P[8 + k] * gGj[11];
        contrib[15 * j + 12 + k] = gtl;
        gJ[3 * j + k] += gtl;
      } else {
        const int e = o - 12, r = e >> 2, k = e & 3;
        float val = gGj[4 * r + 3];
        if (k < 3) {
          const float* Rl = s_R + 9 * j;
          val = gGj[4 * r] * Rl[3 * k] + gGj[4 * r + 1] * Rl[3 * k + 1] + gGj[4 * r + 2] * Rl[3 * k + 2] +
                gGj[4 * r + 3] * (s_J[3 * j + k] - s_J[3 * p + k]);
        }
        contrib[15 * j + e] = val;
      }
    }
    __syncthreads();
    // their parents gather, children in index order
    for (int idx = tid; idx < 15 * NJ; idx += kChainThreads) {
      const int j = idx / 15, e = idx - 15 * j;
      if (s_depth[j] != dl - 1) continue;
      float acc = 0.0f;
      for (int ch = 0; ch < NJ; ++ch)
        if (s_parent[ch] == j) acc += contrib[15 * ch + e];
      if (e < 12) gG[12 * j + e] += acc;
      else gJ[3 * j + e - 12] -= acc;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < 12 * NJ; idx += kChainThreads) {
    const int j = idx / 12, e = idx - 12 * j, r = e >> 2, c = e & 3;
    if (s_depth[j] != 0) continue;
    if (c < 3) gR[9 * j + 3 * r + c] = gG[idx];
    else gJ[3 * j + r] += gG[idx];
  }
  __syncthreads();
  BT(2, 3);
  for (int j = tid; j < NJ; j += kChainThreads) {
    float g[9];
    for (int e = 0; e < 9; ++e) g[e] = gR[9 * j + e] + (j >= 1 ? gcoef[m.NS + 9 * (j - 1) + e] : 0.0f);
    float gvv[3] = {0.f, 0.f, 0.f};
    if (j < n_rot) {
      if (a.g_rots) for (int e = 0; e < 9; ++e) a.g_rots[((size_t)b * n_rot + j) * 9 + e] = g[e];
    } else {
      hfr_rodrigues_bwd(s_full + 3 * j, g, gvv);
    }
    gfull[3 * j] = gvv[0]; gfull[3 * j + 1] = gvv[1]; gfull[3 * j + 2] = gvv[2];
  }
  if (a.g_betas && a.betas && tid >= 64) {   // (the first warps are busy with Rodrigues')
    for (int k = tid - 64; k < m.NS; k += kChainThreads - 64) {
      a.g_betas[(size_t)b * m.NS + k] = gcoef[k] + dot_pred<16>(3 * NJ, 0, 1, [&](int i) { return __ldg(m.J_shapedirs + i * m.NS + k) * gJ[i]; });
    }
  }
  __syncthreads();
  BT(2, 4);
  if (!a.g_pose || n_rot >= NJ) return;
  float* gp = a.g_pose + (size_t)b * pose_dim;
  if (tid < poff) gp[tid] = (n_rot == 0 && tid < 3) ? gfull[tid] : 0.0f;
  if (m.NPC > 0) {
    // 4 lanes per PCA coefficient
    for (int base = 0; base < 4 * m.NPC; base += kChainThreads) {
      const int idx = base + tid, k = idx >> 2, part = idx & 3;
      float acc = 0.0f;
      if (k < m.NPC) acc = dot_pred<12>(NPOSE, part, 4, [&](int o) { return __ldg(m.pca_comps + k * NPOSE + o) * gfull[3 + o]; });
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (k < m.NPC && part == 0) gp[poff + k] = acc;
    }
  } else {
    for (int o = tid; o < NPOSE; o += kChainThreads) gp[poff + o] = gfull[3 + o];
  }
  BT(2, 5);
}

// ---------------------------------------------------------------------------------------------- basis packing
// forward section : NT32 tiles x {hi, lo} x [KP/4 chunks][32 columns][4 coefficients]   value = basis[k][32 t + n]
// backward section: NCH64 chunks x {hi, lo} x [16 chunks][NKP16 coefficients][4 columns] value = basis[n][64 q + c]
__global__ void pack_basis_kernel(HfrHandModel m, BatchedDims d, float* __restrict__ out, size_t fwd_floats, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)m.NJ) {   // depth of joint i in the kinematic tree (root = 0)
    int dd = 0;
    for (int p = m.parents[i]; p >= 0; p = m.parents[p]) ++dd;
    reinterpret_cast<int*>(out + total)[i] = dd;
  }
  if (i >= total) return;
  float x = 0.0f;
  bool lo;
  if (i < fwd_floats) {
    const size_t per = (size_t)kFwdN * d.KP;
    const int t = (int)(i / (2 * per));
    size_t r = i - (size_t)t * 2 * per;
    lo = r >= per;
    if (lo) r -= per;
    const int e = (int)(r & 3), n = (int)((r >> 2) % kFwdN), kc = (int)(r / (4 * kFwdN));
    const int k = 4 * kc + e, c = kFwdN * t + n;
    if (k < d.NK && c < m.C3) x = m.dirs[(size_t)k * m.C3 + c];
  } else {
    const size_t j = i - fwd_floats, per = (size_t)d.NKP16 * kBwdK;
    const int q = (int)(j / (2 * per));
    size_t r = j - (size_t)q * 2 * per;
    lo = r >= per;
    if (lo) r -= per;
    const int e = (int)(r & 3), n = (int)((r >> 2) % d.NKP16), kc = (int)(r / (4 * (size_t)d.NKP16));
    const int c = kBwdK * q + 4 * kc + e;
    if (n < d.NK && c < m.C3) x = m.dirs[(size_t)n * m.C3 + c];
  }
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  out[i] = lo ? __uint_as_float(__float_as_uint(x - hi) & 0xFFFFE000u) : hi;
}

size_t packed_fwd_floats(const BatchedDims& d) { return (size_t)d.NT32 * 2 * kFwdN * d.KP; }
size_t packed_bwd_floats(const BatchedDims& d) { return (size_t)d.NCH64 * 2 * d.NKP16 * kBwdK; }
constexpr int kDepthWords = HFR_MAX_JOINTS;   // tree depth of every joint, int32, stored behind the two basis sections
size_t fwd_smem_bytes(const BatchedDims& d) { return (size_t)(2 * kMTile + 2 * kFwdN) * d.KP * 4 + 64; }
size_t bwd_smem_bytes(const BatchedDims& d) { return (size_t)(2 * kMTile + 2 * d.NKP16) * kBwdK * 4 + 64; }
int skin_bwd_jv_cap(const HfrHandModel& m, const BatchedDims& d) {   // joint-major weight lists staged in shared memory when they fit
  const size_t fixed = (size_t)(3 * d.C3P + 16 * m.NJ + 3 * (kSkinBwdThreads / 32) + 8) * 4;
  const size_t want = (size_t)m.NW * m.V;
  return fixed + want * 8 <= 200 * 1024 ? (int)want : 0;
}
size_t skin_bwd_smem_bytes(const HfrHandModel& m, const BatchedDims& d) {
  return (size_t)(3 * d.C3P + 16 * m.NJ + 3 * (kSkinBwdThreads / 32) + 8) * 4 + (size_t)skin_bwd_jv_cap(m, d) * 8;
}

}  // namespace

bool mano_batched_ok(const HfrHandModel* m, int B, const void* workspace) {
  if (!m->basis_packed || !workspace || !m->jv_ptr || B < 1) return false;
  const BatchedDims d = batched_dims(*m);
  if (m->NT + 2 > 24 || d.NKP16 > 256 || m->NS > 64) return false;
  return fwd_smem_bytes(d) <= 227 * 1024 && bwd_smem_bytes(d) <= 227 * 1024 && skin_bwd_smem_bytes(*m, d) <= 227 * 1024;
}

static int launch_forward_stages(const HfrHandModel* m, const HfrManoFwdArgs& fa, int pose_dim, const BatchedDims& d, const WsPtrs& w,
                                 int want_outputs, cudaStream_t st) {
  const int* depth_tab = reinterpret_cast<const int*>(reinterpret_cast<const float*>(m->basis_packed) + packed_fwd_floats(d) + packed_bwd_floats(d));
  hand_prep_kernel<<<fa.B, kPrepThreads, 0, st>>>(*m, fa, pose_dim, d, w, want_outputs, depth_tab);
  const size_t smem = fwd_smem_bytes(d);
  cudaFuncSetAttribute(blend_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  blend_fwd_kernel<<<dim3(d.NT32, (fa.B + kMTile - 1) / kMTile), kGemmThreads, smem, st>>>(*m, d, w, fa.B);
  return HFR_OK;
}

int mano_batched_forward(const HfrHandModel* m, const HfrManoFwdArgs* a, int pose_dim, cudaStream_t st) {
  const BatchedDims d = batched_dims(*m);
  const WsPtrs w = carve_ws(*m, d, a->B, a->workspace);
  launch_forward_stages(m, *a, pose_dim, d, w, 1, st);
  hand_skin_kernel<<<dim3((m->V + 255) / 256, a->B), 256, 0, st>>>(*m, d, w, a->verts);
  HFR_CHECK_LAUNCH("mano_forward (batched)");
  return HFR_OK;
}

int mano_batched_backward(const HfrHandModel* m, const HfrManoBwdArgs* a, int pose_dim, cudaStream_t st) {
  const BatchedDims d = batched_dims(*m);
  const WsPtrs w = carve_ws(*m, d, a->B, a->workspace);
  if (!a->reuse_forward) {   // rebuild the pose state and v_posed from the inputs
    HfrManoFwdArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.B = a->B; fa.pose = a->pose; fa.betas = a->betas; fa.trans = a->trans; fa.rots = a->rots;
    fa.n_rot_in = a->n_rot_in; fa.pose_off = a->pose_off; fa.root_palm = a->root_palm; fa.workspace = a->workspace;
    launch_forward_stages(m, fa, pose_dim, d, w, 0, st);
  }
  const size_t s1 = skin_bwd_smem_bytes(*m, d);
  if (s1 > 48 * 1024) cudaFuncSetAttribute(hand_skin_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
  hand_skin_bwd_kernel<<<a->B, kSkinBwdThreads, s1, st>>>(*m, *a, d, w, skin_bwd_jv_cap(*m, d));
  const size_t s2 = bwd_smem_bytes(d);
  cudaFuncSetAttribute(blend_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2);
  blend_bwd_kernel<<<dim3(w.ksplit, (a->B + kMTile - 1) / kMTile), kGemmThreads, s2, st>>>(*m, d, w, a->B, packed_fwd_floats(d));
  const int* depth_tab = reinterpret_cast<const int*>(reinterpret_cast<const float*>(m->basis_packed) + packed_fwd_floats(d) + packed_bwd_floats(d));
  hand_chain_bwd_kernel<<<a->B, kChainThreads, 0, st>>>(*m, *a, pose_dim, d, w, depth_tab);
  HFR_CHECK_LAUNCH("mano_backward (batched)");
  return HFR_OK;
}

}  // namespace hfr

extern "C" int64_t hfr_mano_packed_basis_bytes(const HfrHandModel* m) {
  if (!m || m->NJ < 1 || m->C3 <= 0) return 0;
  const hfr::BatchedDims d = hfr::batched_dims(*m);
  return (int64_t)((hfr::packed_fwd_floats(d) + hfr::packed_bwd_floats(d) + hfr::kDepthWords) * sizeof(float));
}

extern "C" int hfr_mano_pack_basis(const HfrHandModel* m, void* packed, void* stream) {
  HFR_CHECK_ARG(m && m->dirs && packed, "mano_pack_basis: null pointer");
  const hfr::BatchedDims d = hfr::batched_dims(*m);
  const size_t fwd = hfr::packed_fwd_floats(d), total = fwd + hfr::packed_bwd_floats(d);
  hfr::pack_basis_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*m, d, reinterpret_cast<float*>(packed), fwd, total);
  HFR_CHECK_LAUNCH("mano_pack_basis");
  return HFR_OK;
}

extern "C" int64_t hfr_mano_workspace_bytes(const HfrHandModel* m, int32_t B) {
  if (!m || B < 1) return 0;
  const hfr::BatchedDims d = hfr::batched_dims(*m);
  return (int64_t)(hfr::carve_ws(*m, d, B, nullptr).floats * sizeof(float));
}

// 0 = no mbarrier wait of the batched kernels ever timed out on the current device (debug / tests)
#ifdef HFR_MANO_TIMING
extern "C" int hfr_debug_batched_times(long long* out) {
  return cudaMemcpyFromSymbol(out, hfr::g_bt, sizeof(long long) * 48) == cudaSuccess ? 0 : 1;
}
#endif
extern "C" int hfr_mano_batched_status(void) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, hfr::g_batched_timeout, sizeof(v)) != cudaSuccess) return -1;
  return (int)v;
}

// Batched hand-layer path (declarations shared by mano.cu and mano_batched.cu).
//
// The per-sample kernels of mano.cu stream the whole blend basis once per sample.  The batched path splits the layer
// into a per-sample pose stage, ONE blend contraction for the whole batch on the tensor cores
//   v_posed[B x 3V] = [betas | R(pose) - I][B x NK] . basis[NK x 3V]          (utils/my_mano.py:386-393)
// (tcgen05.mma kind::tf32 with a 3-term hi/lo split, fp32 accumulators in TMEM; the basis tiles are pre-packed in the
// UMMA shared-memory layout and arrive by cp.async.bulk), and a vertex-parallel skinning stage; the backward mirrors
// it with the transposed product g_coef[B x NK] = g_v_posed[B x 3V] . basis^T split over the 3V dimension.
#pragma once
#include "common.cuh"

namespace hfr {

struct BatchedDims {
  int NK;      // blend coefficients: NS + 9 (NJ - 1)
  int KP;      // NK rounded up to the MMA K granule (8 tf32)
  int NKP16;   // NK rounded up to 16 (N of the backward product)
  int C3P;     // 3V rounded up to 64 (row pitch of v_posed / its gradient in the workspace)
  int NT32;    // 32-column tiles of the forward product
  int NCH64;   // 64-column chunks of the backward product's reduction dimension
  int ST;      // per-sample pose state: full pose 3NJ, R 9NJ, J 3NJ, G 12NJ, A 12NJ (padded to a multiple of 4)
};
__host__ __device__ inline BatchedDims batched_dims(const HfrHandModel& m) {
  BatchedDims d;
  d.NK = m.NS + 9 * (m.NJ - 1);
  d.KP = (d.NK + 7) & ~7;
  d.NKP16 = (d.NK + 15) & ~15;
  d.C3P = (m.C3 + 63) & ~63;
  d.NT32 = (m.C3 + 31) / 32;
  d.NCH64 = d.C3P / 64;
  d.ST = 39 * m.NJ + ((4 - (39 * m.NJ) % 4) % 4);
  return d;
}

// true when the batched path can serve this model / call (else the per-sample kernels run)
bool mano_batched_ok(const HfrHandModel* m, int B, const void* workspace);
int mano_batched_forward(const HfrHandModel* m, const HfrManoFwdArgs* a, int pose_dim, cudaStream_t st);
int mano_batched_backward(const HfrHandModel* m, const HfrManoBwdArgs* a, int pose_dim, cudaStream_t st);

}  // namespace hfr

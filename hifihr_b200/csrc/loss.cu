// SSAA pooling / output split and the render-dependent losses for sm_100a, forward + backward.
//
//   pool : models_res_nimble.py:210-220 (NHWC->NCHW, avg_pool2d(aa), RGB/alpha split, binarise, maskRGBs)
//   loss : losses.py:355-378 (texture L1, mean-RGB, SSIM), :399-408 (silhouette L1, IoU) with
//          utils/losses_util.py:366-378 and utils/pytorch_ssim/__init__.py:17-37.
//
// SSIM is a separable 11x11 Gaussian stencil evaluated per 32x32 tile out of shared memory
// (zero padding as F.conv2d(padding=5)); its backward is the same stencil applied to three
// per-pixel derivative maps, so the whole photometric loss costs two passes over the image.
#include "common.cuh"

namespace hfr {

// ------------------------------------------------------------------------------------- pooling
__global__ void __launch_bounds__(256) pool_fwd_kernel(HfrPoolArgs a) {
  const size_t total = (size_t)a.N * a.H * a.W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % a.W), y = (int)((i / a.W) % a.H), n = (int)(i / ((size_t)a.W * a.H));
  const int aa = a.aa, Wi = a.W * aa, Hi = a.H * aa;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int dy = 0; dy < aa; ++dy)
    for (int dx = 0; dx < aa; ++dx) {
      const float4 v = *reinterpret_cast<const float4*>(a.image + (((size_t)n * Hi + y * aa + dy) * Wi + x * aa + dx) * 4);
      s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    }
  const float inv = (float)(aa * aa);
  s0 /= inv; s1 /= inv; s2 /= inv; s3 /= inv;
  const size_t hw = (size_t)a.H * a.W, p = (size_t)y * a.W + x;
  a.re_img[((size_t)n * 3 + 0) * hw + p] = s0;
  a.re_img[((size_t)n * 3 + 1) * hw + p] = s1;
  a.re_img[((size_t)n * 3 + 2) * hw + p] = s2;
  const float sil = (a.binarize && s3 > 0.0f) ? 255.0f : s3;
  a.re_sil[(size_t)n * hw + p] = sil;
  if (a.mask_rgbs && a.images_in) {
    const float m = sil > 0.0f ? 1.0f : 0.0f;
    for (int c = 0; c < 3; ++c) a.mask_rgbs[((size_t)n * 3 + c) * hw + p] = a.images_in[((size_t)n * 3 + c) * hw + p] * m;
  }
}

__global__ void __launch_bounds__(256) pool_bwd_kernel(HfrPoolBwdArgs a) {
  const int aa = a.aa, Wi = a.W * aa, Hi = a.H * aa;
  const size_t total = (size_t)a.N * Hi * Wi;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xi = (int)(i % Wi), yi = (int)((i / Wi) % Hi), n = (int)(i / ((size_t)Wi * Hi));
  const int x = xi / aa, y = yi / aa;
  const size_t hw = (size_t)a.H * a.W, p = (size_t)y * a.W + x;
  const float inv = 1.0f / (float)(aa * aa);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.g_re_img) {
    g.x = a.g_re_img[((size_t)n * 3 + 0) * hw + p] * inv;
    g.y = a.g_re_img[((size_t)n * 3 + 1) * hw + p] * inv;
    g.z = a.g_re_img[((size_t)n * 3 + 2) * hw + p] * inv;
  }
  if (a.g_re_sil && !a.binarize) g.w = a.g_re_sil[(size_t)n * hw + p] * inv;
  *reinterpret_cast<float4*>(a.g_image + i * 4) = g;
}

// ------------------------------------------------------------------------------------- losses
// One CTA = one 32x32 pixel tile of one sample, one colour channel at a time.  The separable
// 11-tap Gaussian runs out of shared memory with register strips: in the horizontal pass a thread
// owns 8 consecutive outputs of one halo row (18 inputs read once as 128-bit LDS, 5 moment maps x
// 8 x 11 FMAs), in the vertical pass 4 consecutive rows of one column (14 inputs per map), so the
// stencil is FMA-bound instead of LDS-bound.
constexpr int kLT = 32, kR = 5, kLH = kLT + 2 * kR;     // tile, radius, halo (42)
constexpr int kXP = 44;                                  // pitch of the halo arrays (float4-aligned strips)
constexpr int kHP = 36;                                  // pitch of the horizontally blurred maps
constexpr int kLossThreads = 256;
constexpr int kHTasks = kLH * (kLT / 8);                 // 168 (row, strip-of-8) tasks

// rendered colour / silhouette of pixel p of sample n in either layout
__device__ __forceinline__ float ld_rgb(const HfrLossArgs& a, int n, int c, size_t p, size_t hw) {
  return a.nhwc ? __ldg(a.re_img + ((size_t)n * hw + p) * 4 + c) : __ldg(a.re_img + ((size_t)n * 3 + c) * hw + p);
}
__device__ __forceinline__ float ld_sil(const HfrLossArgs& a, int n, size_t p, size_t hw) {
  return a.nhwc ? __ldg(a.re_img + ((size_t)n * hw + p) * 4 + 3) : __ldg(a.re_sil + (size_t)n * hw + p);
}

// target colour / mask of pixel p of sample n, float or 8-bit transport (lut[k] = k / 255.0f, IEEE division)
__device__ __forceinline__ float ld_target(const HfrLossArgs& a, const float* lut, int n, int c, size_t p, size_t hw) {
  const size_t i = ((size_t)n * 3 + c) * hw + p;
  return a.imgs_u8 ? lut[__ldg(a.imgs_u8 + i)] : __ldg(a.imgs + i);
}
__device__ __forceinline__ float ld_seg(const HfrLossArgs& a, int n, size_t p, size_t hw) {
  const size_t i = (size_t)n * hw + p;
  return a.seg_u8 ? (__ldg(a.seg_u8 + i) != 0 ? 1.0f : 0.0f) : __ldg(a.seg + i);
}

// 8 outputs of an 11-tap row filter from 18 inputs, `NM` maps at once: out[o] += g[j-o] * in[j]
template <int NM>
__device__ __forceinline__ void tap_accumulate(float (&acc)[8][NM], const float (&val)[NM], int j, const float (&g)[11]) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const int t = j - o;
    if (t >= 0 && t < 11) {
#pragma unroll
      for (int m = 0; m < NM; ++m) acc[o][m] = fmaf(g[t], val[m], acc[o][m]);
    }
  }
}

// SSIM of one pixel from its five windowed moments, and its derivatives wrt (mu_x, E[x^2], E[xy])
// (utils/pytorch_ssim/__init__.py:17-37).  One copy, so that the all-zero-tile shortcut below yields
// the same bits as the full stencil.
__device__ __forceinline__ float ssim_point(float m1, float m2, float e11, float e22, float e12, float* dm1, float* de11,
                                            float* de12) {
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
  const float s1 = e11 - m11, s2 = e22 - m22, s12 = e12 - m12;
  const float A1 = 2.f * m12 + C1, A2 = 2.f * s12 + C2, B1 = m11 + m22 + C1, B2 = s1 + s2 + C2;
  const float ib1 = 1.0f / B1, ib2 = 1.0f / B2;
  const float S = (A1 * A2) * (ib1 * ib2);
  const float ib = ib1 * ib2;
  *dm1 = 2.f * m2 * (A2 - A1) * ib + 2.f * m1 * S * (ib2 - ib1);
  *de11 = -S * ib2;
  *de12 = 2.f * A1 * ib;
  return S;
}

constexpr size_t kLossFwdSmem = (size_t)(6 * kLH * kXP + 5 * kLH * kHP) * sizeof(float);

// One colour channel of the SSIM stencil over the CTA's 32x32 tile (halo arrays in shared memory): separable 11-tap
// Gaussian of the moment maps, SSIM value summed over this thread's 4 pixels, derivative maps written when `want_d`.
// XZERO: x vanishes on the whole halo, so mu_x = E[x^2] = E[xy] = +0 exactly and only (y, y^2) are filtered.
template <bool XZERO>
__device__ __forceinline__ float ssim_channel(const HfrLossArgs& a, float (*xs)[kXP], float (*ys)[kXP], float (*hb)[kLH][kHP],
                                              const float (&g)[11], int n, int c, int x0, int y0, size_t hw, bool want_d) {
  constexpr int NM = XZERO ? 2 : 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- horizontal pass: (halo row, strip of 8 outputs) per thread ------------------------------
  if (tid < kHTasks) {
    const int row = tid >> 2, strip = tid & 3;
    float acc[8][NM];
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int m = 0; m < NM; ++m) acc[o][m] = 0.f;
    const float4* xr = reinterpret_cast<const float4*>(&xs[row][strip * 8]);
    const float4* yr = reinterpret_cast<const float4*>(&ys[row][strip * 8]);
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float4 yv = yr[q];
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!XZERO) xv = xr[q];
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = 4 * q + e;
        if (j < 18) {
          float val[NM];
          if (XZERO) { val[0] = ya[e]; val[1] = ya[e] * ya[e]; }
          else { val[0] = xa[e]; val[1] = ya[e]; val[2] = xa[e] * xa[e]; val[NM - 2] = ya[e] * ya[e]; val[NM - 1] = xa[e] * ya[e]; }
          tap_accumulate<NM>(acc, val, j, g);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      float4* dst = reinterpret_cast<float4*>(&hb[m][row][strip * 8]);
      dst[0] = make_float4(acc[0][m], acc[1][m], acc[2][m], acc[3][m]);
      dst[1] = make_float4(acc[4][m], acc[5][m], acc[6][m], acc[7][m]);
    }
  }
  __syncthreads();
  // ---- vertical pass: (column, strip of 4 rows) per thread, then the SSIM map --------------------
  const int x = lane, rs = warp;       // 8 warps x 4 rows = 32 rows
  float acc[4][NM];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[o][m] = 0.f;
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    float val[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) val[m] = hb[m][rs * 4 + j][x];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int t = j - o;
      if (t >= 0 && t < 11) {
#pragma unroll
        for (int m = 0; m < NM; ++m) acc[o][m] = fmaf(g[t], val[m], acc[o][m]);
      }
    }
  }
  const int gx = x0 + x;
  float ss = 0.f;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int gy = y0 + rs * 4 + o;
    if (gx < a.W && gy < a.H) {
      float dm1, de11, de12;
      if (XZERO) ss += ssim_point(0.f, acc[o][0], 0.f, acc[o][NM - 1], 0.f, &dm1, &de11, &de12);
      else ss += ssim_point(acc[o][0], acc[o][1], acc[o][2], acc[o][NM - 2], acc[o][NM - 1], &dm1, &de11, &de12);
      if (want_d) {
        const size_t p = (size_t)gy * a.W + gx;
        a.dmaps[((size_t)n * 9 + c * 3 + 0) * hw + p] = dm1;
        a.dmaps[((size_t)n * 9 + c * 3 + 1) * hw + p] = de11;
        a.dmaps[((size_t)n * 9 + c * 3 + 2) * hw + p] = de12;
      }
    }
  }
  return ss;
}

// METRIC = the evaluation-time texture metrics (train_hrnet.py:149-161, compute_texture_metric.py:49-60): both
// images are multiplied by the SAME mask (mask_mode 1: segms_gt, 2: re_sil > 0 as for HO3D) and the sum of squared
// differences (-> L2 / PSNR) is accumulated next to L1 and SSIM.  The training variant compiles without it.
// MODE 2 = the self-supervised photometric terms (losses.py:317-340): x = re_img as rendered (NOT multiplied by the
// silhouette), y = maskRGBs given as `imgs` (models_res_nimble.py:220), and the L1 / mean-RGB sums are kept PER SAMPLE
// (sums[NSUMS + n], [NSUMS + N + n], [NSUMS + 2N + n]) because the reference weights them by texture_con[n]^2.
#ifndef HFR_LOSSF_MINB
#define HFR_LOSSF_MINB 3
#endif
template <int MODE>
__global__ void __launch_bounds__(kLossThreads, HFR_LOSSF_MINB) loss_fwd_kernel(HfrLossArgs a) {
  constexpr bool METRIC = MODE == 1, SELF = MODE == 2;
  extern __shared__ __align__(16) float lsm[];
  float (*xs3)[kLH][kXP] = reinterpret_cast<float (*)[kLH][kXP]>(lsm);                       // [3]
  float (*ys3)[kLH][kXP] = reinterpret_cast<float (*)[kLH][kXP]>(lsm + 3 * kLH * kXP);         // [3]
  float (*hb)[kLH][kHP] = reinterpret_cast<float (*)[kLH][kHP]>(lsm + 6 * kLH * kXP);          // [5]
  __shared__ float red[kLossThreads / 32][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
  const size_t hw = (size_t)a.H * a.W;
  const float inv_scale = 1.0f / a.sil_scale;
  float g[11];
  if (a.want_ssim) {
#pragma unroll
    for (int t = 0; t < 11; ++t) g[t] = __ldg(a.gauss + t);
  }
  float l1 = 0.f, sr = 0.f, st = 0.f, sl = 0.f, ss = 0.f, mul = 0.f, add = 0.f, l2 = 0.f;
  // derivative maps: only where the box-limited backward will read them (within 5 px of the 32x32 tiles touching the mesh box)
  bool want_d = a.dmaps != nullptr;
  if (want_d && a.dmaps_box) {
    const uint4 bx = __ldg(reinterpret_cast<const uint4*>(a.dmaps_box) + n);
    const int aa = a.dmaps_box_aa > 1 ? a.dmaps_box_aa : 1;
    const int bx0 = ((int)bx.x * 16) / aa, bx1 = ((256 - (int)bx.y) * 16 + aa - 1) / aa;   // pixel range [bx0, bx1) of the box
    const int by0 = ((int)bx.z * 16) / aa, by1 = ((256 - (int)bx.w) * 16 + aa - 1) / aa;
    constexpr int kM = kLT + kR + 2;   // a touching tile reaches 31 px beyond the box, its stencil 5 more
    want_d = x0 + kLT > bx0 - kM && x0 < bx1 + kM && y0 + kLT > by0 - kM && y0 < by1 + kM && bx0 < bx1 && by0 < by1;
  }
  bool nz_halo = false;   // any non-zero x / y sample in the tile's halo
  bool nz_x = false;      // any non-zero x sample in the halo
  __shared__ unsigned char sub_nz[64];   // per 4x4 block of the interior: holds a non-zero sample
  __shared__ float lut[256];             // k / 255 for the 8-bit target transport
  if (tid < 64) sub_nz[tid] = 0;
  if (a.imgs_u8) lut[tid] = __fdiv_rn((float)tid, 255.0f);   // kLossThreads == 256
  __syncthreads();
  // ---- halo load, all three channels at once (zero padding as F.conv2d(padding=5)) + the pointwise
  //      sums over the tile's interior
  // The loads of kHaloBatch halo positions are issued back to back before any of them is consumed, so a thread
  // has 5 x kHaloBatch requests in flight instead of paying one round trip per position.
  constexpr int kHaloIters = (kLH * kLH + kLossThreads - 1) / kLossThreads;   // 7
  constexpr int kHaloBatch = 4;
#pragma unroll
  for (int b0 = 0; b0 < kHaloIters; b0 += kHaloBatch) {
    float4 q[kHaloBatch];
    float sg[kHaloBatch], im[kHaloBatch][3];
    bool ok[kHaloBatch];
#pragma unroll
    for (int u = 0; u < kHaloBatch; ++u) {
      const int i = tid + (b0 + u) * kLossThreads;
      const int hy = i / kLH, hx = i - hy * kLH, gx = x0 + hx - kR, gy = y0 + hy - kR;
      ok[u] = (b0 + u) < kHaloIters && i < kLH * kLH && gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
      q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      sg[u] = im[u][0] = im[u][1] = im[u][2] = 0.f;
      if (ok[u]) {
        const size_t p = (size_t)gy * a.W + gx;
        if (a.nhwc) {
          q[u] = __ldg(reinterpret_cast<const float4*>(a.re_img + ((size_t)n * hw + p) * 4));
        } else {
          q[u].w = SELF ? 1.0f : __ldg(a.re_sil + (size_t)n * hw + p);
          q[u].x = __ldg(a.re_img + ((size_t)n * 3 + 0) * hw + p);
          q[u].y = __ldg(a.re_img + ((size_t)n * 3 + 1) * hw + p);
          q[u].z = __ldg(a.re_img + ((size_t)n * 3 + 2) * hw + p);
        }
        sg[u] = SELF ? 1.0f : ld_seg(a, n, p, hw);
#pragma unroll
        for (int c = 0; c < 3; ++c) im[u][c] = ld_target(a, lut, n, c, p, hw);
      }
    }
#pragma unroll
    for (int u = 0; u < kHaloBatch; ++u) {
      const int i = tid + (b0 + u) * kLossThreads;
      if ((b0 + u) >= kHaloIters || i >= kLH * kLH) continue;
      const int hy = i / kLH, hx = i - hy * kLH;
      float vx[3] = {0.f, 0.f, 0.f}, vy[3] = {0.f, 0.f, 0.f};
      if (ok[u]) {
        const float rgb[3] = {q[u].x, q[u].y, q[u].z}, sil = q[u].w, seg = sg[u];
        const float mm = a.mask_mode == 2 ? (sil > 0.0f ? 1.0f : 0.0f) : seg;   // METRIC: one mask for both images
        const float s = SELF ? 1.0f : (METRIC ? mm : sil * inv_scale), sy = SELF ? 1.0f : (METRIC ? mm : seg);
        const bool interior = hx >= kR && hx < kR + kLT && hy >= kR && hy < kR + kLT;
        bool nz_in = false;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          vx[c] = rgb[c] * s;
          vy[c] = sy * im[u][c];
          const bool nzv = vx[c] != 0.0f || vy[c] != 0.0f;
          nz_halo |= nzv;
          nz_x |= vx[c] != 0.0f;
          if (interior) {
            l1 += fabsf(vx[c] - vy[c]); sr += vx[c]; st += vy[c]; nz_in |= nzv;
            if (METRIC) { const float df = vx[c] - vy[c]; l2 += df * df; }
          }
        }
        if (nz_in) sub_nz[((hy - kR) >> 2) * 8 + ((hx - kR) >> 2)] = 1;   // benign race: every writer stores 1
        if (interior && !SELF) { sl += fabsf(sil - seg); mul += sil * seg; add += sil + seg; }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { xs3[c][hy][hx] = vx[c]; ys3[c][hy][hx] = vy[c]; }
    }
  }
  // Both SSIM inputs are masked images (x = rgb * alpha, y = target * seg): wherever the hand and the target
  // mask are absent the whole 42x42 halo is exactly zero, every windowed moment is +0 and the stencil
  // can be skipped - the SSIM value and derivative maps of such a tile are the constants of ssim_point(0...).
  const int any_halo = __syncthreads_or(nz_halo);   // also orders the sub_nz writes
  if (a.tile_flags && tid < 64) {
    const int by = blockIdx.y * 8 + (tid >> 3), bx = blockIdx.x * 8 + (tid & 7);
    const int FH = (a.H + 3) >> 2, FW = (a.W + 3) >> 2;
    if (by < FH && bx < FW) a.tile_flags[((size_t)n * FH + by) * FW + bx] = sub_nz[tid];
  }
  if (a.want_ssim && !any_halo) {
    float dm1, de11, de12;
    const float S = ssim_point(0.f, 0.f, 0.f, 0.f, 0.f, &dm1, &de11, &de12);
    const int gx = x0 + lane;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int gy = y0 + warp * 4 + o;
      if (gx < a.W && gy < a.H) {
        const size_t p = (size_t)gy * a.W + gx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          ss += S;
          if (want_d) {
            a.dmaps[((size_t)n * 9 + c * 3 + 0) * hw + p] = dm1;
            a.dmaps[((size_t)n * 9 + c * 3 + 1) * hw + p] = de11;
            a.dmaps[((size_t)n * 9 + c * 3 + 2) * hw + p] = de12;
          }
        }
      }
    }
  }
  if (a.want_ssim && any_halo) {
    // the rendered image vanishes on the whole halo (a tile of the target mask away from the hand): its three moments
    // are exactly +0 and only the two target moments go through the stencil
    const bool xzero = !__syncthreads_or(nz_x);
    for (int c = 0; c < 3; ++c) {
      if (c > 0) __syncthreads();
      ss += xzero ? ssim_channel<true>(a, xs3[c], ys3[c], hb, g, n, c, x0, y0, hw, want_d)
                  : ssim_channel<false>(a, xs3[c], ys3[c], hb, g, n, c, x0, y0, hw, want_d);
    }
  }
  // ---- block reduction of the partial sums -------------------------------------------------------------------
  float vals[8] = {l1, sr, st, sl, ss, mul, add, l2};
  constexpr int NV = METRIC ? 8 : 7;
#pragma unroll
  for (int i = 0; i < NV; ++i) vals[i] = warp_sum(vals[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i] = vals[i];
  }
  __syncthreads();
  float t = 0.f;
  if (tid < 8) {
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) t += tid < NV ? red[w][tid] : 0.f;
  }
  // where component `comp` of sample `n` goes (-1: unused), and whether it is a per-sample sum
  auto dst_of = [&](int comp, int nn) -> int {
    if (SELF) return comp == 0 ? HFR_LOSS_NSUMS + nn : (comp == 1 ? HFR_LOSS_NSUMS + a.N + nn : (comp == 2 ? HFR_LOSS_NSUMS + 2 * a.N + nn : (comp == 4 ? HFR_LOSS_SSIM : -1)));
    return comp < 5 ? comp : (comp == 5 ? HFR_LOSS_NSUMS + nn : (comp == 6 ? HFR_LOSS_NSUMS + a.N + nn : (METRIC ? HFR_LOSS_L2 : -1)));
  };
  auto per_sample = [&](int comp) -> bool { return SELF ? comp < 3 : (comp == 5 || comp == 6); };
  if (!a.partials) {      // one fp32 atomic per CTA and component: fast, order-dependent in the last bits
    if (tid < NV) {
      const int dst = dst_of(tid, n);
      if (t != 0.0f && dst >= 0) atomicAdd(a.sums + dst, t);
    }
    return;
  }
  // deterministic variant: every CTA leaves its 8 partial sums in a workspace row, the LAST CTA to arrive (ticket)
  // adds all rows in a fixed order and WRITES sums (no atomics on floating point, no zeroing needed by the caller)
  __shared__ int s_last;
  float (*s_tot)[8] = reinterpret_cast<float (*)[8]>(lsm);      // the stencil's shared arrays are dead by now
  const int ctas_per_sample = gridDim.x * gridDim.y, n_cta = ctas_per_sample * gridDim.z;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (tid < 8) a.partials[(size_t)cta * 8 + tid] = t;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(a.ticket, 1u) == (unsigned)(n_cta - 1);
  __syncthreads();             // (also: every thread is done with the stencil arrays that s_tot aliases)
  if (!s_last) return;
  __threadfence();
  {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = tid; r < n_cta; r += kLossThreads) {
      const float4 q0 = __ldcg(reinterpret_cast<const float4*>(a.partials + (size_t)r * 8));
      const float4 q1 = __ldcg(reinterpret_cast<const float4*>(a.partials + (size_t)r * 8 + 4));
      acc[0] += q0.x; acc[1] += q0.y; acc[2] += q0.z; acc[3] += q0.w; acc[4] += q1.x; acc[5] += q1.y; acc[6] += q1.z; acc[7] += q1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_tot[tid][i] = acc[i];
    __syncthreads();
    for (int o = kLossThreads / 2; o > 0; o >>= 1) {
      if (tid < o) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_tot[tid][i] += s_tot[tid + o][i];
      }
      __syncthreads();
    }
    if (tid < 8 && !per_sample(tid)) {
      const int dst = dst_of(tid, 0);
      if (dst >= 0) a.sums[dst] = s_tot[0][tid];
    }
    // per-sample sums: the CTAs of a sample are consecutive rows
    for (int i = tid; i < a.N * 8; i += kLossThreads) {
      const int nn = i >> 3, comp = i & 7;
      if (!per_sample(comp)) continue;
      float sacc = 0.f;
      for (int r = 0; r < ctas_per_sample; ++r) sacc += __ldcg(a.partials + ((size_t)nn * ctas_per_sample + r) * 8 + comp);
      a.sums[dst_of(comp, nn)] = sacc;
    }
    if (tid == 0) *a.ticket = 0u;     // ready for the next launch
  }
}

#ifndef HFR_LOSSB_UNROLLC
#define HFR_LOSSB_UNROLLC 1
#endif
#ifndef HFR_LOSSB_MINB
#define HFR_LOSSB_MINB 3
#endif
// SELF = backward of the self-supervised terms (losses.py:317-340; forward MODE 2): gradients reach re_img only.
template <bool SELF>
__global__ void __launch_bounds__(kLossThreads, HFR_LOSSB_MINB) loss_bwd_kernel(HfrLossBwdArgs b) {
  const HfrLossArgs& a = b.f;
  __shared__ __align__(16) float sd[3][kLH][kXP];
  __shared__ __align__(16) float hb[3][kLH][kHP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT;
  const size_t hw = (size_t)a.H * a.W;
  if (b.tile_box) {
    // Only pixels that hold a fragment pass their gradient on: a 32x32 tile whose rasterizer tiles (16x16 at the
    // rasterised resolution, box_aa x finer than the loss resolution) all lie outside the mesh's tile box is skipped.
    const uint4 bx = __ldg(reinterpret_cast<const uint4*>(b.tile_box) + n);
    const int aa = b.box_aa > 1 ? b.box_aa : 1;
    const int tx0 = (x0 * aa) >> 4, tx1 = ((x0 + kLT) * aa - 1) >> 4, ty0 = (y0 * aa) >> 4, ty1 = ((y0 + kLT) * aa - 1) >> 4;
    if (tx1 < (int)bx.x || tx0 > 255 - (int)bx.y || ty1 < (int)bx.z || ty0 > 255 - (int)bx.w) return;
  }
  bool ssim = a.want_ssim && a.dmaps;
  if (ssim && a.tile_flags) {
    // d(SSIM)/dx at a pixel is r0 + 2 x r1 + y r2 with r = Gauss * dmaps: when x and y vanish within 12 pixels
    // of the tile (>= the 10-pixel reach of the two stacked stencils) dm1 is 0 on the whole halo, so r0 = 0
    // and the other two terms are multiplied by x = y = 0: the stencil contributes exactly nothing.
    int live = 0;
    if (tid < 14 * 14) {   // the 14x14 blocks of 4x4 pixels covering the tile and 12 pixels around it
      const int by = (int)blockIdx.y * 8 - 3 + tid / 14, bx = (int)blockIdx.x * 8 - 3 + tid % 14;
      const int FH = (a.H + 3) >> 2, FW = (a.W + 3) >> 2;
      if (by >= 0 && by < FH && bx >= 0 && bx < FW) live = a.tile_flags[((size_t)n * FH + by) * FW + bx];
    }
    ssim = __syncthreads_or(live) != 0;
  }
  float g[11];
  if (ssim) {
#pragma unroll
    for (int t = 0; t < 11; ++t) g[t] = __ldg(b.gauss + t);
  }
  const float w_tex = __ldg(b.w), w_mrgb = __ldg(b.w + 1), w_ssim = __ldg(b.w + 2), w_sil = __ldg(b.w + 3), w_iou = __ldg(b.w + 4);
  const float cnt = (float)b.count_global, icnt = 1.0f / cnt;
  // (the global sums may be in flight in an all-reduce when skip_mrgb is set: they are not touched then)
  const float mR = b.skip_mrgb ? 0.f : a.sums[HFR_LOSS_SUM_R] * icnt, mT = b.skip_mrgb ? 0.f : a.sums[HFR_LOSS_SUM_T] * icnt;
  // skip_mrgb: the mean-RGB term's gradient is a per-step scalar times (alpha, rgb); the fused backward adds it from
  // the (all-reduced) sums, so that this kernel does not have to wait for them
  float k_mrgb = b.skip_mrgb ? 0.0f : w_mrgb * 2.0f * (mT - mR) * (-icnt);
  float k_l1 = w_tex * icnt;
  if (SELF) {
    // texture_self = sum_n c_n^2 sum|x - y| / (3HW sum_n c_n^2); mrgb_self = sum_n c_n^2 |mean_n x - mean_n y| / sum_n c_n^2
    const float per = 1.0f / (3.0f * (float)hw);
    const float c = __ldg(b.tex_con + n), c2 = c * c / __ldg(b.self_norm);
    const float dm = (a.sums[HFR_LOSS_NSUMS + a.N + n] - a.sums[HFR_LOSS_NSUMS + 2 * a.N + n]) * per;
    k_l1 = w_tex * c2 * per;
    k_mrgb = w_mrgb * c2 * per * (dm > 0.f ? 1.f : (dm < 0.f ? -1.f : 0.f));
  }
  const float inv_scale = SELF ? 1.0f : 1.0f / a.sil_scale;
  // this thread's 4 pixels: column x0+lane, rows y0 + warp*4 + o.  Everything the pointwise part needs (rendered
  // RGBA, mask, target) is requested up front for all channels, so the round trips overlap each other and, on
  // tiles with a live SSIM stencil, the halo loads and the stencil itself.
  const int gx = x0 + lane;
  __shared__ float lut[256];             // k / 255 for the 8-bit target transport
  if (a.imgs_u8) {                       // (uniform branch)
    lut[tid] = __fdiv_rn((float)tid, 255.0f);
    __syncthreads();
  }
  float sil[4], seg[4], rimg[4][3], timg[4][3];
  bool in[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int gy = y0 + warp * 4 + o;
    in[o] = gx < a.W && gy < a.H;
    sil[o] = seg[o] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) rimg[o][c] = timg[o][c] = 0.f;
    if (in[o]) {
      const size_t p = (size_t)gy * a.W + gx;
      if (a.nhwc) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.re_img + ((size_t)n * hw + p) * 4));
        rimg[o][0] = q.x; rimg[o][1] = q.y; rimg[o][2] = q.z; sil[o] = q.w;
      } else {
        if (!SELF) sil[o] = __ldg(a.re_sil + (size_t)n * hw + p);
#pragma unroll
        for (int c = 0; c < 3; ++c) rimg[o][c] = __ldg(a.re_img + ((size_t)n * 3 + c) * hw + p);
      }
      seg[o] = SELF ? 1.0f : ld_seg(a, n, p, hw);
      if (SELF) sil[o] = 1.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) timg[o][c] = ld_target(a, lut, n, c, p, hw);
    }
  }
  const float mulv = a.sums[HFR_LOSS_NSUMS + n], addv = a.sums[HFR_LOSS_NSUMS + a.N + n];
  const float den = addv - mulv;
  float gsil[4], grgb[4][3];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    gsil[o] = 0.f;
    if (in[o] && !SELF) {
      const float d = sil[o] - seg[o];
      gsil[o] = w_sil * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / ((float)b.n_global * (float)hw);
      gsil[o] += w_iou * (-1.0f / (float)b.n_global) * (seg[o] * den - mulv * (1.0f - seg[o])) / (den * den);
    }
  }
#if HFR_LOSSB_UNROLLC
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int c = 0; c < 3; ++c) {
    float r[4][3];
#pragma unroll
    for (int o = 0; o < 4; ++o) r[o][0] = r[o][1] = r[o][2] = 0.f;
    if (ssim) {
      if (c > 0) __syncthreads();
      if ((a.W & 3) == 0) {
        // rows are loaded as 12 aligned float4 covering columns x0-8 .. x0+39 (halo = x0-5 .. x0+36); all of a
        // thread's requests go out before the first one is stored to shared memory
        constexpr int kTot = 3 * kLH * 12, kIt = (kTot + kLossThreads - 1) / kLossThreads;   // 1512, 6
        float4 v[kIt];
#pragma unroll
        for (int u = 0; u < kIt; ++u) {
          const int i = tid + u * kLossThreads;
          const int m = i / (kLH * 12), rem = i - m * (kLH * 12), hy = rem / 12, q = rem - hy * 12;
          const int qy = y0 + hy - kR, qx = x0 - 8 + 4 * q;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < kTot && qy >= 0 && qy < a.H && qx >= 0 && qx < a.W)
            v[u] = __ldg(reinterpret_cast<const float4*>(a.dmaps + ((size_t)n * 9 + c * 3 + m) * hw + (size_t)qy * a.W + qx));
        }
#pragma unroll
        for (int u = 0; u < kIt; ++u) {
          const int i = tid + u * kLossThreads;
          if (i < kTot) {
            const int m = i / (kLH * 12), rem = i - m * (kLH * 12), hy = rem / 12, q = rem - hy * 12;
            const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int hx = 4 * q + t - 3;
              if (hx >= 0 && hx < kLH) sd[m][hy][hx] = e[t];
            }
          }
        }
      } else {
        for (int i = tid; i < kLH * kLH; i += kLossThreads) {
          const int hy = i / kLH, hx = i - hy * kLH, qx = x0 + hx - kR, qy = y0 + hy - kR;
          const bool ok = qx >= 0 && qx < a.W && qy >= 0 && qy < a.H;
          const size_t q = ok ? (size_t)qy * a.W + qx : 0;
#pragma unroll
          for (int m = 0; m < 3; ++m) sd[m][hy][hx] = ok ? __ldg(a.dmaps + ((size_t)n * 9 + c * 3 + m) * hw + q) : 0.f;
        }
      }
      __syncthreads();
      if (tid < kHTasks) {
        const int row = tid >> 2, strip = tid & 3;
        float acc[8][3];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o][0] = acc[o][1] = acc[o][2] = 0.f;
        const float4* r0 = reinterpret_cast<const float4*>(&sd[0][row][strip * 8]);
        const float4* r1 = reinterpret_cast<const float4*>(&sd[1][row][strip * 8]);
        const float4* r2 = reinterpret_cast<const float4*>(&sd[2][row][strip * 8]);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float4 v0 = r0[q], v1 = r1[q], v2 = r2[q];
          const float a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w}, a2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            if (j < 18) {
              const float val[3] = {a0[e], a1[e], a2[e]};
              tap_accumulate<3>(acc, val, j, g);
            }
          }
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          float4* dst = reinterpret_cast<float4*>(&hb[m][row][strip * 8]);
          dst[0] = make_float4(acc[0][m], acc[1][m], acc[2][m], acc[3][m]);
          dst[1] = make_float4(acc[4][m], acc[5][m], acc[6][m], acc[7][m]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 14; ++j) {
        const float v0 = hb[0][warp * 4 + j][lane], v1 = hb[1][warp * 4 + j][lane], v2 = hb[2][warp * 4 + j][lane];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int t = j - o;
          if (t >= 0 && t < 11) {
            r[o][0] = fmaf(g[t], v0, r[o][0]); r[o][1] = fmaf(g[t], v1, r[o][1]); r[o][2] = fmaf(g[t], v2, r[o][2]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      // channel c of the per-pixel registers (select chain: c is a runtime index)
      const float rim = c == 0 ? rimg[o][0] : (c == 1 ? rimg[o][1] : rimg[o][2]);
      const float tim = c == 0 ? timg[o][0] : (c == 1 ? timg[o][1] : timg[o][2]);
      const float so = sil[o] * inv_scale;
      const float xv = rim * so, yv = seg[o] * tim;
      const float gS = r[o][0] + 2.0f * xv * r[o][1] + yv * r[o][2];
      const float d = xv - yv;
      const float grim = k_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) + k_mrgb + w_ssim * (-icnt) * gS;
      const float gc = in[o] ? grim * so : 0.f;
      if (c == 0) grgb[o][0] = gc; else if (c == 1) grgb[o][1] = gc; else grgb[o][2] = gc;
      if (in[o] && !SELF) gsil[o] += grim * rim * inv_scale;
    }
  }
  if (b.gmax_bits) {   // max |gradient| of this launch (positive floats order like their bit patterns): order-independent
    float m = 0.f;
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if (in[o]) m = fmaxf(fmaxf(fmaxf(m, fabsf(grgb[o][0])), fmaxf(fabsf(grgb[o][1]), fabsf(grgb[o][2]))), fabsf(gsil[o]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.f) atomicMax(b.gmax_bits, __float_as_uint(m));
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    if (in[o]) {
      const size_t p = (size_t)(y0 + warp * 4 + o) * a.W + gx;
      if (a.nhwc) {
        *reinterpret_cast<float4*>(b.g_re_img + ((size_t)n * hw + p) * 4) = make_float4(grgb[o][0], grgb[o][1], grgb[o][2], gsil[o]);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) b.g_re_img[((size_t)n * 3 + c) * hw + p] = grgb[o][c];
        if (!SELF) b.g_re_sil[n * hw + p] = gsil[o];
      }
    }
  }
}

}  // namespace hfr

extern "C" int hfr_pool_forward(const HfrPoolArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->N >= 0 && a->H > 0 && a->W > 0 && a->aa >= 1, "pool_forward: bad dims");
  if (a->N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->image && a->re_img && a->re_sil, "pool_forward: null pointer");
  const size_t total = (size_t)a->N * a->H * a->W;
  pool_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("pool_forward");
  return HFR_OK;
}

extern "C" int hfr_pool_backward(const HfrPoolBwdArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->N >= 0 && a->H > 0 && a->W > 0 && a->aa >= 1, "pool_backward: bad dims");
  if (a->N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->g_image, "pool_backward: null pointer");
  const size_t total = (size_t)a->N * a->H * a->aa * a->W * a->aa;
  pool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("pool_backward");
  return HFR_OK;
}

extern "C" int64_t hfr_loss_partials_floats(int32_t N, int32_t H, int32_t W) {
  using namespace hfr;
  return (int64_t)N * ((W + kLT - 1) / kLT) * ((H + kLT - 1) / kLT) * 8;
}

extern "C" int hfr_loss_forward(const HfrLossArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->N >= 0 && a->H > 0 && a->W > 0 && a->sil_scale > 0.f, "loss_forward: bad dims");
  if (a->N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->re_img && (a->nhwc || a->re_sil || a->mask_mode == 3) && (a->imgs || a->imgs_u8) && (a->seg || a->seg_u8 || a->mask_mode == 3) && a->sums, "loss_forward: null pointer");
  HFR_CHECK_ARG(!a->want_ssim || a->gauss, "loss_forward: SSIM needs the Gaussian taps");
  dim3 grid((a->W + kLT - 1) / kLT, (a->H + kLT - 1) / kLT, a->N);
  HFR_CHECK_ARG(a->mask_mode >= 0 && a->mask_mode <= 3, "loss_forward: mask_mode must be 0..3");
  HFR_CHECK_ARG(!a->partials || (a->ticket && (reinterpret_cast<uintptr_t>(a->partials) & 15) == 0),
                "loss_forward: the deterministic reduction needs a 16-byte aligned partials workspace and a ticket word");
  HFR_CHECK_ARG(a->mask_mode == 0 || a->mask_mode == 3 || !(a->want_grad && a->dmaps), "loss_forward: the metric modes have no backward");
  static bool attr_set = false;   // benign race: the attribute is idempotent
  if (!attr_set) {
    cudaFuncSetAttribute(loss_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLossFwdSmem);
    cudaFuncSetAttribute(loss_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLossFwdSmem);
    cudaFuncSetAttribute(loss_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLossFwdSmem);
    attr_set = true;
  }
  if (a->mask_mode == 0) loss_fwd_kernel<0><<<grid, kLossThreads, kLossFwdSmem, (cudaStream_t)stream>>>(*a);
  else if (a->mask_mode == 3) loss_fwd_kernel<2><<<grid, kLossThreads, kLossFwdSmem, (cudaStream_t)stream>>>(*a);
  else loss_fwd_kernel<1><<<grid, kLossThreads, kLossFwdSmem, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("loss_forward");
  return HFR_OK;
}

extern "C" int hfr_loss_backward(const HfrLossBwdArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->f.N >= 0 && a->f.H > 0 && a->f.W > 0, "loss_backward: bad dims");
  if (a->f.N == 0) return HFR_OK;
  const bool self = a->f.mask_mode == 3;
  HFR_CHECK_ARG(a->f.mask_mode == 0 || self, "loss_backward: mask_mode must be 0 (training terms) or 3 (self-supervised terms)");
  HFR_CHECK_ARG(a->f.re_img && (a->f.nhwc || self || (a->f.re_sil && a->g_re_sil)) && (a->f.imgs || a->f.imgs_u8) && (self || a->f.seg || a->f.seg_u8) && a->f.sums && a->w && a->g_re_img,
                "loss_backward: null pointer");
  HFR_CHECK_ARG(!self || (a->tex_con && a->self_norm && !a->f.nhwc), "loss_backward: the self-supervised terms need tex_con, self_norm and NCHW images");
  HFR_CHECK_ARG(!(a->f.want_ssim && !a->f.dmaps), "loss_backward: want_ssim is set but the forward wrote no derivative maps (dmaps)");
  HFR_CHECK_ARG(!(a->f.want_ssim && a->f.dmaps) || a->gauss, "loss_backward: SSIM needs the Gaussian taps");
  HFR_CHECK_ARG(a->count_global > 0 && a->n_global > 0, "loss_backward: bad global counts");
  dim3 grid((a->f.W + kLT - 1) / kLT, (a->f.H + kLT - 1) / kLT, a->f.N);
  if (self) loss_bwd_kernel<true><<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(*a);
  else loss_bwd_kernel<false><<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("loss_backward");
  return HFR_OK;
}

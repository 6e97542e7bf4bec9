// SSAA pooling / output split and the render-dependent losses for sm_100a, forward + backward.
//
//   pool : models_res_nimble.py:210-220 (NHWC->NCHW, avg_pool2d(aa), RGB/alpha split, binarise, maskRGBs)
//   loss : losses.py:355-378 (texture L1, mean-RGB, SSIM), :399-408 (silhouette L1, IoU) with
//          utils/losses_util.py:366-378 and utils/pytorch_ssim/__init__.py:17-37.
//
// SSIM is a separable 11x11 Gaussian stencil evaluated per 16x16 tile out of shared memory
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
constexpr int kT = 16, kR = 5, kHalo = kT + 2 * kR;   // 16x16 tile, 11-tap window

// rendered colour / silhouette of pixel p of sample n in either layout
__device__ __forceinline__ float ld_rgb(const HfrLossArgs& a, int n, int c, size_t p, size_t hw) {
  return a.nhwc ? a.re_img[((size_t)n * hw + p) * 4 + c] : a.re_img[((size_t)n * 3 + c) * hw + p];
}
__device__ __forceinline__ float ld_sil(const HfrLossArgs& a, int n, size_t p, size_t hw) {
  return a.nhwc ? a.re_img[((size_t)n * hw + p) * 4 + 3] : a.re_sil[(size_t)n * hw + p];
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

// grid (tiles_x, tiles_y, N), 256 threads = one 16x16 tile.  Pointwise sums + SSIM.
__global__ void __launch_bounds__(256) loss_fwd_kernel(HfrLossArgs a) {
  __shared__ float sx[kHalo][kHalo + 1], sy[kHalo][kHalo + 1];
  __shared__ float hbuf[5][kHalo][kT + 1];
  __shared__ float g[11];
  __shared__ float scratch[8];
  const int n = blockIdx.z, tx = threadIdx.x % kT, ty = threadIdx.x / kT;
  const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
  const int x = x0 + tx, y = y0 + ty;
  const bool in = x < a.W && y < a.H;
  const size_t hw = (size_t)a.H * a.W;
  if (threadIdx.x < 11 && a.want_ssim) g[threadIdx.x] = a.gauss[threadIdx.x];
  float l1 = 0.f, sr = 0.f, st = 0.f, sl = 0.f, ss = 0.f, mul = 0.f, add = 0.f;
  if (in) {
    const size_t p = (size_t)y * a.W + x;
    const float sil = ld_sil(a, n, p, hw), seg = a.seg[n * hw + p];
    const float s = sil / a.sil_scale;
    for (int c = 0; c < 3; ++c) {
      const float rim = ld_rgb(a, n, c, p, hw) * s;
      const float tgt = seg * a.imgs[((size_t)n * 3 + c) * hw + p];
      l1 += fabsf(rim - tgt); sr += rim; st += tgt;
    }
    sl = fabsf(sil - seg); mul = sil * seg; add = sil + seg;
  }
  if (a.want_ssim) {
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    for (int c = 0; c < 3; ++c) {
      __syncthreads();
      for (int i = threadIdx.x; i < kHalo * kHalo; i += 256) {
        const int hx = i % kHalo, hy = i / kHalo, gx = x0 + hx - kR, gy = y0 + hy - kR;
        float vx = 0.f, vy = 0.f;
        if (gx >= 0 && gx < a.W && gy >= 0 && gy < a.H) {
          const size_t p = (size_t)gy * a.W + gx;
          vx = ld_rgb(a, n, c, p, hw) * (ld_sil(a, n, p, hw) / a.sil_scale);
          vy = a.seg[n * hw + p] * a.imgs[((size_t)n * 3 + c) * hw + p];
        }
        sx[hy][hx] = vx; sy[hy][hx] = vy;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < kHalo * kT; i += 256) {   // horizontal pass
        const int ox = i % kT, hy = i / kT;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; ++t) {
          const float w = g[t], vx = sx[hy][ox + t], vy = sy[hy][ox + t];
          m1 += w * vx; m2 += w * vy; e11 += w * (vx * vx); e22 += w * (vy * vy); e12 += w * (vx * vy);
        }
        hbuf[0][hy][ox] = m1; hbuf[1][hy][ox] = m2; hbuf[2][hy][ox] = e11; hbuf[3][hy][ox] = e22; hbuf[4][hy][ox] = e12;
      }
      __syncthreads();
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) {                           // vertical pass
        const float w = g[t];
        m1 += w * hbuf[0][ty + t][tx]; m2 += w * hbuf[1][ty + t][tx]; e11 += w * hbuf[2][ty + t][tx];
        e22 += w * hbuf[3][ty + t][tx]; e12 += w * hbuf[4][ty + t][tx];
      }
      if (in) {
        const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
        const float s1 = e11 - m11, s2 = e22 - m22, s12 = e12 - m12;
        const float A1 = 2.f * m12 + C1, A2 = 2.f * s12 + C2, B1 = m11 + m22 + C1, B2 = s1 + s2 + C2;
        const float S = (A1 * A2) / (B1 * B2);
        ss += S;
        if (a.dmaps) {
          const float ib = 1.0f / (B1 * B2);
          const float dm1 = 2.f * m2 * A2 * ib - 2.f * m2 * A1 * ib - S * 2.f * m1 / B1 + S * 2.f * m1 / B2;
          const float de11 = -S / B2, de12 = 2.f * A1 * ib;
          const size_t p = (size_t)y * a.W + x;
          a.dmaps[((size_t)n * 9 + c * 3 + 0) * hw + p] = dm1;
          a.dmaps[((size_t)n * 9 + c * 3 + 1) * hw + p] = de11;
          a.dmaps[((size_t)n * 9 + c * 3 + 2) * hw + p] = de12;
        }
      }
    }
  }
  const float vals[7] = {l1, sr, st, sl, ss, mul, add};
  const int dst[7] = {HFR_LOSS_L1, HFR_LOSS_SUM_R, HFR_LOSS_SUM_T, HFR_LOSS_SIL, HFR_LOSS_SSIM,
                      HFR_LOSS_NSUMS + n, HFR_LOSS_NSUMS + a.N + n};
  for (int i = 0; i < 7; ++i) {
    const float t = block_sum(vals[i], scratch);
    if (threadIdx.x == 0 && t != 0.0f) atomicAdd(a.sums + dst[i], t);
  }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(HfrLossBwdArgs b) {
  const HfrLossArgs& a = b.f;
  __shared__ float sd[3][kHalo][kHalo + 1];
  __shared__ float hbuf[3][kHalo][kT + 1];
  __shared__ float g[11];
  const int n = blockIdx.z, tx = threadIdx.x % kT, ty = threadIdx.x / kT;
  const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
  const int x = x0 + tx, y = y0 + ty;
  const bool in = x < a.W && y < a.H;
  const size_t hw = (size_t)a.H * a.W;
  const bool ssim = a.want_ssim && a.dmaps;
  if (threadIdx.x < 11 && ssim) g[threadIdx.x] = b.gauss[threadIdx.x];
  const float w_tex = b.w[0], w_mrgb = b.w[1], w_ssim = b.w[2], w_sil = b.w[3], w_iou = b.w[4];
  const float cnt = (float)b.count_global;
  const float mR = a.sums[HFR_LOSS_SUM_R] / cnt, mT = a.sums[HFR_LOSS_SUM_T] / cnt;
  const float k_mrgb = w_mrgb * 2.0f * (mT - mR) * (-1.0f / cnt);
  float sil = 0.f, seg = 0.f, s = 0.f, gsil = 0.f;
  size_t p = 0;
  if (in) {
    p = (size_t)y * a.W + x;
    sil = ld_sil(a, n, p, hw); seg = a.seg[n * hw + p];
    s = sil / a.sil_scale;
    const float d = sil - seg;
    gsil = w_sil * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / ((float)b.n_global * (float)hw);
    const float mul = a.sums[HFR_LOSS_NSUMS + n], add = a.sums[HFR_LOSS_NSUMS + a.N + n];
    const float den = add - mul;
    gsil += w_iou * (-1.0f / (float)b.n_global) * (seg * den - mul * (1.0f - seg)) / (den * den);
  }
  for (int c = 0; c < 3; ++c) {
    float gS = 0.0f, xv = 0.f, yv = 0.f, rimg = 0.f;
    if (in) {
      rimg = ld_rgb(a, n, c, p, hw);
      xv = rimg * s;
      yv = seg * a.imgs[((size_t)n * 3 + c) * hw + p];
    }
    if (ssim) {
      __syncthreads();
      for (int i = threadIdx.x; i < kHalo * kHalo; i += 256) {
        const int hx = i % kHalo, hy = i / kHalo, gx = x0 + hx - kR, gy = y0 + hy - kR;
        const bool ok = gx >= 0 && gx < a.W && gy >= 0 && gy < a.H;
        const size_t q = ok ? (size_t)gy * a.W + gx : 0;
        for (int m = 0; m < 3; ++m) sd[m][hy][hx] = ok ? a.dmaps[((size_t)n * 9 + c * 3 + m) * hw + q] : 0.f;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < kHalo * kT; i += 256) {
        const int ox = i % kT, hy = i / kT;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; ++t) { const float w = g[t]; r0 += w * sd[0][hy][ox + t]; r1 += w * sd[1][hy][ox + t]; r2 += w * sd[2][hy][ox + t]; }
        hbuf[0][hy][ox] = r0; hbuf[1][hy][ox] = r1; hbuf[2][hy][ox] = r2;
      }
      __syncthreads();
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int t = 0; t < 11; ++t) { const float w = g[t]; r0 += w * hbuf[0][ty + t][tx]; r1 += w * hbuf[1][ty + t][tx]; r2 += w * hbuf[2][ty + t][tx]; }
      gS = r0 + 2.0f * xv * r1 + yv * r2;
    }
    if (in) {
      const float d = xv - yv;
      float grim = w_tex * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / cnt + k_mrgb + w_ssim * (-1.0f / cnt) * gS;
      if (a.nhwc) b.g_re_img[((size_t)n * hw + p) * 4 + c] = grim * s;
      else b.g_re_img[((size_t)n * 3 + c) * hw + p] = grim * s;
      gsil += grim * rimg / a.sil_scale;
    }
  }
  if (in) {
    if (a.nhwc) b.g_re_img[((size_t)n * hw + p) * 4 + 3] = gsil;
    else b.g_re_sil[n * hw + p] = gsil;
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

extern "C" int hfr_loss_forward(const HfrLossArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->N >= 0 && a->H > 0 && a->W > 0 && a->sil_scale > 0.f, "loss_forward: bad dims");
  if (a->N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->re_img && (a->nhwc || a->re_sil) && a->imgs && a->seg && a->sums, "loss_forward: null pointer");
  HFR_CHECK_ARG(!a->want_ssim || a->gauss, "loss_forward: SSIM needs the Gaussian taps");
  dim3 grid((a->W + kT - 1) / kT, (a->H + kT - 1) / kT, a->N);
  loss_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("loss_forward");
  return HFR_OK;
}

extern "C" int hfr_loss_backward(const HfrLossBwdArgs* a, void* stream) {
  using namespace hfr;
  HFR_CHECK_ARG(a && a->f.N >= 0 && a->f.H > 0 && a->f.W > 0, "loss_backward: bad dims");
  if (a->f.N == 0) return HFR_OK;
  HFR_CHECK_ARG(a->f.re_img && (a->f.nhwc || (a->f.re_sil && a->g_re_sil)) && a->f.imgs && a->f.seg && a->f.sums && a->w && a->g_re_img,
                "loss_backward: null pointer");
  HFR_CHECK_ARG(!(a->f.want_ssim && a->f.dmaps) || a->gauss, "loss_backward: SSIM needs the Gaussian taps");
  HFR_CHECK_ARG(a->count_global > 0 && a->n_global > 0, "loss_backward: bad global counts");
  dim3 grid((a->f.W + kT - 1) / kT, (a->f.H + kT - 1) / kT, a->f.N);
  loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  HFR_CHECK_LAUNCH("loss_backward");
  return HFR_OK;
}

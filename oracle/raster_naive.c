/* TEST INFRASTRUCTURE — scalar C restatement of PyTorch3D's naive CPU rasterizer.
 *
 * PARITY UNPINNED (see oracle/__init__.py): restates the published algorithm of
 * RasterizeMeshesNaiveCpu (pytorch3d/csrc/rasterize_meshes/rasterize_meshes_cpu.cpp)
 * and csrc/utils/geometry_utils.h as summarised in SURVEY.md Appendix A.3-A.4.
 * The reference reaches it through models_res_nimble.py:208 when tensors are on
 * the CPU (bin_size forced to 0).  O(P*F) per mesh, one priority list per pixel.
 *
 * Build (no FMA contraction, so fp32 results are reproducible):
 *   gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC raster_naive.c -o _build/libraster_naive.so
 * `threads` = 1 reproduces upstream's single-threaded behaviour; >1 splits image
 * rows over OpenMP threads (results are identical: pixels are independent).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KEPS 1e-8f
#define MAXK 64

static inline float pix_to_ndc(int i, int S1, int S2) {
  const float range = S1 > S2 ? (2.0f * (float)S1) / (float)S2 : 2.0f;
  const float offset = range / 2.0f;
  return -offset + (range * (float)i + offset) / (float)S1;
}
static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}
static inline float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }
static inline float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
static inline float clamp01(float t) { return fminf(fmaxf(t, 0.0f), 1.0f); }

static inline float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= KEPS) {
    const float ex = px - bx, ey = py - by;
    return ex * ex + ey * ey;
  }
  float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  t = clamp01(t);
  const float qx = ax + t * bax, qy = ay + t * bay;
  const float dx = qx - px, dy = qy - py;
  return dx * dx + dy * dy;
}

typedef struct { float z; int64_t f; float d, b0, b1, b2; } Hit;

/* face_verts: (F_total,3,3) NDC xy + view z.  Outputs (N,H,W,K[,3]) pre-sized by the caller. */
int hfr_oracle_rasterize_naive(const float* face_verts, const int64_t* mesh_first, const int64_t* mesh_nf,
                               int N, int H, int W, int K, float blur_radius, int perspective_correct,
                               int clip_bary, int cull_backfaces, int threads,
                               int64_t* pix_to_face, float* zbuf, float* bary, float* dists) {
  if (K > MAXK || K < 1) return 1;
  const float r = sqrtf(blur_radius);
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  const long rows = (long)N * H;
#pragma omp parallel for schedule(dynamic, 4)
  for (long row = 0; row < rows; ++row) {
    const int n = (int)(row / H), yi = (int)(row % H);
    const float yf = pix_to_ndc(H - 1 - yi, H, W);
    const int64_t f0 = mesh_first[n], nf = mesh_nf[n];
    for (int xi = 0; xi < W; ++xi) {
      const float xf = pix_to_ndc(W - 1 - xi, W, H);
      Hit q[MAXK];
      int cnt = 0;
      for (int64_t f = f0; f < f0 + nf; ++f) {
        const float* v = face_verts + f * 9;
        const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7],
                    z2 = v[8];
        const float xmin = fmin3(x0, x1, x2) - r, xmax = fmax3(x0, x1, x2) + r;
        const float ymin = fmin3(y0, y1, y2) - r, ymax = fmax3(y0, y1, y2) + r;
        const float zmin = fmin3(z0, z1, z2);
        if (xf < xmin || xf > xmax || yf < ymin || yf > ymax || zmin < KEPS) continue;
        const float face_area = edge_fn(x0, y0, x1, y1, x2, y2);
        if (face_area <= KEPS && face_area >= -KEPS) continue;
        if (cull_backfaces && face_area < 0.0f) continue;
        const float area = edge_fn(x2, y2, x0, y0, x1, y1) + KEPS;
        float b0 = edge_fn(xf, yf, x1, y1, x2, y2) / area;
        float b1 = edge_fn(xf, yf, x2, y2, x0, y0) / area;
        float b2 = edge_fn(xf, yf, x0, y0, x1, y1) / area;
        if (perspective_correct) {
          const float t0 = b0 * z1 * z2, t1 = z0 * b1 * z2, t2 = z0 * z1 * b2;
          const float den = fmaxf(t0 + t1 + t2, KEPS);
          b0 = t0 / den; b1 = t1 / den; b2 = t2 / den;
        }
        float c0 = b0, c1 = b1, c2 = b2;
        if (clip_bary) {
          c0 = clamp01(b0); c1 = clamp01(b1); c2 = clamp01(b2);
          const float s = fmaxf(c0 + c1 + c2, 1e-5f);
          c0 /= s; c1 /= s; c2 /= s;
        }
        const float pz = c0 * z0 + c1 * z1 + c2 * z2;
        if (pz < 0.0f) continue;
        const float e01 = seg_dist2(xf, yf, x0, y0, x1, y1);
        const float e02 = seg_dist2(xf, yf, x0, y0, x2, y2);
        const float e12 = seg_dist2(xf, yf, x1, y1, x2, y2);
        const float dist = fminf(fminf(e01, e02), e12);
        const int inside = b0 > 0.0f && b1 > 0.0f && b2 > 0.0f;
        if (!inside && dist >= blur_radius) continue;
        /* keep the K smallest (z, f): faces arrive in increasing f, so strict < keeps ties in f order */
        if (cnt == K && !(pz < q[K - 1].z)) continue;
        int pos = cnt < K ? cnt : K - 1;
        while (pos > 0 && pz < q[pos - 1].z) { q[pos] = q[pos - 1]; --pos; }
        q[pos].z = pz; q[pos].f = f; q[pos].d = inside ? -dist : dist;
        q[pos].b0 = c0; q[pos].b1 = c1; q[pos].b2 = c2;
        if (cnt < K) ++cnt;
      }
      const long base = ((long)row * W + xi) * K;
      for (int k = 0; k < K; ++k) {
        if (k < cnt) {
          pix_to_face[base + k] = q[k].f; zbuf[base + k] = q[k].z; dists[base + k] = q[k].d;
          bary[(base + k) * 3 + 0] = q[k].b0; bary[(base + k) * 3 + 1] = q[k].b1;
          bary[(base + k) * 3 + 2] = q[k].b2;
        } else {
          pix_to_face[base + k] = -1; zbuf[base + k] = -1.0f; dists[base + k] = -1.0f;
          bary[(base + k) * 3 + 0] = bary[(base + k) * 3 + 1] = bary[(base + k) * 3 + 2] = -1.0f;
        }
      }
    }
  }
  return 0;
}

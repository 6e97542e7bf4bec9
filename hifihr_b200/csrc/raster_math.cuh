// Per (pixel, face) rasterization math, host + device, forward and backward.
//
// Semantics: PyTorch3D rasterize_meshes (csrc/utils/geometry_utils.h, rasterize_meshes_cpu.cpp)
// as restated in SURVEY.md Appendix A.3-A.5; reached from models_res_nimble.py:208.
// Every operation goes through the X* wrappers (common.cuh) in the oracle's order, so the
// fp32 results are bit-identical to oracle/raster_naive.c and oracle/p3d.py.
#pragma once
#include "common.cuh"

#define HFR_KEPS 1e-8f

HFR_HD float hfr_pix_to_ndc(int i, int S1, int S2) {
  const float range = S1 > S2 ? XDIV(XMUL(2.0f, (float)S1), (float)S2) : 2.0f;
  const float offset = XDIV(range, 2.0f);
  return XADD(-offset, XDIV(XADD(XMUL(range, (float)i), offset), (float)S1));
}

HFR_HD float hfr_edge(float px, float py, float ax, float ay, float bx, float by) {
  return XSUB(XMUL(XSUB(px, ax), XSUB(by, ay)), XMUL(XSUB(py, ay), XSUB(bx, ax)));
}

HFR_HD float hfr_seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
  const float bax = XSUB(bx, ax), bay = XSUB(by, ay);
  const float l2 = XADD(XMUL(bax, bax), XMUL(bay, bay));
  if (l2 <= HFR_KEPS) {
    const float ex = XSUB(px, bx), ey = XSUB(py, by);
    return XADD(XMUL(ex, ex), XMUL(ey, ey));
  }
  float t = XDIV(XADD(XMUL(bax, XSUB(px, ax)), XMUL(bay, XSUB(py, ay))), l2);
  t = hfr_clamp01(t);
  const float qx = XADD(ax, XMUL(t, bax)), qy = XADD(ay, XMUL(t, bay));
  const float dx = XSUB(qx, px), dy = XSUB(qy, py);
  return XADD(XMUL(dx, dx), XMUL(dy, dy));
}

// Exact early rejection for blur_radius == 0 (only a pixel INSIDE the face can be rasterised): true when the pixel is
// certainly not inside.  b_i = e_i / area (IEEE division) is <= 0 - or NaN - as soon as e_i is zero, NaN, or differs in
// sign from a non-zero area (a quotient never changes sign; it may underflow to +-0, which is not > 0 either), and the
// perspective-corrected coordinate b_i z_j z_k / max(sum, eps) with z > 0 keeps that sign.  False means "run the exact
// math" (hfr_raster_bary decides), never "inside".  area as hfr_raster_bary takes it; v = the 9 packed face floats.
HFR_HD bool hfr_edge_sign_outside(float px, float py, float x0, float y0, float x1, float y1, float x2, float y2, float area) {
  const float e0 = hfr_edge(px, py, x1, y1, x2, y2), e1 = hfr_edge(px, py, x2, y2, x0, y0), e2 = hfr_edge(px, py, x0, y0, x1, y1);
  const bool pos = area > 0.0f;
  return (area > 0.0f || area < 0.0f) &&
         (!(e0 > 0.0f || e0 < 0.0f) || !(e1 > 0.0f || e1 < 0.0f) || !(e2 > 0.0f || e2 < 0.0f) ||
          (e0 > 0.0f) != pos || (e1 > 0.0f) != pos || (e2 > 0.0f) != pos);
}

// Face-only validity (independent of the pixel): z in front, non-degenerate, not culled.
HFR_HD bool hfr_face_valid(const float* v, int cull_backfaces) {
  const float zmin = hfr_min3(v[2], v[5], v[8]);
  if (zmin < HFR_KEPS) return false;
  const float fa = hfr_edge(v[0], v[1], v[3], v[4], v[6], v[7]);
  if (fa <= HFR_KEPS && fa >= -HFR_KEPS) return false;
  if (cull_backfaces && fa < 0.0f) return false;
  if (!(fa == fa)) return false;  // NaN
  return true;
}

// Barycentrics / depth part.  Returns false when the candidate is rejected before the
// distance test.  `inside` uses the UNCLIPPED perspective-corrected barycentrics.
HFR_HD bool hfr_raster_bary(float px, float py, const float* v, float area, int pc, int clip, float* pz, float* bc,
                            bool* inside) {
  const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7], z2 = v[8];
  float b0 = XDIV(hfr_edge(px, py, x1, y1, x2, y2), area);
  float b1 = XDIV(hfr_edge(px, py, x2, y2, x0, y0), area);
  float b2 = XDIV(hfr_edge(px, py, x0, y0, x1, y1), area);
  if (pc) {
    const float t0 = XMUL(XMUL(b0, z1), z2), t1 = XMUL(XMUL(z0, b1), z2), t2 = XMUL(XMUL(z0, z1), b2);
    const float den = fmaxf(XADD(XADD(t0, t1), t2), HFR_KEPS);
    b0 = XDIV(t0, den); b1 = XDIV(t1, den); b2 = XDIV(t2, den);
  }
  *inside = b0 > 0.0f && b1 > 0.0f && b2 > 0.0f;
  float c0 = b0, c1 = b1, c2 = b2;
  if (clip) {
    c0 = hfr_clamp01(b0); c1 = hfr_clamp01(b1); c2 = hfr_clamp01(b2);
    const float s = fmaxf(XADD(XADD(c0, c1), c2), 1e-5f);
    // a clipped coordinate is very often exactly 0, and 0/s would take div.rn's slow path (FCHK): divide a
    // stand-in instead and select the exact +0 afterwards (s > 0, so 0/s = +0)
    const float q0 = XDIV(c0 == 0.0f ? 1.0f : c0, s), q1 = XDIV(c1 == 0.0f ? 1.0f : c1, s),
                q2 = XDIV(c2 == 0.0f ? 1.0f : c2, s);
    c0 = c0 == 0.0f ? 0.0f : q0; c1 = c1 == 0.0f ? 0.0f : q1; c2 = c2 == 0.0f ? 0.0f : q2;
  }
  bc[0] = c0; bc[1] = c1; bc[2] = c2;
  *pz = XADD(XADD(XMUL(c0, z0), XMUL(c1, z1)), XMUL(c2, z2));
  return !(*pz < 0.0f);
}

HFR_HD float hfr_tri_dist2(float px, float py, const float* v) {
  const float e01 = hfr_seg_dist2(px, py, v[0], v[1], v[3], v[4]);
  const float e02 = hfr_seg_dist2(px, py, v[0], v[1], v[6], v[7]);
  const float e12 = hfr_seg_dist2(px, py, v[3], v[4], v[6], v[7]);
  return fminf(fminf(e01, e02), e12);
}

// Full evaluation from the 9 packed face floats (used by the epilogue and the host emulation).
HFR_HD bool hfr_raster_eval(float px, float py, const float* v, float blur_radius, float sqrt_blur, int pc, int clip,
                            int cull, float* pz, float* bc, float* sdist) {
  if (!hfr_face_valid(v, cull)) return false;
  const float xmin = XSUB(hfr_min3(v[0], v[3], v[6]), sqrt_blur), xmax = XADD(hfr_max3(v[0], v[3], v[6]), sqrt_blur);
  const float ymin = XSUB(hfr_min3(v[1], v[4], v[7]), sqrt_blur), ymax = XADD(hfr_max3(v[1], v[4], v[7]), sqrt_blur);
  if (px < xmin || px > xmax || py < ymin || py > ymax) return false;
  const float area = XADD(hfr_edge(v[6], v[7], v[0], v[1], v[3], v[4]), HFR_KEPS);
  bool inside;
  if (!hfr_raster_bary(px, py, v, area, pc, clip, pz, bc, &inside)) return false;
  const float d = hfr_tri_dist2(px, py, v);
  if (!inside && d >= blur_radius) return false;
  *sdist = inside ? -d : d;
  return true;
}

// ------------------------------------------------------------------------------------ backward
// Plain-arithmetic variants for the backward (FMA contraction allowed, no IEEE division): gradients need fp32
// accuracy, not the forward's bit pattern.
HFR_HD float hfr_edge_f(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}
HFR_HD float hfr_seg_dist2_f(float px, float py, float ax, float ay, float bx, float by) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= HFR_KEPS) {
    const float ex = px - bx, ey = py - by;
    return ex * ex + ey * ey;
  }
  const float t = hfr_clamp01(HFR_FDIV(bax * (px - ax) + bay * (py - ay), l2));
  const float dx = ax + t * bax - px, dy = ay + t * bay - py;
  return dx * dx + dy * dy;
}

// d(seg_dist2)/d(a,b) accumulated into ga[2], gb[2] with upstream g.
HFR_HD void hfr_seg_dist2_bwd(float px, float py, float ax, float ay, float bx, float by, float g, float* ga,
                              float* gb) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= HFR_KEPS) {
    gb[0] += -2.0f * (px - bx) * g;
    gb[1] += -2.0f * (py - by) * g;
    return;
  }
  const float pax = px - ax, pay = py - ay;
  const float t = HFR_FDIV(bax * pax + bay * pay, l2);
  const float tt = hfr_clamp01(t);
  const float qx = ax + tt * bax, qy = ay + tt * bay;
  const float gqx = 2.0f * (qx - px) * g, gqy = 2.0f * (qy - py) * g;
  float gax = gqx, gay = gqy;                       // q = a + tt*ba
  float gbax = tt * gqx, gbay = tt * gqy;
  const float gtt = gqx * bax + gqy * bay;
  const float gt = (t >= 0.0f && t <= 1.0f) ? gtt : 0.0f;
  const float gnum = HFR_FDIV(gt, l2), gl2 = -gnum * t;
  gbax += gnum * pax + 2.0f * gl2 * bax;
  gbay += gnum * pay + 2.0f * gl2 * bay;
  gax -= gnum * bax; gay -= gnum * bay;             // pa = p - a
  ga[0] += gax - gbax; ga[1] += gay - gbay;         // ba = b - a
  gb[0] += gbax; gb[1] += gbay;
}

// Gradients of (bary_clip, pz, signed dist) wrt the 9 face floats, accumulated into gv[9].
// Pixel position is a constant; clamps / max / min are sub-gradients as autograd takes them
// on the forward formulas.
HFR_HD void hfr_raster_eval_bwd(float px, float py, const float* v, int pc, int clip, const float* g_bc, float g_pz,
                                float g_sd, float* gv) {
  const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7], z2 = v[8];
  const float area = hfr_edge_f(x2, y2, x0, y0, x1, y1) + HFR_KEPS;
  const float E0 = hfr_edge_f(px, py, x1, y1, x2, y2), E1 = hfr_edge_f(px, py, x2, y2, x0, y0),
              E2 = hfr_edge_f(px, py, x0, y0, x1, y1);
  const float ia = HFR_RCP(area);
  const float w0 = E0 * ia, w1 = E1 * ia, w2 = E2 * ia;
  float b0 = w0, b1 = w1, b2 = w2, t0 = 0.f, t1 = 0.f, t2 = 0.f, tsum = 0.f, den = 1.f;
  if (pc) {
    t0 = w0 * z1 * z2; t1 = z0 * w1 * z2; t2 = z0 * z1 * w2;
    tsum = t0 + t1 + t2;
    den = fmaxf(tsum, HFR_KEPS);
    const float id = HFR_RCP(den);
    b0 = t0 * id; b1 = t1 * id; b2 = t2 * id;
  }
  const bool inside = b0 > 0.0f && b1 > 0.0f && b2 > 0.0f;
  float c0 = b0, c1 = b1, c2 = b2, csum = 1.f, s = 1.f;
  if (clip) {
    c0 = hfr_clamp01(b0); c1 = hfr_clamp01(b1); c2 = hfr_clamp01(b2);
    csum = c0 + c1 + c2;
    s = fmaxf(csum, 1e-5f);
  }
  const float is = HFR_RCP(s);
  const float bc0 = c0 * is, bc1 = c1 * is, bc2 = c2 * is;
  // pz = sum bc_i z_i
  float gbc0 = g_bc[0] + g_pz * z0, gbc1 = g_bc[1] + g_pz * z1, gbc2 = g_bc[2] + g_pz * z2;
  float gz0 = g_pz * bc0, gz1 = g_pz * bc1, gz2 = g_pz * bc2;
  float gb0 = gbc0, gb1 = gbc1, gb2 = gbc2;
  if (clip) {
    const float gs = (csum >= 1e-5f) ? -(gbc0 * c0 + gbc1 * c1 + gbc2 * c2) * (is * is) : 0.0f;
    const float gc0 = gbc0 * is + gs, gc1 = gbc1 * is + gs, gc2 = gbc2 * is + gs;
    gb0 = (b0 >= 0.0f && b0 <= 1.0f) ? gc0 : 0.0f;
    gb1 = (b1 >= 0.0f && b1 <= 1.0f) ? gc1 : 0.0f;
    gb2 = (b2 >= 0.0f && b2 <= 1.0f) ? gc2 : 0.0f;
  }
  float gw0 = gb0, gw1 = gb1, gw2 = gb2;
  if (pc) {
    const float id = HFR_RCP(den);
    const float gden = (tsum >= HFR_KEPS) ? -(gb0 * t0 + gb1 * t1 + gb2 * t2) * (id * id) : 0.0f;
    const float gt0 = gb0 * id + gden, gt1 = gb1 * id + gden, gt2 = gb2 * id + gden;
    gw0 = gt0 * z1 * z2; gz1 += gt0 * w0 * z2; gz2 += gt0 * w0 * z1;
    gw1 = gt1 * z0 * z2; gz0 += gt1 * w1 * z2; gz2 += gt1 * z0 * w1;
    gw2 = gt2 * z0 * z1; gz0 += gt2 * z1 * w2; gz1 += gt2 * z0 * w2;
  }
  // w_i = E_i / area
  const float gE0 = gw0 * ia, gE1 = gw1 * ia, gE2 = gw2 * ia;
  const float garea = -(gw0 * w0 + gw1 * w1 + gw2 * w2) * ia;
  float gx0 = 0.f, gy0 = 0.f, gx1 = 0.f, gy1 = 0.f, gx2 = 0.f, gy2 = 0.f;
  // E(p;a,b): dE/dax = py-by, dE/day = bx-px, dE/dbx = -(py-ay), dE/dby = px-ax
  // E0 = E(p; v1, v2)
  gx1 += gE0 * (py - y2); gy1 += gE0 * (x2 - px); gx2 += gE0 * -(py - y1); gy2 += gE0 * (px - x1);
  // E1 = E(p; v2, v0)
  gx2 += gE1 * (py - y0); gy2 += gE1 * (x0 - px); gx0 += gE1 * -(py - y2); gy0 += gE1 * (px - x2);
  // E2 = E(p; v0, v1)
  gx0 += gE2 * (py - y1); gy0 += gE2 * (x1 - px); gx1 += gE2 * -(py - y0); gy1 += gE2 * (px - x0);
  // area = E(v2; v0, v1): p = v2 also moves: dE/dpx = by-ay, dE/dpy = -(bx-ax)
  gx0 += garea * (y2 - y1); gy0 += garea * (x1 - x2); gx1 += garea * -(y2 - y0); gy1 += garea * (x2 - x0);
  gx2 += garea * (y1 - y0); gy2 += garea * -(x1 - x0);
  // signed distance: first-match priority e01, e02, e12
  if (g_sd != 0.0f) {
    const float e01 = hfr_seg_dist2_f(px, py, x0, y0, x1, y1), e02 = hfr_seg_dist2_f(px, py, x0, y0, x2, y2),
                e12 = hfr_seg_dist2_f(px, py, x1, y1, x2, y2);
    const float g = inside ? -g_sd : g_sd;
    float ga[2] = {0.f, 0.f}, gb[2] = {0.f, 0.f};
    if (e01 <= e02 && e01 <= e12) {
      hfr_seg_dist2_bwd(px, py, x0, y0, x1, y1, g, ga, gb);
      gx0 += ga[0]; gy0 += ga[1]; gx1 += gb[0]; gy1 += gb[1];
    } else if (e02 <= e01 && e02 <= e12) {
      hfr_seg_dist2_bwd(px, py, x0, y0, x2, y2, g, ga, gb);
      gx0 += ga[0]; gy0 += ga[1]; gx2 += gb[0]; gy2 += gb[1];
    } else {
      hfr_seg_dist2_bwd(px, py, x1, y1, x2, y2, g, ga, gb);
      gx1 += ga[0]; gy1 += ga[1]; gx2 += gb[0]; gy2 += gb[1];
    }
  }
  gv[0] += gx0; gv[1] += gy0; gv[2] += gz0;
  gv[3] += gx1; gv[4] += gy1; gv[5] += gz1;
  gv[6] += gx2; gv[7] += gy2; gv[8] += gz2;
}

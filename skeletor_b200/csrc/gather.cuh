// gather.cuh — CIC / TSC field gather from a shared-memory window (or HBM).
// Restates gather_cic / gather_tsc, reference skeletor/cython/particle_push.pxd:3-67.
#pragma once
#include "common.cuh"

// Load the NS x NS stencil of a Float3 field around (ix, iy) [lower-left cell for
// CIC, centre cell for TSC] from the shared window or, failing that, from HBM.
template <int NS>
__device__ __forceinline__ void load_stencil(const double *__restrict__ sw,
                                             const Window &w, int wstride,
                                             const double *__restrict__ F,
                                             const DevGrid &g, int ix, int iy,
                                             double (&v)[NS * NS][3]) {
  const int lo = (NS == 3) ? 1 : 0;
  const int x_lo = ix - lo, y_lo = iy - lo;
  // x_lo in [w.x0, w.x1 - NS] and y_lo in [w.y0, w.y1 - NS]: one unsigned compare each
  const unsigned rx = (unsigned)(x_lo - w.x0), ry = (unsigned)(y_lo - w.y0);
  if (rx <= (unsigned)(w.x1 - w.x0 - NS) && ry <= (unsigned)(w.y1 - w.y0 - NS) &&
      w.x1 - w.x0 >= NS && w.y1 - w.y0 >= NS) {
    const double *b = sw + ((size_t)ry * wstride + rx) * 3;
#pragma unroll
    for (int a = 0; a < NS; a++)
#pragma unroll
      for (int c = 0; c < NS; c++)
#pragma unroll
        for (int k = 0; k < 3; k++) v[a * NS + c][k] = b[(a * wstride + c) * 3 + k];
  } else if (x_lo >= 0 && x_lo + NS <= g.mx && y_lo >= 0 && y_lo + NS <= g.myp) {
    const double *b = F + ((size_t)y_lo * g.mx + x_lo) * 3;
#pragma unroll
    for (int a = 0; a < NS; a++)
#pragma unroll
      for (int c = 0; c < NS; c++)
#pragma unroll
        for (int k = 0; k < 3; k++) v[a * NS + c][k] = __ldg(b + ((size_t)a * g.mx + c) * 3 + k);
  } else {
    // outside the array: the reference would read out of bounds; read zeros
#pragma unroll
    for (int a = 0; a < NS * NS; a++)
#pragma unroll
      for (int k = 0; k < 3; k++) v[a][k] = 0.0;
  }
}

// gather_cic, particle_push.pxd:3-27
__device__ __forceinline__ void gather_cic(const double *sw, const Window &w, int ws,
                                           const double *F, const DevGrid &g,
                                           double xs, double ys, double (&f)[3]) {
  int ix, iy;
  double dx, tx, dy, ty;
  cic_weights(xs, ix, dx, tx);
  cic_weights(ys, iy, dy, ty);
  double v[4][3];
  load_stencil<2>(sw, w, ws, F, g, ix, iy, v);
#pragma unroll
  for (int k = 0; k < 3; k++)
    f[k] = dy * (dx * v[3][k] + tx * v[2][k]) + ty * (dx * v[1][k] + tx * v[0][k]);
}

// gather_tsc, particle_push.pxd:29-67
__device__ __forceinline__ void gather_tsc(const double *sw, const Window &w, int ws,
                                           const double *F, const DevGrid &g,
                                           double xs, double ys, double (&f)[3]) {
  int ix, iy;
  double wmx, w0x, wpx, wmy, w0y, wpy;
  tsc_weights(xs + 0.5, ix, wmx, w0x, wpx);
  tsc_weights(ys + 0.5, iy, wmy, w0y, wpy);
  double v[9][3];
  load_stencil<3>(sw, w, ws, F, g, ix, iy, v);
#pragma unroll
  for (int k = 0; k < 3; k++)
    f[k] = wmy * (wmx * v[0][k] + w0x * v[1][k] + wpx * v[2][k]) +
           w0y * (wmx * v[3][k] + w0x * v[4][k] + wpx * v[5][k]) +
           wpy * (wmx * v[6][k] + w0x * v[7][k] + wpx * v[8][k]);
}

// Copy the window rows of a Float3 field into shared memory: one warp per row,
// lanes along the row (coalesced, no integer division, all loads independent so the
// whole window is in flight at once).
__device__ __forceinline__ void stage_window(double *sw, const double *__restrict__ F,
                                             const Window &w, int wstride,
                                             const DevGrid &g) {
  const int wx = (w.x1 - w.x0) * 3, wy = w.y1 - w.y0;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = wv; r < wy; r += nw) {
    const double *src = F + ((size_t)(w.y0 + r) * g.mx + w.x0) * 3;
    double *dst = sw + (size_t)r * wstride * 3;
#pragma unroll 3
    for (int c = lane; c < wx; c += 32) dst[c] = __ldg(src + c);
  }
}

struct KickParams {
  double qtmh, dt, Omega, S;
  double offBx, offBy, offEx, offEy;  // particle_push.pyx:15-19
};

static inline KickParams make_kick(const DevGrid &g, double qtmh, double dt,
                                   double Omega, double S) {
  KickParams k;
  k.qtmh = qtmh; k.dt = dt; k.Omega = Omega; k.S = S;
  k.offBx = (double)g.lbx;
  k.offBy = (double)(g.lby - g.noff);
  k.offEx = k.offBx - 0.5;
  k.offEy = k.offBy - 0.5;
  return k;
}

// rescale (particle_push.pxd:93-97), [rotation / shear terms, particle_push.pyx:75-76],
// Boris kick (kick_particle, particle_push.pxd:69-86).  `y` is the position BEFORE the
// drift.
template <bool MODIFIED>
__device__ __forceinline__ void rescale_and_kick(double (&e)[3], double (&b)[3],
                                                 const DevGrid &g, const KickParams &q,
                                                 double y, double &vx, double &vy,
                                                 double &vz) {
#pragma unroll
  for (int k = 0; k < 3; k++) { e[k] = e[k] * q.qtmh; b[k] = b[k] * q.qtmh; }
  if (MODIFIED) {  // particle_push.pyx:75-76
    b[2] = b[2] + q.Omega * q.dt;
    e[1] = e[1] - q.S * (g.y0 + y * g.dy) * b[2];
  }
  const double vmx = vx + e[0], vmy = vy + e[1], vmz = vz + e[2];
  const double vpx = vmx + (vmy * b[2] - vmz * b[1]);
  const double vpy = vmy + (vmz * b[0] - vmx * b[2]);
  const double vpz = vmz + (vmx * b[1] - vmy * b[0]);
  const double fac = 2. / (1. + b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  vx = vmx + fac * (vpy * b[2] - vpz * b[1]) + e[0];
  vy = vmy + fac * (vpz * b[0] - vpx * b[2]) + e[1];
  vz = vmz + fac * (vpx * b[1] - vpy * b[0]) + e[2];
}

// gather E and B at (x, y), rescale, [rotation/shear terms], Boris kick.
// particle_push.pyx:28-36 (+ :75-76) and kick_particle, particle_push.pxd:69-86
template <int ORDER, bool MODIFIED>
__device__ __forceinline__ void fields_and_kick(const double *sE, const double *sB,
                                                const Window &w, int ws,
                                                const double *E, const double *B,
                                                const DevGrid &g, const KickParams &q,
                                                double x, double y, double &vx,
                                                double &vy, double &vz) {
  double e[3], b[3];
  if (ORDER == 1) {
    gather_cic(sE, w, ws, E, g, x + q.offEx, y + q.offEy, e);
    gather_cic(sB, w, ws, B, g, x + q.offBx, y + q.offBy, b);
  } else {
    gather_tsc(sE, w, ws, E, g, x + q.offEx, y + q.offEy, e);
    gather_tsc(sB, w, ws, B, g, x + q.offBx, y + q.offBy, b);
  }
  rescale_and_kick<MODIFIED>(e, b, g, q, y, vx, vy, vz);
}

// deposit.cuh — device helpers shared by the deposition kernels (deposit.cu) and the
// fused push_and_deposit on the gapped layout (gapped.cu): per-cell register
// accumulators, warp reductions, shared-memory window emit / flush.
// Restates deposit_particle_cic/tsc, reference skeletor/cython/deposit.pxd:3-118.
#pragma once
#include "common.cuh"
#include "gather.cuh"
#include <limits.h>

#define ACC_EMPTY INT_MIN  // Acc.iy of a lane that holds no data

#define DEP_THREADS 256

struct DepParams {
  double offx, offy;  // deposit.pyx:14-15
  double S;
};

template <int NS>
struct Acc {
  double v[NS * NS * 4];
  int ix, iy;  // stencil base cell (lower-left for CIC, centre for TSC); iy == ACC_EMPTY: no data
};

// Add the per-run totals to the shared window (or HBM).  Called by run tails only.
template <int NS>
__device__ __forceinline__ void emit(const Acc<NS> &a, double *sw, const Window &w,
                                     int wstride, double *__restrict__ cur,
                                     const DevGrid &g) {
  const int lo = (NS == 3) ? 1 : 0;
  const int x_lo = a.ix - lo, y_lo = a.iy - lo;
  if (x_lo >= w.x0 && x_lo + NS <= w.x1 && y_lo >= w.y0 && y_lo + NS <= w.y1) {
    double *b = sw + ((size_t)(y_lo - w.y0) * wstride + (x_lo - w.x0)) * 4;
#pragma unroll
    for (int r = 0; r < NS; r++)
#pragma unroll
      for (int c = 0; c < NS; c++)
#pragma unroll
        for (int k = 0; k < 4; k++)
          atomicAdd(b + (r * wstride + c) * 4 + k, a.v[(r * NS + c) * 4 + k]);
  } else if (x_lo >= 0 && x_lo + NS <= g.mx && y_lo >= 0 && y_lo + NS <= g.myp) {
    double *b = cur + ((size_t)y_lo * g.mx + x_lo) * 4;
#pragma unroll
    for (int r = 0; r < NS; r++)
#pragma unroll
      for (int c = 0; c < NS; c++)
#pragma unroll
        for (int k = 0; k < 4; k++)
          atomicAdd(b + ((size_t)r * g.mx + c) * 4 + k, a.v[(r * NS + c) * 4 + k]);
  }
  // else: outside the array (the reference would write out of bounds) — dropped
}

// Segmented reduction of every lane's accumulator over runs of equal cell, then
// the run tails emit.  All 32 lanes must call this.
template <int NS>
__device__ __forceinline__ void warp_flush(Acc<NS> &a, double *sw, const Window &w,
                                           int wstride, double *cur, const DevGrid &g) {
  const int lane = threadIdx.x & 31;
  const bool has = a.iy != ACC_EMPTY;
  const int pix = __shfl_up_sync(SKB_FULL, a.ix, 1);
  const int piy = __shfl_up_sync(SKB_FULL, a.iy, 1);
  const bool head = (lane == 0) || (pix != a.ix) || (piy != a.iy);
  const unsigned heads = __ballot_sync(SKB_FULL, head);
  if (heads == 1u) {
    // whole warp in one cell (the common case): plain butterfly-free reduction
    if (has) {
#pragma unroll
      for (int i = 0; i < NS * NS * 4; i++) {
        double s = a.v[i];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) s += __shfl_down_sync(SKB_FULL, s, d);
        a.v[i] = s;
      }
      if (lane == 0) emit<NS>(a, sw, w, wstride, cur, g);
    }
  } else {
    // start lane of my run = highest head at or below my lane
    const int seg0 = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
    for (int i = 0; i < NS * NS * 4; i++) {
      double s = a.v[i];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double t = __shfl_up_sync(SKB_FULL, s, d);
        if (lane - d >= seg0) s += t;
      }
      a.v[i] = s;
    }
    const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
    if (tail && has) emit<NS>(a, sw, w, wstride, cur, g);
  }
#pragma unroll
  for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
  a.iy = ACC_EMPTY;
  a.ix = 0;
}

// Accumulate one particle (already positioned: xs = x + offx [+0.5 for TSC]).
// deposit_particle_cic, deposit.pxd:3-44 / deposit_particle_tsc, deposit.pxd:46-118
template <int ORDER>
__device__ __forceinline__ void particle_terms(double xs, double ys, int &ix, int &iy,
                                               double (&wx)[ORDER + 1],
                                               double (&wy)[ORDER + 1]) {
  if (ORDER == 1) {
    double d, t;
    cic_weights(xs, ix, d, t); wx[0] = t; wx[1] = d;
    cic_weights(ys, iy, d, t); wy[0] = t; wy[1] = d;
  } else {
    tsc_weights(xs, ix, wx[0], wx[1], wx[2]);
    tsc_weights(ys, iy, wy[0], wy[1], wy[2]);
  }
}

template <int ORDER>
__device__ __forceinline__ void accumulate(Acc<ORDER + 1> &a, const double (&wx)[ORDER + 1],
                                           const double (&wy)[ORDER + 1], double vxr,
                                           double vy, double vz) {
  constexpr int NS = ORDER + 1;
#pragma unroll
  for (int r = 0; r < NS; r++)
#pragma unroll
    for (int c = 0; c < NS; c++) {
      const double wgt = wy[r] * wx[c];
      double *v = a.v + (r * NS + c) * 4;
      v[0] += wgt;
      v[1] += wgt * vxr;
      v[2] += wgt * vy;
      v[3] += wgt * vz;
    }
}

__device__ __forceinline__ void zero_window(double *sw, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) sw[i] = 0.0;
}

// add the shared window to the HBM source grid; zero entries (halo cells that no
// particle touched) are skipped
__device__ __forceinline__ void flush_window(const double *sw, const Window &w,
                                             int wstride, double *__restrict__ cur,
                                             const DevGrid &g) {
  const int wx = (w.x1 - w.x0) * 4, wy = w.y1 - w.y0;
  for (int idx = threadIdx.x; idx < wx * wy; idx += blockDim.x) {
    int r = idx / wx, c = idx - r * wx;
    double v = sw[(size_t)r * wstride * 4 + c];
    if (v != 0.0) atomicAdd(cur + ((size_t)(w.y0 + r) * g.mx + w.x0) * 4 + c, v);
  }
}

// MODE 0: deposit only (deposit.pyx:6-34)
// MODE 1: push_and_deposit, update = False (predictor: particles untouched)
// MODE 2: push_and_deposit, update = True            (push_and_deposit.pyx:10-170)
struct FusedParams {
  KickParams k;
  double d2x, d2y;  // 0.5*dt/dx, 0.5*dt/dy, push_and_deposit.pyx:37-38
  int *ihole;
  int ntmax;
};

#define SKB_CFL_BIT 0x40000000

#ifndef DEP_PREFETCH_LINES
#define DEP_PREFETCH_LINES 16   // 16 lines x 16 doubles = the next 256 particles
#endif
__device__ __forceinline__ void dep_prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int ORDER> struct DepUnroll { static constexpr int value = (ORDER == 1) ? 4 : 2; };

template <int NS>
__device__ __forceinline__ void single_particle_emit(const double (&wx)[NS],
                                                     const double (&wy)[NS], int ix, int iy,
                                                     double vxr, double vy, double vz,
                                                     double *__restrict__ cur,
                                                     const DevGrid &g) {
  const int lo = (NS == 3) ? 1 : 0;
  const int x_lo = ix - lo, y_lo = iy - lo;
  if (!(x_lo >= 0 && x_lo + NS <= g.mx && y_lo >= 0 && y_lo + NS <= g.myp)) return;
  double *b = cur + ((size_t)y_lo * g.mx + x_lo) * 4;
#pragma unroll
  for (int r = 0; r < NS; r++)
#pragma unroll
    for (int c = 0; c < NS; c++) {
      const double wgt = wy[r] * wx[c];
      double *v = b + ((size_t)r * g.mx + c) * 4;
      atomicAdd(v + 0, wgt);
      atomicAdd(v + 1, wgt * vxr);
      atomicAdd(v + 2, wgt * vy);
      atomicAdd(v + 3, wgt * vz);
    }
}

// Warp reduce-scatter of V (power of two <= 32) per-lane values: V-1 + (5 - log2 V)
// shuffle exchanges instead of 5 V.  On return v[0] of lane l holds the sum over all
// 32 lanes of value number scatter_index<V>(l); lanes l and l + V hold the same.
template <int V>
__device__ __forceinline__ int scatter_index(int lane) {
  int idx = 0;
#pragma unroll
  for (int s = 0, h = V / 2; h >= 1; s++, h >>= 1) idx += ((lane >> s) & 1) * h;
  return idx;
}

// HALF: the two half-warps reduce independently (V = 16 values over 16 lanes each): used
// when every half-warp accumulates a cell of its own.
template <int V, bool HALF = false>
__device__ __forceinline__ void warp_reduce_scatter(double *v, int lane) {
  int bit = 1;
#pragma unroll
  for (int h = V / 2; h >= 1; h >>= 1, bit <<= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int j = 0; j < h; j++) {
      const double a = v[j], b = v[j + h];
      const double send = up ? a : b;
      const double keep = up ? b : a;
      v[j] = keep + __shfl_xor_sync(SKB_FULL, send, bit);
    }
  }
#pragma unroll
  for (; bit < (HALF ? 16 : 32); bit <<= 1) v[0] += __shfl_xor_sync(SKB_FULL, v[0], bit);
}

// add value number idx (= (r*NS + c)*4 + k) of a cell's stencil sums to the window
template <int NS>
__device__ __forceinline__ void emit_one(double val, int idx, bool in_window, int ix, int iy,
                                         double *sw, const Window &w, int wstride,
                                         double *__restrict__ cur, const DevGrid &g) {
  const int lo = (NS == 3) ? 1 : 0;
  const int cell = idx >> 2, k = idx & 3;
  const int r = cell / NS, c = cell - r * NS;
  const int x = ix - lo + c, y = iy - lo + r;
  if (in_window)
    atomicAdd(sw + ((size_t)(y - w.y0) * wstride + (x - w.x0)) * 4 + k, val);
  else if (x >= 0 && x < g.mx && y >= 0 && y < g.myp)
    atomicAdd(cur + ((size_t)y * g.mx + x) * 4 + k, val);
}

template <int NS>
__device__ __forceinline__ void stray_particle_emit(const double (&wx)[NS],
                                                    const double (&wy)[NS], int ix, int iy,
                                                    double vxr, double vy, double vz,
                                                    double *sw, const Window &w, int wstride,
                                                    double *__restrict__ cur,
                                                    const DevGrid &g) {
  const int lo = (NS == 3) ? 1 : 0;
  const int x_lo = ix - lo, y_lo = iy - lo;
  if (x_lo >= w.x0 && x_lo + NS <= w.x1 && y_lo >= w.y0 && y_lo + NS <= w.y1) {
    double *b = sw + ((size_t)(y_lo - w.y0) * wstride + (x_lo - w.x0)) * 4;
#pragma unroll
    for (int r = 0; r < NS; r++)
#pragma unroll
      for (int c = 0; c < NS; c++) {
        const double wgt = wy[r] * wx[c];
        double *v = b + (r * wstride + c) * 4;
        atomicAdd(v + 0, wgt);
        atomicAdd(v + 1, wgt * vxr);
        atomicAdd(v + 2, wgt * vy);
        atomicAdd(v + 3, wgt * vz);
      }
  } else {
    single_particle_emit<NS>(wx, wy, ix, iy, vxr, vy, vz, cur, g);
  }
}


// common.cuh — shared device helpers for libskeletor_b200 (sm_100a, float64).
//
// Everything that touches particle coordinates keeps the reference's operation
// order and the translation unit is compiled with
// -fmad=false, so per-particle results are bit-identical to the reference's
// gcc -O2 (SSE2, no FMA) code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "skeletor_b200.h"

#define SKB_FULL 0xffffffffu

// Halo of the shared-memory field / source window around a tile of stencil-base
// cells.  Low side: TSC reaches base-1, +1 cell of slack for particles that have
// moved since the last sort.  High side: the B gather (offset lbx instead of
// lbx-1/2) reaches base+2, +1 cell of slack.
#define SKB_HALO_LO 2
#define SKB_HALO_HI 3

struct DevGrid {
  int nx, ny, nyp, noff, lbx, lby, ubx, uby, mx, myp;
  double dx, dy, Lx, Ly, x0, y0, e0, e1;
};

static inline DevGrid make_grid(const skb_grid_t *g) {
  DevGrid d;
  d.nx = g->nx; d.ny = g->ny; d.nyp = g->nyp; d.noff = g->noff;
  d.lbx = g->lbx; d.lby = g->lby; d.ubx = g->ubx; d.uby = g->uby;
  d.mx = g->nx + 2 * g->lbx; d.myp = g->nyp + 2 * g->lby;
  d.dx = g->dx; d.dy = g->dy; d.Lx = g->Lx; d.Ly = g->Ly;
  d.x0 = g->x0; d.y0 = g->y0; d.e0 = g->edges[0]; d.e1 = g->edges[1];
  return d;
}

struct DevTiling {
  const int *tile_offsets;
  const int *chunk_first_tile;
  const int *cell_end;
  const int *gap_start, *gap_count;
  int ntx, nty, tlx, tly, chunk;
  long long n_sorted;
};

static inline DevTiling make_tiling(const skb_tiling_t *t) {
  DevTiling d;
  if (t && (t->tile_offsets || t->gap_start)) {
    d.tile_offsets = t->tile_offsets; d.chunk_first_tile = t->chunk_first_tile;
    d.cell_end = t->cell_end;
    d.gap_start = t->gap_start; d.gap_count = t->gap_count;
    d.ntx = t->ntx; d.nty = t->nty; d.tlx = t->tlx; d.tly = t->tly;
    d.chunk = t->chunk; d.n_sorted = t->n_sorted;
  } else {
    d.tile_offsets = nullptr; d.chunk_first_tile = nullptr; d.cell_end = nullptr;
    d.gap_start = nullptr; d.gap_count = nullptr;
    d.ntx = d.nty = 1; d.tlx = d.tly = 4; d.chunk = 2048; d.n_sorted = 0;
  }
  return d;
}

// Window of the extended [myp][mx] array staged in shared memory for one tile.
struct Window {
  int x0, y0, x1, y1;  // [x0,x1) x [y0,y1) in array index space; empty if x1<=x0
};

__device__ __forceinline__ Window tile_window(int tile, const DevTiling &t,
                                              const DevGrid &g) {
  Window w;
  if (tile < 0) { w.x0 = w.y0 = 0; w.x1 = w.y1 = 0; return w; }
  int tx = tile % t.ntx, ty = tile / t.ntx;
  int bx = tx << t.tlx, by = ty << t.tly;
  w.x0 = max(bx - SKB_HALO_LO, 0);
  w.y0 = max(by - SKB_HALO_LO, 0);
  w.x1 = min(bx + (1 << t.tlx) + SKB_HALO_HI, g.mx);
  w.y1 = min(by + (1 << t.tly) + SKB_HALO_HI, g.myp);
  return w;
}

static inline int window_stride(const DevTiling &t) {
  return (1 << t.tlx) + SKB_HALO_LO + SKB_HALO_HI;
}
static inline int window_rows(const DevTiling &t) {
  return (1 << t.tly) + SKB_HALO_LO + SKB_HALO_HI;
}

// Iterates over the (tile, particle-range) segments of the work item of one CTA:
// particles [first, last) where first = blockIdx.x * span.  Uniform across the CTA.
struct SegmentIter {
  long long p, last, sorted_last;
  int tile;
  __device__ __forceinline__ void init(const DevTiling &t, long long np, int span) {
    p = (long long)blockIdx.x * span;
    last = min(p + (long long)span, np);
    sorted_last = 0;
    tile = -1;
    if (t.tile_offsets && p < t.n_sorted) {
      sorted_last = min(last, t.n_sorted);
      tile = t.chunk_first_tile[p / t.chunk];
    }
  }
  // Returns false when done; otherwise sets [s0,s1) and the tile (-1: unsorted)
  __device__ __forceinline__ bool next(const DevTiling &t, long long &s0,
                                       long long &s1, int &seg_tile) {
    if (p >= last) return false;
    if (p < sorted_last) {
      // skip tiles that end at or before p (empty tiles), 32 at a time
      const int ntiles = t.ntx * t.nty;
      const int lane = threadIdx.x & 31;
      while (true) {
        int cand = tile + lane;
        bool ends_after = (cand < ntiles) ? (t.tile_offsets[cand + 1] > p) : true;
        unsigned m = __ballot_sync(SKB_FULL, ends_after);
        if (m) { tile += __ffs(m) - 1; break; }
        tile += 32;
      }
      long long tend = (tile < ntiles) ? (long long)t.tile_offsets[tile + 1] : sorted_last;
      s0 = p;
      s1 = min(sorted_last, tend);
      seg_tile = (tile < ntiles) ? tile : -1;
      p = s1;
      return true;
    }
    s0 = p; s1 = last; seg_tile = -1; p = last;
    return true;
  }
};

// ---- interpolation weights, reference order -----------------------------------
// CIC: particle_push.pxd:10-21 / deposit.pxd:11-21.  xs = x + offset (done by caller)
__device__ __forceinline__ void cic_weights(double xs, int &i, double &d, double &t) {
  i = (int)xs;              // truncation, as the C cast
  d = xs - (double)i;
  t = 1.0 - d;
}
// TSC: particle_push.pxd:37-57 / deposit.pxd:55-72.  xs = x + offset + 0.5 (caller)
__device__ __forceinline__ void tsc_weights(double xs, int &i, double &wm, double &w0,
                                            double &wp) {
  i = (int)xs;
  double d = xs - (double)i - 0.5;
  w0 = 0.75 - d * d;
  double h = 0.5 + d;
  wp = 0.5 * (h * h);
  wm = 1.0 - (w0 + wp);
}

// periodic_x_cdef, particle_boundary.pxd:3-7 (guarded against non-finite x, for
// which the reference loops forever)
__device__ __forceinline__ double wrap_x(double x, double nx) {
  if (!isfinite(x)) return x;
  while (x < 0.0) x = x + nx;
  while (x >= nx) x = x - nx;
  return x;
}

// ---- sort key (see sort.cu) ----------------------------------------------------
struct KeyParams {
  double offx, offy;
  int order, tlx, tly, ntx, mx, myp;
};

static inline KeyParams make_keyparams(const DevGrid &g, int order, int tlx, int tly) {
  KeyParams k;
  k.offx = g.lbx - 0.5;
  k.offy = g.lby - 0.5 - g.noff;
  k.order = order; k.tlx = tlx; k.tly = tly;
  k.mx = g.mx; k.myp = g.myp;
  k.ntx = (g.mx + (1 << tlx) - 1) >> tlx;
  return k;
}

__device__ __forceinline__ int cell_key(double x, double y, const KeyParams &k) {
  double xs = x + k.offx, ys = y + k.offy;
  if (k.order == 2) { xs = xs + 0.5; ys = ys + 0.5; }
  int ix = (int)xs, iy = (int)ys;
  ix = min(max(ix, 0), k.mx - 1);
  iy = min(max(iy, 0), k.myp - 1);
  const int mxm = (1 << k.tlx) - 1, mym = (1 << k.tly) - 1;
  return ((((iy >> k.tly) * k.ntx + (ix >> k.tlx)) << (k.tlx + k.tly)) |
          ((iy & mym) << k.tlx) | (ix & mxm));
}

#define SKB_CHECK_LAUNCH()                      \
  do {                                          \
    cudaError_t e_ = cudaGetLastError();        \
    if (e_ != cudaSuccess) return (int)e_;      \
  } while (0)

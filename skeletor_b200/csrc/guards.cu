// guards.cu — guard-cell kernels on the extended [myp][mx] grid.
//
// Replaces the NumPy slicing of Field.copy_guards_x/_y (reference
// skeletor/field.py:73-98), Sources.add_guards_x/_y + guard zeroing
// (skeletor/sources.py:91-150) and Sources.normalize's scaling (sources.py:61-63).
// The y-direction neighbour exchange (mpi4py sendrecv, field.py:52-58) is done by
// the caller over NCCL; these kernels consume the received packed rows, or wrap
// periodically inside the slab when there is a single rank.
#include "common.cuh"

#define GT 256

// One thread per guard cell (all components).  Every source is an ACTIVE cell (or a
// received row), never a cell written by this kernel, so y-then-x ordering of the
// reference (corners!) is reproduced without a barrier.
__global__ void __launch_bounds__(GT)
copy_guards_kernel(double *f, int nc, DevGrid g, const double *__restrict__ below,
                   const double *__restrict__ above) {
  long long idx = (long long)blockIdx.x * GT + threadIdx.x;
  if (idx >= (long long)g.mx * g.myp) return;
  int iy = (int)(idx / g.mx), ix = (int)(idx - (long long)iy * g.mx);
  const bool gy = iy < g.lby || iy >= g.uby, gx = ix < g.lbx || ix >= g.ubx;
  if (!gy && !gx) return;
  int sx = ix;                       // source column (active), field.py:79-83
  if (ix < g.lbx) sx = ix + g.nx;
  else if (ix >= g.ubx) sx = ix - g.nx;
  const double *src;
  if (!gy) {
    src = f + ((size_t)iy * g.mx + sx) * nc;
  } else if (iy < g.lby) {           // lower guard row: field.py:91-92, 98
    src = below ? below + ((size_t)iy * g.nx + (sx - g.lbx)) * nc
                : f + ((size_t)(iy + g.nyp) * g.mx + sx) * nc;
  } else {                           // upper guard row: field.py:93-94, 97
    src = above ? above + ((size_t)(iy - g.uby) * g.nx + (sx - g.lbx)) * nc
                : f + ((size_t)(iy - g.nyp) * g.mx + sx) * nc;
  }
  double *dst = f + ((size_t)iy * g.mx + ix) * nc;
  for (int c = 0; c < nc; c++) dst[c] = src[c];
}

// x guards of rows [iy0, iy0+nrows) only (after the spectral remap of a guard row)
__global__ void __launch_bounds__(GT)
copy_guards_x_rows_kernel(double *f, int nc, DevGrid g, int iy0, int nrows) {
  int idx = blockIdx.x * GT + threadIdx.x;
  if (idx >= nrows * 2 * g.lbx) return;
  int r = idx / (2 * g.lbx), k = idx - r * 2 * g.lbx;
  int iy = iy0 + r;
  int ix = (k < g.lbx) ? k : g.ubx + (k - g.lbx);
  int sx = (k < g.lbx) ? ix + g.nx : ix - g.nx;
  for (int c = 0; c < nc; c++)
    f[((size_t)iy * g.mx + ix) * nc + c] = f[((size_t)iy * g.mx + sx) * nc + c];
}

// add_guards_x: sources.py:91-101, all rows.  One thread per (row, component).
__global__ void __launch_bounds__(GT)
add_guards_x_kernel(double *f, int nc, DevGrid g) {
  int idx = blockIdx.x * GT + threadIdx.x;
  if (idx >= g.myp * nc) return;
  int iy = idx / nc, c = idx - iy * nc;
  double *row = f + (size_t)iy * g.mx * nc + c;
  for (int ix = 0; ix < g.lbx; ix++) row[(size_t)(ix + g.nx) * nc] += row[(size_t)ix * nc];
  for (int ix = g.ubx + g.lbx - 1; ix >= g.ubx; ix--)
    row[(size_t)(ix - g.nx) * nc] += row[(size_t)ix * nc];
}

// add_guards_y + zeroing: sources.py:103-115, 147-150.  One thread per (column,
// component); it owns the whole column, so the reference's loop order is kept.
__global__ void __launch_bounds__(GT)
add_guards_y_kernel(double *f, int nc, DevGrid g, const double *__restrict__ from_below,
                    const double *__restrict__ from_above) {
  int idx = blockIdx.x * GT + threadIdx.x;
  if (idx >= g.mx * nc) return;
  int ix = idx / nc, c = idx - ix * nc;
  double *col = f + (size_t)ix * nc + c;
  const size_t rs = (size_t)g.mx * nc;
  if (ix >= g.lbx && ix < g.ubx) {
    // lower guard rows hold what came from the rank above (its lower guards)
    for (int iy = 0; iy < g.lby; iy++) {
      double v = from_above ? from_above[((size_t)iy * g.nx + (ix - g.lbx)) * nc + c]
                            : col[(size_t)iy * rs];
      col[(size_t)(iy + g.nyp) * rs] += v;
    }
    for (int iy = g.uby + g.lby - 1; iy >= g.uby; iy--) {
      double v = from_below ? from_below[((size_t)(iy - g.uby) * g.nx + (ix - g.lbx)) * nc + c]
                            : col[(size_t)iy * rs];
      col[(size_t)(iy - g.nyp) * rs] += v;
    }
    for (int iy = 0; iy < g.lby; iy++) col[(size_t)iy * rs] = 0.0;
    for (int iy = g.uby; iy < g.myp; iy++) col[(size_t)iy * rs] = 0.0;
  } else {
    for (int iy = 0; iy < g.myp; iy++) col[(size_t)iy * rs] = 0.0;
  }
}

__global__ void __launch_bounds__(GT)
pack_rows_kernel(const double *__restrict__ f, int nc, DevGrid g, int iy0, int nrows,
                 double *out) {
  long long idx = (long long)blockIdx.x * GT + threadIdx.x;
  long long n = (long long)nrows * g.nx * nc;
  if (idx >= n) return;
  int r = (int)(idx / ((long long)g.nx * nc));
  long long k = idx - (long long)r * g.nx * nc;
  out[idx] = f[((size_t)(iy0 + r) * g.mx + g.lbx) * nc + k];
}

__global__ void __launch_bounds__(GT)
scale_kernel(double *f, long long n, double fac) {
  long long i = (long long)blockIdx.x * GT + threadIdx.x;
  if (i < n) f[i] = f[i] * fac;
}

static inline unsigned gblk(long long n) { return (unsigned)((n + GT - 1) / GT); }

extern "C" int skb_copy_guards(double *f, int nc, const skb_grid_t *grid,
                               const double *from_below, const double *from_above,
                               void *stream) {
  DevGrid g = make_grid(grid);
  copy_guards_kernel<<<gblk((long long)g.mx * g.myp), GT, 0, (cudaStream_t)stream>>>(
      f, nc, g, from_below, from_above);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_copy_guards_x_rows(double *f, int nc, const skb_grid_t *grid, int iy0,
                                      int nrows, void *stream) {
  DevGrid g = make_grid(grid);
  if (nrows <= 0) return 0;
  copy_guards_x_rows_kernel<<<gblk(nrows * 2 * g.lbx), GT, 0, (cudaStream_t)stream>>>(
      f, nc, g, iy0, nrows);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_add_guards(double *f, int nc, const skb_grid_t *grid, int phase,
                              const double *from_below, const double *from_above,
                              void *stream) {
  DevGrid g = make_grid(grid);
  cudaStream_t st = (cudaStream_t)stream;
  if (phase == 0) {
    add_guards_x_kernel<<<gblk(g.myp * nc), GT, 0, st>>>(f, nc, g);
  } else {
    add_guards_y_kernel<<<gblk(g.mx * nc), GT, 0, st>>>(f, nc, g, from_below, from_above);
  }
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_pack_rows(const double *f, int nc, const skb_grid_t *grid, int iy0,
                             int nrows, double *out, void *stream) {
  DevGrid g = make_grid(grid);
  long long n = (long long)nrows * g.nx * nc;
  if (n <= 0) return 0;
  pack_rows_kernel<<<gblk(n), GT, 0, (cudaStream_t)stream>>>(f, nc, g, iy0, nrows, out);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_scale(double *f, long long n, double fac, void *stream) {
  if (n <= 0) return 0;
  scale_kernel<<<gblk(n), GT, 0, (cudaStream_t)stream>>>(f, n, fac);
  SKB_CHECK_LAUNCH();
  return 0;
}

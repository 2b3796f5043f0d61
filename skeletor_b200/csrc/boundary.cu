// boundary.cu — particle boundary conditions and the local halves of the particle
// manager.
//
// Replaces periodic_x / calculate_ihole / shear_periodic_y (reference
// skeletor/cython/particle_boundary.pyx:5-49, particle_boundary.pxd:3-22) and the
// pack / redistribute parts of ppic2's cppmove2 (picksc/ppic2/pplib2.c:607-981).
// The neighbour exchange that sits between pack and unpack in cppmove2
// (MPI_Isend/Irecv, pplib2.c:741-753) is done by the caller over NCCL.
#include "common.cuh"

#define BT 256

__global__ void __launch_bounds__(BT)
periodic_x_kernel(double *x, long long np, double nx) {
  long long i = (long long)blockIdx.x * BT + threadIdx.x;
  if (i >= np) return;
  double v = x[i];
  double w = wrap_x(v, nx);
  if (w != v) x[i] = w;
}

// particle_boundary.pyx:26-49
__global__ void __launch_bounds__(BT)
shear_y_kernel(skb_particles_t P, long long np, double ny, double vx_boost,
               double x_boost) {
  long long i = (long long)blockIdx.x * BT + threadIdx.x;
  if (i >= np) return;
  double y = P.y[i];
  if (y < 0.0) {
    P.x[i] = P.x[i] - x_boost;
    P.vx[i] = P.vx[i] - vx_boost;
  }
  if (y >= ny) {
    P.x[i] = P.x[i] + x_boost;
    P.vx[i] = P.vx[i] + vx_boost;
  }
}

// ---- deterministic calculate_ihole: count / scan / ordered write --------------
__device__ __forceinline__ bool outside(double y, double e0, double e1) {
  return (y < e0) || (y >= e1);   // particle_boundary.pxd:15
}

__global__ void __launch_bounds__(BT)
hole_count_kernel(const double *__restrict__ y, long long np, double e0, double e1,
                  int *block_counts) {
  long long i = (long long)blockIdx.x * BT + threadIdx.x;
  bool out = (i < np) && outside(y[i], e0, e1);
  int c = __syncthreads_count(out);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// single CTA: exclusive scan of block_counts[nb] in place; total -> ihole[0] with the
// reference's overflow encoding (-(total-1) when total > ntmax, pxd:17-20)
__global__ void __launch_bounds__(1024)
hole_scan_kernel(int *block_counts, int nb, int *ihole, int ntmax) {
  __shared__ int wtot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nb) ? block_counts[i] : 0, s = v;
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(SKB_FULL, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) wtot[wv] = s;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wv; w++) woff += wtot[w];
    int excl = carry_s + woff + s - v;
    if (i < nb) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int total = carry_s;
    ihole[0] = (total > ntmax) ? -(total - 1) : total;
  }
}

__global__ void __launch_bounds__(BT)
hole_write_kernel(const double *__restrict__ y, long long np, double e0, double e1,
                  const int *__restrict__ block_offsets, int *ihole, int ntmax) {
  __shared__ int wtot[BT / 32];
  long long i = (long long)blockIdx.x * BT + threadIdx.x;
  bool out = (i < np) && outside(y[i], e0, e1);
  unsigned m = __ballot_sync(SKB_FULL, out);
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  if (lane == 0) wtot[wv] = __popc(m);
  __syncthreads();
  if (out) {
    int off = block_offsets[blockIdx.x];
    for (int w = 0; w < wv; w++) off += wtot[w];
    off += __popc(m & ((1u << lane) - 1u));
    if (off < ntmax) ihole[off + 1] = (int)i + 1;
  }
}

// ---- cppmove2, local halves ----------------------------------------------------
__device__ __forceinline__ void put_row(double *buf, int slot, double x, double y,
                                        double vx, double vy, double vz) {
  double *r = buf + (size_t)slot * 5;
  r[0] = x; r[1] = y; r[2] = vx; r[3] = vy; r[4] = vz;
}

// pplib2.c:666-707.  nh < 0: the hole count is read from ihole[0] on the device (no
// host round trip between the push and the pack); it is echoed into counts[3].
__global__ void __launch_bounds__(BT)
move_pack_kernel(skb_particles_t P, const int *__restrict__ ihole, int nh, double *sbufl,
                 double *sbufr, int nbmax, int *counts, double e0, double ny, int rank,
                 int nvp, int ntmax) {
  if (nh < 0) {
    nh = ihole[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[3] = nh;
    if (nh < 0) return;               // overflow / CFL flag: the host raises
    nh = min(nh, ntmax);
  }
  for (int j = blockIdx.x * BT + threadIdx.x; j < nh; j += gridDim.x * BT) {
    long long i = (long long)ihole[j + 1] - 1;
    double y = P.y[i];
    if (y < e0) {                       // going down
      if (rank == 0) y += ny;
      int slot = atomicAdd(counts + 0, 1);
      if (slot < nbmax) put_row(sbufl, slot, P.x[i], y, P.vx[i], P.vy[i], P.vz[i]);
      else counts[2] = 1;
    } else {                            // going up
      if (rank == nvp - 1) y -= ny;
      int slot = atomicAdd(counts + 1, 1);
      if (slot < nbmax) put_row(sbufr, slot, P.x[i], y, P.vx[i], P.vy[i], P.vz[i]);
      else counts[2] = 1;
    }
  }
}

// pplib2.c:756-866: particles that arrived but do not belong here are passed on.
// nrecv < 0: header mode (peer-memory exchange) — the count is the first double of
// rbuf, the rows follow the 5-double header; -nrecv - 1 is the row capacity.
__global__ void __launch_bounds__(BT)
move_classify_kernel(const double *__restrict__ rbuf, int nrecv, double *keep,
                     double *sbufl, double *sbufr, int nbmax, int *counts, double e0,
                     double e1, double ny, int rank, int nvp) {
  if (nrecv < 0) {
    nrecv = min((int)rbuf[0], -nrecv - 1);
    rbuf += 5;
  }
  for (int j = blockIdx.x * BT + threadIdx.x; j < nrecv; j += gridDim.x * BT) {
    const double *r = rbuf + (size_t)j * 5;
    double y = r[1];
    if (y < e0) {
      if (rank == 0) y += ny;
      int slot = atomicAdd(counts + 1, 1);
      if (slot < nbmax) put_row(sbufl, slot, r[0], y, r[2], r[3], r[4]);
      else counts[3] = 1;
    } else if (y >= e1) {
      if (rank == nvp - 1) y -= ny;
      int slot = atomicAdd(counts + 2, 1);
      if (slot < nbmax) put_row(sbufr, slot, r[0], y, r[2], r[3], r[4]);
      else counts[3] = 1;
    } else {
      // keep holds 2 * nbmax rows (both neighbours' buffers can arrive full); the count
      // runs on across forwarding rounds, so it is checked like the send buffers
      int slot = atomicAdd(counts + 0, 1);
      if (slot < 2 * nbmax) put_row(keep, slot, r[0], y, r[2], r[3], r[4]);
      else counts[3] = 1;
    }
  }
}

// incoming particle j -> hole j, or appended at np + (j - nh)   (pplib2.c:883-926)
__global__ void __launch_bounds__(BT)
move_fill_kernel(skb_particles_t P, long long np, const int *__restrict__ ihole, int nh,
                 const double *__restrict__ in, int nin) {
  int j = blockIdx.x * BT + threadIdx.x;
  if (j >= nin) return;
  long long d = (j < nh) ? (long long)ihole[j + 1] - 1 : np + (j - nh);
  const double *r = in + (size_t)j * 5;
  P.x[d] = r[0]; P.y[d] = r[1]; P.vx[d] = r[2]; P.vy[d] = r[3]; P.vz[d] = r[4];
}

// holes left over (nh > nin): the array shrinks to np2 = np - (nh - nin); live
// particles in the tail [np2, np) move into the leftover holes below np2
// (pplib2.c:927-952).  scratch: mark[r] | low[r] | nlow
__global__ void __launch_bounds__(BT)
move_mark_kernel(const int *__restrict__ ihole, int nin, int nh, long long np2,
                 int *mark, int *low, int *nlow) {
  int j = blockIdx.x * BT + threadIdx.x;
  int r = nh - nin;
  if (j >= r) return;
  long long h = (long long)ihole[nin + j + 1] - 1;
  if (h >= np2) mark[h - np2] = 1;
  else low[atomicAdd(nlow, 1)] = (int)h;
}

__global__ void __launch_bounds__(1024)
move_compact_kernel(skb_particles_t P, long long np2, int r, const int *__restrict__ mark,
                    const int *__restrict__ low) {
  // single CTA; k-th live tail slot (ascending) -> low[k]
  __shared__ int wtot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  for (int base = 0; base < r; base += 1024) {
    int s = base + threadIdx.x;
    bool live = (s < r) && (mark[s] == 0);
    unsigned m = __ballot_sync(SKB_FULL, live);
    if (lane == 0) wtot[wv] = __popc(m);
    __syncthreads();
    int off = carry_s;
    for (int w = 0; w < wv; w++) off += wtot[w];
    if (live) {
      int k = off + __popc(m & ((1u << lane) - 1u));
      long long src = np2 + s, dst = low[k];
      P.x[dst] = P.x[src]; P.y[dst] = P.y[src]; P.vx[dst] = P.vx[src];
      P.vy[dst] = P.vy[src]; P.vz[dst] = P.vz[src];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 32; w++) t += wtot[w];
      carry_s += t;
    }
    __syncthreads();
  }
}

// ---- C ABI ---------------------------------------------------------------------
static inline unsigned nblk(long long n) { return (unsigned)((n + BT - 1) / BT); }

extern "C" int skb_periodic_x(skb_particles_t p, long long np, const skb_grid_t *grid,
                              void *stream) {
  if (np <= 0) return 0;
  periodic_x_kernel<<<nblk(np), BT, 0, (cudaStream_t)stream>>>(p.x, np, (double)grid->nx);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_shear_periodic_y(skb_particles_t p, long long np,
                                    const skb_grid_t *grid, double S, double t,
                                    void *stream) {
  if (np <= 0) return 0;
  double vx_boost = S * grid->Ly;               // particle_boundary.pyx:37-38
  double x_boost = vx_boost * t / grid->dx;
  shear_y_kernel<<<nblk(np), BT, 0, (cudaStream_t)stream>>>(p, np, (double)grid->ny,
                                                            vx_boost, x_boost);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" long long skb_ihole_scratch_ints(long long np) { return (np + BT - 1) / BT + 1; }

extern "C" int skb_calculate_ihole(skb_particles_t p, long long np, int *ihole, int ntmax,
                                   const skb_grid_t *grid, int *scratch, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const double e0 = grid->edges[0], e1 = grid->edges[1];
  unsigned nb = nblk(np);
  if (np > 0) {
    hole_count_kernel<<<nb, BT, 0, st>>>(p.y, np, e0, e1, scratch);
    SKB_CHECK_LAUNCH();
  }
  hole_scan_kernel<<<1, 1024, 0, st>>>(scratch, (int)nb * (np > 0), ihole, ntmax);
  SKB_CHECK_LAUNCH();
  if (np > 0) {
    hole_write_kernel<<<nb, BT, 0, st>>>(p.y, np, e0, e1, scratch, ihole, ntmax);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int skb_move_pack(skb_particles_t p, const int *ihole, int nh, double *sbufl,
                             double *sbufr, int nbmax, int *counts,
                             const skb_grid_t *grid, int rank, int nvp, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, 4 * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  if (nh == 0) return 0;
  // nh < 0: count on the device; -nh - 1 is the capacity of the hole list (ntmax)
  const int ntmax = (nh < 0) ? -nh - 1 : nh;
  const unsigned blocks = (nh < 0) ? min(nblk(ntmax), 1184u) : nblk(nh);
  if (blocks == 0) return 0;
  move_pack_kernel<<<blocks, BT, 0, st>>>(p, ihole, nh < 0 ? -1 : nh, sbufl, sbufr, nbmax,
                                          counts, grid->edges[0], (double)grid->ny, rank,
                                          nvp, ntmax);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_move_classify(const double *rbuf, int nrecv, double *keep,
                                 double *sbufl, double *sbufr, int nbmax, int *counts,
                                 const skb_grid_t *grid, int rank, int nvp, void *stream) {
  if (nrecv == 0) return 0;
  const unsigned blocks = (nrecv < 0) ? min(nblk(-nrecv - 1), 592u) : nblk(nrecv);
  if (blocks == 0) return 0;
  move_classify_kernel<<<blocks, BT, 0, (cudaStream_t)stream>>>(
      rbuf, nrecv, keep, sbufl, sbufr, nbmax, counts, grid->edges[0], grid->edges[1],
      (double)grid->ny, rank, nvp);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_move_unpack(skb_particles_t p, long long np, const int *ihole, int nh,
                               const double *in, int nin, int *scratch, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nin > 0) {
    move_fill_kernel<<<nblk(nin), BT, 0, st>>>(p, np, ihole, nh, in, nin);
    SKB_CHECK_LAUNCH();
  }
  if (nh > nin) {
    const int r = nh - nin;
    const long long np2 = np - r;
    int *mark = scratch, *low = scratch + r, *nlow = scratch + 2 * r;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(int) * (size_t)(2 * r + 1), st);
    if (e != cudaSuccess) return (int)e;
    move_mark_kernel<<<nblk(r), BT, 0, st>>>(ihole, nin, nh, np2, mark, low, nlow);
    SKB_CHECK_LAUNCH();
    move_compact_kernel<<<1, 1024, 0, st>>>(p, np2, r, mark, low);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}

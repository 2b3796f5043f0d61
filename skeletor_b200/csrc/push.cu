// push.cu — field gather + Boris / shearing-sheet push (+ fused particle-boundary
// epilogue) and drift.
//
// Replaces boris_push_cic/tsc, modified_boris_push_cic/tsc, drift
// (reference skeletor/cython/particle_push.pyx:4-169, particle_push.pxd:3-97) and,
// through the optional epilogue, shear_periodic_y / periodic_x / calculate_ihole
// (particle_boundary.pyx:5-49).
//
// Design: one CTA per work item of `span` (16384) consecutive particles.  Particles
// are tile-ordered (skb_tile_sort), so a work item covers one or a few tiles; for
// each the CTA stages the (tile + halo) window of E and B (interleaved x,y,z, as in
// HBM) in shared memory, one warp per row, and gathers from there.  A particle whose
// stencil is not inside the window (stale ordering, unsorted tail) gathers from
// global memory instead — ordering is a performance property, never a correctness
// requirement.  The particle coordinates of the NEXT iteration arrive through a
// two-deep cp.async ring in shared memory (no registers held while the load is in
// flight) and the lines of the iteration after that are prefetched into L2.
// Particle traffic is 5 coalesced 8-byte loads + 5 stores = 80 B/particle, the
// algorithmic minimum.  The optional epilogue folds shear_periodic_y, periodic_x,
// calculate_ihole and the first pass of the tile sort (histogram of the NEW cell
// keys) into the same pass.
#include "common.cuh"
#include "gather.cuh"

#define PUSH_THREADS 256
#ifndef PUSH_SPAN_CHUNKS
#define PUSH_SPAN_CHUNKS 8   // 16384 particles per CTA work item with chunk = 2048
#endif
#ifndef PUSH_PREFETCH
#define PUSH_PREFETCH 1
#endif

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// 8-byte asynchronous global -> shared copy (LDGSTS): no register is held while the
// load is in flight
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// slot layout of the ring: [2 buffers][5 coordinates][PUSH_THREADS]
__device__ __forceinline__ void async_load_particle(double *pbuf, int buf, const skb_particles_t &P,
                                                    long long i) {
  double *d = pbuf + buf * 5 * PUSH_THREADS + threadIdx.x;
  cp_async8(d, P.x + i);
  cp_async8(d + PUSH_THREADS, P.y + i);
  cp_async8(d + 2 * PUSH_THREADS, P.vx + i);
  cp_async8(d + 3 * PUSH_THREADS, P.vy + i);
  cp_async8(d + 4 * PUSH_THREADS, P.vz + i);
}

struct PushParams {
  KickParams k;
  double dtdsx, dtdsy;
  // epilogue
  int flags, ntmax;
  double vx_boost, x_boost;
  int *ihole;
  int *cell_counts;   // SKB_EPI_COUNT
  KeyParams key;
};

// shear_periodic_y + calculate_ihole + periodic_x for one particle
// (particle_boundary.pyx:26-49, pxd:10-22, pxd:3-7; order as Particles.push,
// particles.py:181-188: shear boost, hole detection, then x wrap)
// FLAGS >= 0: the epilogue flags are compile-time constants (the production
// combinations get their own kernel instances); FLAGS < 0: read q.flags at run time.
template <int FLAGS = -1>
__device__ __forceinline__ void boundary_epilogue(const PushParams &q, const DevGrid &g,
                                                  long long i, double &x, double y,
                                                  double &vx) {
  const int flags = (FLAGS >= 0) ? FLAGS : q.flags;
  if (flags & SKB_EPI_SHEAR) {
    if (y < 0.0) { x = x - q.x_boost; vx = vx - q.vx_boost; }
    if (y >= (double)g.ny) { x = x + q.x_boost; vx = vx + q.vx_boost; }
  }
  if (flags & SKB_EPI_HOLES) {
    if (y < g.e0 || y >= g.e1) {
      int slot = atomicAdd(q.ihole, 1);
      if (slot < q.ntmax) q.ihole[slot + 1] = (int)i + 1;
    }
  }
  if (flags & SKB_EPI_PERIODIC_X) x = wrap_x(x, (double)g.nx);
}

// gather + kick + drift + boundary epilogue for one particle
template <int ORDER, bool MODIFIED, int FLAGS = -1>
__device__ __forceinline__ void push_one(const double *sE, const double *sB, const Window &w,
                                         int wstride, const double *E, const double *B,
                                         const DevGrid &g, const PushParams &q, long long i,
                                         double &x, double &y, double &vx, double &vy,
                                         double &vz) {
  fields_and_kick<ORDER, MODIFIED>(sE, sB, w, wstride, E, B, g, q.k, x, y, vx, vy, vz);
  // drift_particle, particle_push.pxd:88-91
  x = x + vx * q.dtdsx;
  y = y + vy * q.dtdsy;
  if (FLAGS >= 0) { if (FLAGS) boundary_epilogue<FLAGS>(q, g, i, x, y, vx); }
  else if (q.flags) boundary_epilogue<-1>(q, g, i, x, y, vx);
}

// resident CTAs per SM the register allocation aims for: 3 for CIC (80 registers, no
// spills), 2 for TSC (118); measured best on B200
template <int ORDER, bool MODIFIED, int FLAGS>
__global__ void __launch_bounds__(PUSH_THREADS, (ORDER == 1) ? 3 : 2)
push_kernel(skb_particles_t P, long long np, const double *__restrict__ E,
            const double *__restrict__ B, DevGrid g, DevTiling tl, PushParams q,
            int span, int wstride, int wrows) {
  extern __shared__ double smem[];
  double *sE = smem;
  double *sB = smem + (size_t)wstride * wrows * 3;
  double *pbuf = sB + (size_t)wstride * wrows * 3;   // [2][5][PUSH_THREADS] particle ring
  const int flags = (FLAGS >= 0) ? FLAGS : q.flags;

  SegmentIter it;
  it.init(tl, np, span);
  long long s0, s1;
  int tile;
  while (it.next(tl, s0, s1, tile)) {
    Window w = tile_window(tile, tl, g);
    if (tile >= 0) {
      __syncthreads();  // previous segment's readers are done
      stage_window(sE, E, w, wstride, g);
      stage_window(sB, B, w, wstride, g);
      __syncthreads();
    }
    // Software pipeline through shared memory: the five coordinates of the particle
    // this thread handles in the NEXT iteration are already on their way
    // (cp.async, no registers held) while the current one is pushed; the lines of
    // the iteration after that are being pulled into L2.
    // Warp-uniform trip count (the fused histogram below is a warp collective).
    const int lane = threadIdx.x & 31;
    const long long first = s0 + (threadIdx.x & ~31);
    if (first + lane < s1) async_load_particle(pbuf, 0, P, first + lane);
    cp_async_commit();
    int buf = 0;
    for (long long base = first; base < s1; base += PUSH_THREADS, buf ^= 1) {
      const long long i = base + lane;
      const bool act = i < s1;
      const long long in = i + PUSH_THREADS;
      if (in < s1) async_load_particle(pbuf, buf ^ 1, P, in);
      cp_async_commit();
      const long long ip = i + (PUSH_PREFETCH + 1) * PUSH_THREADS;
      if (ip < s1) {
        prefetch_l2(P.x + ip); prefetch_l2(P.y + ip); prefetch_l2(P.vx + ip);
        prefetch_l2(P.vy + ip); prefetch_l2(P.vz + ip);
      }
      cp_async_wait_one();   // everything but the group just committed has landed
      int key = -1;
      if (act) {
        const double *pb = pbuf + buf * 5 * PUSH_THREADS + threadIdx.x;
        double x = pb[0], y = pb[PUSH_THREADS], vx = pb[2 * PUSH_THREADS],
               vy = pb[3 * PUSH_THREADS], vz = pb[4 * PUSH_THREADS];
        push_one<ORDER, MODIFIED, FLAGS>(sE, sB, w, wstride, E, B, g, q, i, x, y, vx, vy, vz);
        P.x[i] = x; P.y[i] = y; P.vx[i] = vx; P.vy[i] = vy; P.vz[i] = vz;
        if ((flags & SKB_EPI_COUNT) && !(y < g.e0 || y >= g.e1)) key = cell_key(x, y, q.key);
      }
      if (flags & SKB_EPI_COUNT) {
        // first pass of the tile sort: one integer atomic per distinct new cell
        const unsigned peers = __match_any_sync(SKB_FULL, key);
        if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(q.cell_counts + key, __popc(peers));
      }
    }
    cp_async_wait_all();
  }
}

// ---------------------------------------------------------------------------------
// Fused push + tile sort in two passes that both RECOMPUTE the push from the old
// state instead of writing it back in between:
//   PASS 1 (push_count):   read 40 B, push in registers; leavers are packed straight
//                          into the exchange buffers (cppmove2's pack, pplib2.c:666-707),
//                          everybody else adds to the histogram of NEW cell keys.
//   PASS 2 (push_scatter): read 40 B, push again (bit-identical arithmetic), claim a
//                          slot in the new cell and write 40 B to the sorted position.
// 120 B/particle instead of 80 (push) + 16 (key pass) + 80 (move) = 176 B, no hole
// list and no hole filling.  Arrivals are counted / scattered by the small row kernels
// of sort.cu between and after the passes.
struct SortParams {
  KeyParams key;
  int *cells;          // PASS 1: histogram; PASS 2: next free slot per cell
  double *sbufl, *sbufr;
  int *counts;         // [0] = #sbufl, [1] = #sbufr, [2] = overflow
  int nbmax, rank, nvp;
  skb_particles_t out;
};

template <int ORDER, bool MODIFIED, int PASS>
__global__ void __launch_bounds__(PUSH_THREADS)
push_sort_kernel(skb_particles_t P, long long np, const double *__restrict__ E,
                 const double *__restrict__ B, DevGrid g, DevTiling tl, PushParams q,
                 SortParams sp, int span, int wstride, int wrows) {
  extern __shared__ double smem[];
  double *sE = smem;
  double *sB = smem + (size_t)wstride * wrows * 3;
  const int lane = threadIdx.x & 31;

  SegmentIter it;
  it.init(tl, np, span);
  long long s0, s1;
  int tile;
  while (it.next(tl, s0, s1, tile)) {
    Window w = tile_window(tile, tl, g);
    if (tile >= 0) {
      __syncthreads();
      stage_window(sE, E, w, wstride, g);
      stage_window(sB, B, w, wstride, g);
      __syncthreads();
    }
    // warp-uniform trip count: the slot claim below is a warp collective
    for (long long base = s0 + (threadIdx.x & ~31); base < s1; base += PUSH_THREADS) {
      const long long i = base + lane;
      const bool act = i < s1;
      const long long ip = i + PUSH_PREFETCH * PUSH_THREADS;
      if (ip < s1) {
        prefetch_l2(P.x + ip); prefetch_l2(P.y + ip); prefetch_l2(P.vx + ip);
        prefetch_l2(P.vy + ip); prefetch_l2(P.vz + ip);
      }
      double x = 0, y = 0, vx = 0, vy = 0, vz = 0;
      int key = -1;
      if (act) {
        x = P.x[i]; y = P.y[i]; vx = P.vx[i]; vy = P.vy[i]; vz = P.vz[i];
        push_one<ORDER, MODIFIED>(sE, sB, w, wstride, E, B, g, q, i, x, y, vx, vy, vz);
        if (y < g.e0 || y >= g.e1) {            // leaves the slab
          if (PASS == 1) {
            double *buf; int slot;
            if (y < g.e0) {                      // going down, pplib2.c:674-688
              if (sp.rank == 0) y += (double)g.ny;
              slot = atomicAdd(sp.counts + 0, 1); buf = sp.sbufl;
            } else {                             // going up, pplib2.c:690-705
              if (sp.rank == sp.nvp - 1) y -= (double)g.ny;
              slot = atomicAdd(sp.counts + 1, 1); buf = sp.sbufr;
            }
            if (slot < sp.nbmax) {
              double *r = buf + (size_t)slot * 5;
              r[0] = x; r[1] = y; r[2] = vx; r[3] = vy; r[4] = vz;
            } else {
              sp.counts[2] = 1;
            }
          }
        } else {
          key = cell_key(x, y, sp.key);
        }
      }
      const unsigned peers = __match_any_sync(SKB_FULL, key);
      if (key >= 0) {
        const int leader = __ffs(peers) - 1;
        if (PASS == 1) {
          if (lane == leader) atomicAdd(sp.cells + key, __popc(peers));
        } else {
          int slot0 = 0;
          if (lane == leader) slot0 = atomicAdd(sp.cells + key, __popc(peers));
          slot0 = __shfl_sync(peers, slot0, leader);
          const long long d = (long long)slot0 + __popc(peers & ((1u << lane) - 1u));
          sp.out.x[d] = x; sp.out.y[d] = y; sp.out.vx[d] = vx; sp.out.vy[d] = vy;
          sp.out.vz[d] = vz;
        }
      }
    }
  }
}

typedef void (*PushKernel)(skb_particles_t, long long, const double *, const double *,
                           DevGrid, DevTiling, PushParams, int, int, int);

template <int FLAGS>
static PushKernel push_kernel_for(int order, bool modified) {
  if (order == 1) return modified ? push_kernel<1, true, FLAGS> : push_kernel<1, false, FLAGS>;
  return modified ? push_kernel<2, true, FLAGS> : push_kernel<2, false, FLAGS>;
}

// kernel instances with compile-time epilogue flags for the combinations Particles.push
// uses (holes + x wrap [+ shear] [+ sort histogram]) and for "no epilogue"; anything
// else runs the generic instance that reads the flags at run time
static PushKernel select_push_kernel(int order, bool modified, int flags) {
  const int HP = SKB_EPI_HOLES | SKB_EPI_PERIODIC_X;
  switch (flags) {
    case 0: return push_kernel_for<0>(order, modified);
    case HP: return push_kernel_for<HP>(order, modified);
    case HP | SKB_EPI_SHEAR: return push_kernel_for<HP | SKB_EPI_SHEAR>(order, modified);
    case HP | SKB_EPI_COUNT: return push_kernel_for<HP | SKB_EPI_COUNT>(order, modified);
    case HP | SKB_EPI_SHEAR | SKB_EPI_COUNT:
      return push_kernel_for<HP | SKB_EPI_SHEAR | SKB_EPI_COUNT>(order, modified);
    default: return push_kernel_for<-1>(order, modified);
  }
}

// ihole[0] holds the raw count after the kernel; give it the reference's in-band
// overflow encoding (particle_boundary.pxd:17-20: -ih of the last overflowing
// particle, i.e. -(count-1))
__global__ void finalize_ihole_kernel(int *ihole, int ntmax) {
  int n = ihole[0];
  if (n > ntmax) ihole[0] = -(n - 1);
}

__global__ void __launch_bounds__(256)
drift_kernel(skb_particles_t P, long long np, DevGrid g, PushParams q) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  double x = P.x[i], y = P.y[i];
  x = x + P.vx[i] * q.dtdsx;
  y = y + P.vy[i] * q.dtdsy;
  if (q.flags) {
    double vx = P.vx[i], vx0 = vx;
    boundary_epilogue(q, g, i, x, y, vx);
    if (vx != vx0) P.vx[i] = vx;
  }
  P.x[i] = x;
  P.y[i] = y;
}

static void fill_epilogue(PushParams &q, const skb_epilogue_t *epi, const DevGrid &g) {
  q.flags = 0; q.ntmax = 0; q.ihole = nullptr; q.vx_boost = 0; q.x_boost = 0;
  q.cell_counts = nullptr;
  q.key = make_keyparams(g, 1, 4, 4);
  if (!epi) return;
  q.flags = epi->flags;
  q.ntmax = epi->ntmax;
  q.ihole = epi->ihole;
  // particle_boundary.pyx:37-38
  q.vx_boost = epi->S * g.Ly;
  q.x_boost = q.vx_boost * epi->t / g.dx;
  if (q.flags & SKB_EPI_COUNT) {
    q.cell_counts = epi->cell_counts;
    q.key = make_keyparams(g, epi->key_order, epi->key_tlx, epi->key_tly);
  }
}

extern "C" int skb_boris_push(skb_particles_t p, long long np, const double *E,
                              const double *B, const skb_grid_t *grid, int order,
                              double qtmh, double dt, int modified, double Omega,
                              double S, const skb_tiling_t *tiling,
                              const skb_epilogue_t *epi, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DevGrid g = make_grid(grid);
  DevTiling tl = make_tiling(tiling);
  PushParams q;
  q.k = make_kick(g, qtmh, dt, Omega, S);
  q.dtdsx = dt / g.dx; q.dtdsy = dt / g.dy;     // particle_push.pyx:24-26
  fill_epilogue(q, epi, g);
  if (q.flags & SKB_EPI_HOLES) {
    cudaError_t e = cudaMemsetAsync(q.ihole, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  if (np > 0) {
    int mult = PUSH_SPAN_CHUNKS;   // fewer chunks per CTA when that would idle SMs
    while (mult > 1 && (np / ((long long)tl.chunk * mult)) < 4 * 148) mult >>= 1;
    const int span = tl.chunk * mult;
    const int ws = window_stride(tl), wr = window_rows(tl);
    size_t smem = ((size_t)ws * wr * 3 * 2 + 2 * 5 * PUSH_THREADS) * sizeof(double);
    long long nblk = (np + span - 1) / span;
    if (order != 1 && order != 2) return (int)cudaErrorInvalidValue;
    PushKernel k = select_push_kernel(order, modified != 0, q.flags);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    k<<<(unsigned)nblk, PUSH_THREADS, smem, st>>>(p, np, E, B, g, tl, q, span, ws, wr);
    SKB_CHECK_LAUNCH();
  }
  if (q.flags & SKB_EPI_HOLES) {
    finalize_ihole_kernel<<<1, 1, 0, st>>>(q.ihole, q.ntmax);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int skb_drift(skb_particles_t p, long long np, double dt,
                         const skb_grid_t *grid, const skb_epilogue_t *epi,
                         void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DevGrid g = make_grid(grid);
  PushParams q = {};
  q.dtdsx = dt / g.dx; q.dtdsy = dt / g.dy;     // particle_push.pyx:165-166
  fill_epilogue(q, epi, g);
  if (q.flags & SKB_EPI_HOLES) {
    cudaError_t e = cudaMemsetAsync(q.ihole, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  if (np > 0) {
    long long nblk = (np + 255) / 256;
    drift_kernel<<<(unsigned)nblk, 256, 0, st>>>(p, np, g, q);
    SKB_CHECK_LAUNCH();
  }
  if (q.flags & SKB_EPI_HOLES) {
    finalize_ihole_kernel<<<1, 1, 0, st>>>(q.ihole, q.ntmax);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}


template <int PASS>
static int launch_push_sort(skb_particles_t p, long long np, const double *E,
                            const double *B, const DevGrid &g, int order, int modified,
                            const DevTiling &tl, const PushParams &q, const SortParams &sp,
                            cudaStream_t st) {
  if (np <= 0) return 0;
  int mult = PUSH_SPAN_CHUNKS;
  while (mult > 1 && (np / ((long long)tl.chunk * mult)) < 4 * 148) mult >>= 1;
  const int span = tl.chunk * mult;
  const int ws = window_stride(tl), wr = window_rows(tl);
  size_t smem = (size_t)ws * wr * 3 * 2 * sizeof(double);
  long long nblk = (np + span - 1) / span;
  void (*k)(skb_particles_t, long long, const double *, const double *, DevGrid, DevTiling,
            PushParams, SortParams, int, int, int);
  if (order == 1) k = modified ? push_sort_kernel<1, true, PASS> : push_sort_kernel<1, false, PASS>;
  else if (order == 2) k = modified ? push_sort_kernel<2, true, PASS> : push_sort_kernel<2, false, PASS>;
  else return (int)cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  k<<<(unsigned)nblk, PUSH_THREADS, smem, st>>>(p, np, E, B, g, tl, q, sp, span, ws, wr);
  SKB_CHECK_LAUNCH();
  return 0;
}

static PushParams fused_params(const DevGrid &g, double qtmh, double dt, double Omega,
                               double S, const skb_epilogue_t *epi) {
  PushParams q;
  q.k = make_kick(g, qtmh, dt, Omega, S);
  q.dtdsx = dt / g.dx; q.dtdsy = dt / g.dy;
  fill_epilogue(q, epi, g);
  q.flags &= ~(SKB_EPI_HOLES | SKB_EPI_COUNT);   // leavers are packed directly
  return q;
}

extern "C" int skb_push_count(skb_particles_t p, long long np, const double *E,
                              const double *B, const skb_grid_t *grid, int order,
                              double qtmh, double dt, int modified, double Omega, double S,
                              const skb_tiling_t *tiling, const skb_epilogue_t *epi,
                              int tlx, int tly, int *cell_counts, double *sbufl,
                              double *sbufr, int nbmax, int *counts, int rank, int nvp,
                              void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DevGrid g = make_grid(grid);
  DevTiling tl = make_tiling(tiling);
  PushParams q = fused_params(g, qtmh, dt, Omega, S, epi);
  SortParams sp = {};
  sp.key = make_keyparams(g, order, tlx, tly);
  sp.cells = cell_counts; sp.sbufl = sbufl; sp.sbufr = sbufr; sp.counts = counts;
  sp.nbmax = nbmax; sp.rank = rank; sp.nvp = nvp;
  cudaError_t e = cudaMemsetAsync(counts, 0, 4 * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  int rc = skb_sort_clear(cell_counts, grid, tlx, tly, stream);
  if (rc) return rc;
  return launch_push_sort<1>(p, np, E, B, g, order, modified, tl, q, sp, st);
}

extern "C" int skb_push_scatter(skb_particles_t p, skb_particles_t out, long long np,
                                const double *E, const double *B, const skb_grid_t *grid,
                                int order, double qtmh, double dt, int modified,
                                double Omega, double S, const skb_tiling_t *tiling,
                                const skb_epilogue_t *epi, int tlx, int tly, int *cell_pos,
                                void *stream) {
  DevGrid g = make_grid(grid);
  DevTiling tl = make_tiling(tiling);
  PushParams q = fused_params(g, qtmh, dt, Omega, S, epi);
  SortParams sp = {};
  sp.key = make_keyparams(g, order, tlx, tly);
  sp.cells = cell_pos; sp.out = out;
  return launch_push_sort<2>(p, np, E, B, g, order, modified, tl, q, sp, (cudaStream_t)stream);
}

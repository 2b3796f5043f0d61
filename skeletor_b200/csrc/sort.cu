// sort.cu — GPU counting sort of the SoA particle arrays by tile-major cell key.
//
// New component (the reference's cppdsortp2yl / particle_sort.py is unused and
// broken); it exists to give push and deposit their locality:
// after the sort every tile's particles are one contiguous range (tile_offsets)
// and, inside a tile, particles of one stencil-base cell are consecutive.
//
// key(x, y) = tile-major index of (ixs, iys) = the E-gather / deposit stencil base
// cell in the extended [myp][mx] array, computed with exactly the arithmetic the
// deposit uses ((int)(x + (lbx - 0.5)) for CIC, (int)(x + (lbx - 0.5) + 0.5) for
// TSC), clamped into the array.
//
// Passes: (1) histogram of keys (warp-aggregated integer atomics, 16 B/particle
// read); (2) three-kernel exclusive scan over the cells -> cell starts, tile
// offsets; (3) scatter: each warp claims slots per distinct key with one atomic
// and moves the five arrays out of place (40 B read + 40 B written per particle);
// (4) chunk -> first tile table for the CTA work items of push/deposit.
// Order of particles inside one cell follows claim order (not contractual, like
// the reference's particle order, tests/test_skeletor.py:144-147).
#include "common.cuh"

#define SORT_THREADS 256
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SORT_THREADS)
keys_kernel(skb_particles_t P, long long np, KeyParams kp, int *keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) keys[i] = cell_key(P.x[i], P.y[i], kp);
}

__global__ void __launch_bounds__(SORT_THREADS)
count_kernel(skb_particles_t P, long long np, KeyParams kp, int *counts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < np;
  int key = act ? cell_key(P.x[i], P.y[i], kp) : -1;
  // one atomic per distinct key in the warp
  unsigned peers = __match_any_sync(SKB_FULL, key);
  const int lane = threadIdx.x & 31;
  if (act && lane == __ffs(peers) - 1) atomicAdd(counts + key, __popc(peers));
}

// ---- exclusive scan over n ints, in place, 3 kernels -------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const int *__restrict__ a, int n, int *block_sums) {
  __shared__ int red[SCAN_THREADS / 32];
  long long base = (long long)blockIdx.x * SCAN_TILE;
  int s = 0;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += a[i];
  }
  for (int d = 16; d >= 1; d >>= 1) s += __shfl_down_sync(SKB_FULL, s, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; w++) t += red[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024)
scan_top_kernel(int *block_sums, int nb) {
  // single CTA, exclusive scan of nb <= a few thousand entries
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nb) ? block_sums[i] : 0;
    int s = v;
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(SKB_FULL, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) warp_tot[wv] = s;
    __syncthreads();
    if (wv == 0) {
      int t = warp_tot[lane];
      int u = t;
      for (int d = 1; d < 32; d <<= 1) {
        int r = __shfl_up_sync(SKB_FULL, u, d);
        if (lane >= d) u += r;
      }
      warp_tot[lane] = u - t;  // exclusive warp offsets
      if (lane == 31) warp_tot[31] = u - t;
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_tot[wv] + s - v;
    if (i < nb) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
}

// in-place: a[i] <- exclusive prefix; also records tile starts
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(int *a, int n, const int *__restrict__ block_sums, int cells_log2,
                  int *tile_offsets) {
  __shared__ int warp_tot[SCAN_THREADS / 32];
  __shared__ int carry_s;
  long long base = (long long)blockIdx.x * SCAN_TILE;
  if (threadIdx.x == 0) carry_s = block_sums[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = base + k * SCAN_THREADS + threadIdx.x;
    int v = (i < n) ? a[i] : 0;
    int s = v;
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(SKB_FULL, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) warp_tot[wv] = s;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wv; w++) woff += warp_tot[w];
    int excl = carry_s + woff + s - v;
    if (i < n) {
      a[i] = excl;
      if (tile_offsets && (i & ((1LL << cells_log2) - 1)) == 0) tile_offsets[i >> cells_log2] = excl;
    }
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) carry_s = excl + v;
    __syncthreads();
  }
}

// scatter: cell_pos[key] holds the next free slot of the cell (starts at the
// exclusive prefix); a warp claims a block of slots per distinct key.  Each thread
// moves SCATTER_ITEMS particles (warp-strided), with all their loads issued before
// the first dependent atomic: more bytes in flight per warp.
#ifndef SCATTER_ITEMS
#define SCATTER_ITEMS 2
#endif
__global__ void __launch_bounds__(SORT_THREADS)
scatter_kernel(skb_particles_t in, skb_particles_t out, long long np, KeyParams kp,
               int *cell_pos) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * SORT_THREADS + (threadIdx.x & ~31)) *
                          SCATTER_ITEMS;
  double x[SCATTER_ITEMS], y[SCATTER_ITEMS], vx[SCATTER_ITEMS], vy[SCATTER_ITEMS],
      vz[SCATTER_ITEMS];
#pragma unroll
  for (int u = 0; u < SCATTER_ITEMS; u++) {
    const long long i = warp0 + u * 32 + lane;
    if (i < np) { x[u] = in.x[i]; y[u] = in.y[i]; vx[u] = in.vx[i]; vy[u] = in.vy[i]; vz[u] = in.vz[i]; }
  }
#pragma unroll
  for (int u = 0; u < SCATTER_ITEMS; u++) {
    const long long i = warp0 + u * 32 + lane;
    const bool act = i < np;
    int key = act ? cell_key(x[u], y[u], kp) : -1;
    unsigned peers = __match_any_sync(SKB_FULL, key);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (act && lane == leader) base = atomicAdd(cell_pos + key, __popc(peers));
    base = __shfl_sync(SKB_FULL, base, leader);
    if (act) {
      long long d = (long long)base + __popc(peers & ((1u << lane) - 1u));
      out.x[d] = x[u]; out.y[d] = y[u]; out.vx[d] = vx[u]; out.vy[d] = vy[u]; out.vz[d] = vz[u];
    }
  }
}

// chunk c (particles [c*chunk, (c+1)*chunk)) -> tile that holds its first particle
__global__ void __launch_bounds__(256)
chunk_table_kernel(const int *__restrict__ tile_offsets, int ntiles, int chunk,
                   int *chunk_first_tile) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  long long a = tile_offsets[t], b = tile_offsets[t + 1];
  if (b <= a) return;
  // chunks whose first particle c*chunk lies in [a, b)
  for (long long c = (a + chunk - 1) / chunk; c * chunk < b; c++) chunk_first_tile[c] = t;
}

extern "C" int skb_tile_geometry(const skb_grid_t *grid, int tlx, int tly, int *ntx,
                                 int *nty) {
  int mx = grid->nx + 2 * grid->lbx, myp = grid->nyp + 2 * grid->lby;
  *ntx = (mx + (1 << tlx) - 1) >> tlx;
  *nty = (myp + (1 << tly) - 1) >> tly;
  return 0;
}

extern "C" int skb_cell_keys(skb_particles_t p, long long np, const skb_grid_t *grid,
                             int order, int tlx, int tly, int *keys, void *stream) {
  if (np <= 0) return 0;
  DevGrid g = make_grid(grid);
  KeyParams kp = make_keyparams(g, order, tlx, tly);
  keys_kernel<<<(unsigned)((np + SORT_THREADS - 1) / SORT_THREADS), SORT_THREADS, 0,
                (cudaStream_t)stream>>>(p, np, kp, keys);
  SKB_CHECK_LAUNCH();
  return 0;
}

// ---- canonical order inside every cell (optional, for bitwise reproducibility) ----
// The scatter leaves the particles of one cell in claim order, which depends on warp
// scheduling.  This pass rewrites every cell's range in lexicographic order of
// (x, y, vx, vy, vz): the stored ARRAY then depends only on the set of particles, so
// the deposit's summation order — and with it every later bit — is reproducible from
// run to run and testable against np.lexsort.  One warp per cell; rank by counting
// (O(n^2 / 32) per cell: n is the number of particles per cell), out of place.
__device__ __forceinline__ bool row_less(double ax, double ay, double avx, double avy,
                                         double avz, int ai, double bx, double by,
                                         double bvx, double bvy, double bvz, int bi) {
  if (ax != bx) return ax < bx;
  if (ay != by) return ay < by;
  if (avx != bvx) return avx < bvx;
  if (avy != bvy) return avy < bvy;
  if (avz != bvz) return avz < bvz;
  return ai < bi;     // identical rows: any order gives the same array
}

#define CANON_MAX 1024   // cells up to this many particles are sorted in shared memory

__device__ __forceinline__ unsigned long long sortable_bits(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);   // total order == numeric order
}

// rank-by-counting fallback for very crowded cells (n > CANON_MAX): O(n^2 / 32)
__device__ __forceinline__ void canonical_by_counting(const skb_particles_t &in,
                                                      const skb_particles_t &out, int s,
                                                      int e, int lane) {
  for (int i = s + lane; i < e; i += 32) {
    const double x = in.x[i], y = in.y[i], vx = in.vx[i], vy = in.vy[i], vz = in.vz[i];
    int rank = 0;
    for (int j = s; j < e; j++) {
      const double xj = in.x[j];
      if (xj < x) rank++;
      else if (xj == x && j != i &&
               row_less(xj, in.y[j], in.vx[j], in.vy[j], in.vz[j], j, x, y, vx, vy, vz, i))
        rank++;
    }
    const int d = s + rank;
    out.x[d] = x; out.y[d] = y; out.vx[d] = vx; out.vy[d] = vy; out.vz[d] = vz;
  }
}

// One warp per cell: bitonic sort of (sortable x bits, local index) pairs in shared
// memory; ties in x (quiet starts put many particles on the same x) are resolved by
// comparing the remaining coordinates from global memory.
__global__ void __launch_bounds__(SORT_THREADS)
canonical_cells_kernel(skb_particles_t in, skb_particles_t out,
                       const int *__restrict__ cell_end, int ncells) {
  extern __shared__ unsigned long long canon_smem[];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  unsigned long long *keys = canon_smem + (size_t)wv * CANON_MAX;
  int *idx = (int *)(canon_smem + (size_t)(SORT_THREADS / 32) * CANON_MAX) + wv * CANON_MAX;
  const int warps = (gridDim.x * SORT_THREADS) >> 5;
  for (int cell = (blockIdx.x * SORT_THREADS + threadIdx.x) >> 5; cell < ncells;
       cell += warps) {
    const int s = cell ? cell_end[cell - 1] : 0;
    const int e = cell_end[cell];
    const int n = e - s;
    if (n <= 0) continue;
    if (n == 1) {
      if (lane == 0) {
        out.x[s] = in.x[s]; out.y[s] = in.y[s]; out.vx[s] = in.vx[s];
        out.vy[s] = in.vy[s]; out.vz[s] = in.vz[s];
      }
      continue;
    }
    if (n > CANON_MAX) { canonical_by_counting(in, out, s, e, lane); continue; }
    int P = 32;
    while (P < n) P <<= 1;
    for (int t = lane; t < P; t += 32) {
      keys[t] = (t < n) ? sortable_bits(in.x[s + t]) : ~0ull;   // padding sorts last
      idx[t] = t;
    }
    __syncwarp();
    for (int k = 2; k <= P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < (P >> 1); t += 32) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int l = i | j;
          const bool asc = (i & k) == 0;
          const unsigned long long ki = keys[i], kl = keys[l];
          const int ii = idx[i], il = idx[l];
          bool gt;                       // element i sorts after element l ?
          if (ki != kl) gt = ki > kl;
          else if (ii >= n || il >= n) gt = ii > il;   // padding vs padding / tie
          else {
            const int a = s + ii, b = s + il;
            gt = row_less(in.x[b], in.y[b], in.vx[b], in.vy[b], in.vz[b], il,
                          in.x[a], in.y[a], in.vx[a], in.vy[a], in.vz[a], ii);
          }
          if (gt == asc) { keys[i] = kl; keys[l] = ki; idx[i] = il; idx[l] = ii; }
        }
        __syncwarp();
      }
    for (int t = lane; t < n; t += 32) {
      const int src = s + idx[t], d = s + t;
      out.x[d] = in.x[src]; out.y[d] = in.y[src]; out.vx[d] = in.vx[src];
      out.vy[d] = in.vy[src]; out.vz[d] = in.vz[src];
    }
    __syncwarp();
  }
}

// AoS rows (migration arrivals): histogram / scatter into the sorted SoA arrays
__global__ void __launch_bounds__(SORT_THREADS)
count_rows_kernel(const double *__restrict__ rows, int n, KeyParams kp, int *counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(counts + cell_key(rows[(size_t)i * 5], rows[(size_t)i * 5 + 1], kp), 1);
}

__global__ void __launch_bounds__(SORT_THREADS)
scatter_rows_kernel(const double *__restrict__ rows, int n, skb_particles_t out,
                    KeyParams kp, int *cell_pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *r = rows + (size_t)i * 5;
  long long d = atomicAdd(cell_pos + cell_key(r[0], r[1], kp), 1);
  out.x[d] = r[0]; out.y[d] = r[1]; out.vx[d] = r[2]; out.vy[d] = r[3]; out.vz[d] = r[4];
}

static int scan_cells(int *cell_counts, const skb_grid_t *grid, int tlx, int tly, int chunk,
                      int *block_sums, int *tile_offsets, int *chunk_first_tile,
                      cudaStream_t st) {
  int ntx, nty;
  skb_tile_geometry(grid, tlx, tly, &ntx, &nty);
  const int ntiles = ntx * nty;
  const long long ncells = (long long)ntiles << (tlx + tly);
  const int n = (int)ncells + 1;  // one extra entry: total
  const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb > 4096) return (int)cudaErrorInvalidValue;
  scan_reduce_kernel<<<nb, SCAN_THREADS, 0, st>>>(cell_counts, n, block_sums);
  SKB_CHECK_LAUNCH();
  scan_top_kernel<<<1, 1024, 0, st>>>(block_sums, nb);
  SKB_CHECK_LAUNCH();
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(cell_counts, n, block_sums, tlx + tly,
                                                tile_offsets);
  SKB_CHECK_LAUNCH();
  chunk_table_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(tile_offsets, ntiles, chunk,
                                                           chunk_first_tile);
  SKB_CHECK_LAUNCH();
  return 0;
}

static size_t cells_bytes(const skb_grid_t *grid, int tlx, int tly) {
  int ntx, nty;
  skb_tile_geometry(grid, tlx, tly, &ntx, &nty);
  return sizeof(int) * ((((size_t)ntx * nty) << (tlx + tly)) + 1);
}

extern "C" int skb_sort_clear(int *cell_counts, const skb_grid_t *grid, int tlx, int tly,
                              void *stream) {
  return (int)cudaMemsetAsync(cell_counts, 0, cells_bytes(grid, tlx, tly),
                              (cudaStream_t)stream);
}

extern "C" int skb_sort_scan(int *cell_counts, const skb_grid_t *grid, int tlx, int tly,
                             int chunk, int *block_sums, int *tile_offsets,
                             int *chunk_first_tile, void *stream) {
  return scan_cells(cell_counts, grid, tlx, tly, chunk, block_sums, tile_offsets,
                    chunk_first_tile, (cudaStream_t)stream);
}

extern "C" int skb_sort_count_rows(const double *rows, int n, const skb_grid_t *grid,
                                   int order, int tlx, int tly, int *cell_counts,
                                   void *stream) {
  if (n <= 0) return 0;
  DevGrid g = make_grid(grid);
  KeyParams kp = make_keyparams(g, order, tlx, tly);
  count_rows_kernel<<<(n + SORT_THREADS - 1) / SORT_THREADS, SORT_THREADS, 0,
                      (cudaStream_t)stream>>>(rows, n, kp, cell_counts);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_sort_scatter_rows(const double *rows, int n, skb_particles_t out,
                                     const skb_grid_t *grid, int order, int tlx, int tly,
                                     int *cell_pos, void *stream) {
  if (n <= 0) return 0;
  DevGrid g = make_grid(grid);
  KeyParams kp = make_keyparams(g, order, tlx, tly);
  scatter_rows_kernel<<<(n + SORT_THREADS - 1) / SORT_THREADS, SORT_THREADS, 0,
                        (cudaStream_t)stream>>>(rows, n, out, kp, cell_pos);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_tile_sort(skb_particles_t in, skb_particles_t out, long long np,
                             const skb_grid_t *grid, int order, int tlx, int tly,
                             int chunk, int *cell_counts, int *block_sums,
                             int *tile_offsets, int *chunk_first_tile, int stable,
                             int *perm, void *stream) {
  (void)perm;
  if (stable) return (int)cudaErrorNotSupported;  // reserved
  cudaStream_t st = (cudaStream_t)stream;
  DevGrid g = make_grid(grid);
  KeyParams kp = make_keyparams(g, order, tlx, tly);
  cudaError_t e = cudaMemsetAsync(cell_counts, 0, cells_bytes(grid, tlx, tly), st);
  if (e != cudaSuccess) return (int)e;
  const unsigned pblk = (unsigned)((np + SORT_THREADS - 1) / SORT_THREADS);
  if (np > 0) {
    count_kernel<<<pblk, SORT_THREADS, 0, st>>>(in, np, kp, cell_counts);
    SKB_CHECK_LAUNCH();
  }
  int rc = scan_cells(cell_counts, grid, tlx, tly, chunk, block_sums, tile_offsets,
                      chunk_first_tile, st);
  if (rc) return rc;
  if (np > 0) {
    const unsigned sblk = (unsigned)((np + SORT_THREADS * SCATTER_ITEMS - 1) /
                                     (SORT_THREADS * SCATTER_ITEMS));
    scatter_kernel<<<sblk, SORT_THREADS, 0, st>>>(in, out, np, kp, cell_counts);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int skb_tile_sort_precounted(skb_particles_t in, skb_particles_t out,
                                        long long np, const skb_grid_t *grid, int order,
                                        int tlx, int tly, int chunk, int *cell_counts,
                                        int *block_sums, int *tile_offsets,
                                        int *chunk_first_tile, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DevGrid g = make_grid(grid);
  KeyParams kp = make_keyparams(g, order, tlx, tly);
  int rc = scan_cells(cell_counts, grid, tlx, tly, chunk, block_sums, tile_offsets,
                      chunk_first_tile, st);
  if (rc) return rc;
  if (np > 0) {
    const unsigned sblk = (unsigned)((np + SORT_THREADS * SCATTER_ITEMS - 1) /
                                     (SORT_THREADS * SCATTER_ITEMS));
    scatter_kernel<<<sblk, SORT_THREADS, 0, st>>>(in, out, np, kp, cell_counts);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int skb_canonical_cells(skb_particles_t in, skb_particles_t out,
                                   const int *cell_end, const skb_grid_t *grid, int tlx,
                                   int tly, void *stream) {
  int ntx, nty;
  skb_tile_geometry(grid, tlx, tly, &ntx, &nty);
  const int ncells = (ntx * nty) << (tlx + tly);
  int blocks = (ncells + (SORT_THREADS / 32) - 1) / (SORT_THREADS / 32);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const size_t smem = (size_t)(SORT_THREADS / 32) * CANON_MAX * (sizeof(unsigned long long) + sizeof(int));
  cudaError_t e = cudaFuncSetAttribute(canonical_cells_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  canonical_cells_kernel<<<blocks, SORT_THREADS, smem, (cudaStream_t)stream>>>(in, out,
                                                                             cell_end, ncells);
  SKB_CHECK_LAUNCH();
  return 0;
}

// in-place exclusive scan of n ints (n <= 4096 * 4096); block_sums: >= 4100 ints
extern "C" int skb_exclusive_scan(int *a, int n, int *block_sums, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb > 4096) return (int)cudaErrorInvalidValue;
  scan_reduce_kernel<<<nb, SCAN_THREADS, 0, st>>>(a, n, block_sums);
  SKB_CHECK_LAUNCH();
  scan_top_kernel<<<1, 1024, 0, st>>>(block_sums, nb);
  SKB_CHECK_LAUNCH();
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(a, n, block_sums, 30, nullptr);
  SKB_CHECK_LAUNCH();
  return 0;
}

// chunk -> first tile table for a given dense ordering (tile_offsets)
extern "C" int skb_chunk_table(const int *tile_offsets, const skb_grid_t *grid, int tlx,
                               int tly, int chunk, int *chunk_first_tile, void *stream) {
  int ntx, nty;
  skb_tile_geometry(grid, tlx, tly, &ntx, &nty);
  const int ntiles = ntx * nty;
  chunk_table_kernel<<<(ntiles + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      tile_offsets, ntiles, chunk, chunk_first_tile);
  SKB_CHECK_LAUNCH();
  return 0;
}

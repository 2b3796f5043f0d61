/* skeletor_b200.h — C ABI of libskeletor_b200.so
 *
 * Drop-in replacement for the compiled layer of nbia-astro/skeletor's particle
 * hot path: the Cython module-level functions in skeletor/cython/*.pyx and
 * ppic2's particle manager cppmove2 (picksc/ppic2/pplib2.c:607-981).  Each entry
 * point cites the reference interface it replaces (paths relative to the
 * reference root).
 *
 * Conventions (same as the reference):
 *  - every buffer is allocated and owned by the caller (PyTorch tensors on the
 *    Python side); the library never allocates, frees or keeps pointers between
 *    calls.  All pointers are DEVICE pointers unless named h_*.
 *  - all calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *    the return value is a cudaError_t (0 = success) from launch-time checks.
 *  - errors that the reference reports in-band (ihole[0] < 0) stay in-band.
 *  - float64 only (reference types.pxd:6-8: real_t = double).  Arithmetic keeps
 *    the reference's operation order and is compiled with -fmad=false, so
 *    per-particle results are bit-identical to the reference's gcc -O2 x86-64
 *    code; only deposition sums differ (summation order).
 *
 * Data layout in HBM:
 *  - particles: structure of arrays, five contiguous double arrays x,y,vx,vy,vz
 *    (reference: AoS particle_t, types.pxd:10-11).  x,y in grid units, y global.
 *  - fields: C-order [myp][mx] of interleaved structs exactly as the reference's
 *    Float3 (E,B: x,y,z) / Float4 (sources: t=rho,x,y,z) NumPy dtypes
 *    (types.pyx:9-10, field.py:6-9); scalar fields [myp][mx] doubles.
 *  - migration buffers: AoS rows of 5 doubles (x,y,vx,vy,vz), as sbufl/sbufr.
 */
#ifndef SKELETOR_B200_H
#define SKELETOR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* grid_t of the reference (skeletor/cython/types.pxd:25-37, grid.py:7-67) */
typedef struct {
  int nx, ny;     /* global grid size */
  int nyp, noff;  /* rows owned by this slab, first global row */
  int lbx, lby;   /* guard layers = first active index */
  int ubx, uby;   /* first upper guard index */
  double dx, dy, Lx, Ly, x0, y0;
  double edges[2]; /* [noff, noff+nyp] as doubles */
} skb_grid_t;

/* SoA particle arrays (device pointers) */
typedef struct {
  double *x, *y, *vx, *vy, *vz;
} skb_particles_t;

/* Tile ordering produced by skb_tile_sort().  A "cell key" is the tile-major
 * index of the particle's E-gather/deposit stencil base cell in the extended
 * [myp][mx] array; tiles are 2^tlx x 2^tly cells.  Particles [0, n_sorted) are
 * ordered by key; [n_sorted, np) (e.g. arrivals not yet sorted) are unordered.
 * Kernels never REQUIRE the ordering for correctness: a particle whose stencil
 * is outside the shared-memory window of the tile it is filed under takes a
 * global-memory path.  tile_offsets == NULL means "no ordering known". */
typedef struct {
  const int *tile_offsets;     /* [ntx*nty + 1] first particle of each tile */
  const int *chunk_first_tile; /* [ceil(n_sorted/chunk)] tile of particle c*chunk */
  const int *cell_end;         /* [ncells] end of each cell's particle range (cell k =
                                  [cell_end[k-1], cell_end[k])), or NULL.  Lets the
                                  deposit give every warp whole cells. */
  int ntx, nty, tlx, tly;
  int chunk;                   /* particles per CTA work item */
  long long n_sorted;
  /* gapped layout (skb_gap_*): cell k owns slots [gap_start[k], gap_start[k+1]) of the
   * particle arrays and its gap_count[k] live particles sit at the front of that
   * range.  NULL = dense layout (cell_end). */
  const int *gap_start;
  const int *gap_count;
} skb_tiling_t;

/* flags for the fused particle-boundary epilogue of skb_boris_push/skb_drift */
#define SKB_EPI_NONE 0
#define SKB_EPI_SHEAR 1      /* shear_periodic_y, particle_boundary.pyx:26-49 */
#define SKB_EPI_PERIODIC_X 2 /* periodic_x,       particle_boundary.pyx:5-11  */
#define SKB_EPI_HOLES 4      /* calculate_ihole (unordered list + count)      */
#define SKB_EPI_COUNT 8      /* histogram of the NEW cell keys of the particles that
                                stay in the slab (first pass of the tile sort, fused) */
#define SKB_EPI_DRIFT_ONLY 16 /* (gapped sweeps) no gather / kick: drift, particle_push.pyx:159-169 */

typedef struct {
  int flags;
  double S, t;   /* shear rate and particle time AFTER the push (particles.py:154-155,175) */
  int *ihole;    /* [ntmax+1]: ihole[0] = count (negated on overflow), then 1-based
                    indices of particles with y < edges[0] or y >= edges[1] */
  int ntmax;
  int *cell_counts; /* SKB_EPI_COUNT: histogram [ncells+1], cleared by the caller */
  int key_order, key_tlx, key_tly; /* key definition, as skb_tile_sort */
} skb_epilogue_t;

int skb_version(void);
const char *skb_error_string(int err);

/* ---- push ----------------------------------------------------------------
 * boris_push_cic/tsc(particles, E, B, qtmh, dt, grid)       particle_push.pyx:4,81
 * modified_boris_push_cic/tsc(..., grid, Omega, S)          particle_push.pyx:40,117
 * order: 1 = CIC, 2 = TSC.  modified != 0 adds the rotation/shear terms.
 * tiling / epi may be NULL (plain reference semantics). */
int skb_boris_push(skb_particles_t p, long long np, const double *E,
                   const double *B, const skb_grid_t *grid, int order,
                   double qtmh, double dt, int modified, double Omega, double S,
                   const skb_tiling_t *tiling, const skb_epilogue_t *epi,
                   void *stream);

/* drift(particles, dt, grid)                                particle_push.pyx:159 */
int skb_drift(skb_particles_t p, long long np, double dt, const skb_grid_t *grid,
              const skb_epilogue_t *epi, void *stream);

/* ---- particle boundaries ---------------------------------------------------
 * periodic_x(particles, grid)                               particle_boundary.pyx:5  */
int skb_periodic_x(skb_particles_t p, long long np, const skb_grid_t *grid,
                   void *stream);
/* shear_periodic_y(particles, grid, S, t)                   particle_boundary.pyx:26 */
int skb_shear_periodic_y(skb_particles_t p, long long np, const skb_grid_t *grid,
                         double S, double t, void *stream);
/* calculate_ihole(particles, ihole, grid)                   particle_boundary.pyx:14
 * Deterministic: the list is in ascending particle order, bit-identical to the
 * reference's serial loop.  scratch: >= skb_ihole_scratch_ints(np) ints. */
long long skb_ihole_scratch_ints(long long np);
int skb_calculate_ihole(skb_particles_t p, long long np, int *ihole, int ntmax,
                        const skb_grid_t *grid, int *scratch, void *stream);

/* ---- particle manager: cppmove2(particles, npp, sbufl, sbufr, rbufl, rbufr,
 *      ihole, info, grid)           ppic2_wrapper.pyx:51-67, pplib2.c:607-981
 * split into its two local halves; the neighbour exchange between them
 * (MPI_Isend/Irecv, pplib2.c:741-753) is done by the caller (copies into the
 * neighbours' NVLink peer-memory slots, NCCL send/recv as fallback, or a device copy
 * when nvp == 1).
 *
 * skb_move_pack: for each listed hole, copy the particle into sbufl (y <
 *   edges[0]; y += ny if rank == 0, pplib2.c:676-677) or sbufr (otherwise; y -=
 *   ny if rank == nvp-1, :692-693).  counts[0] = #sbufl, counts[1] = #sbufr
 *   (device ints, zeroed by the call).  A buffer overflow (> nbmax) is reported
 *   as counts[2] = 1 and the surplus is NOT packed.  nh = -(ntmax + 1) means "read
 *   the hole count from ihole[0] on the device" (no host round trip between push and
 *   pack); it is echoed into counts[3].
 * skb_move_classify: multi-hop support (pplib2.c:756-866): split a received
 *   buffer into particles that belong here (copied to `keep`, 2 * nbmax rows; count in
 *   counts[0], which may carry on from earlier calls) and particles to pass further
 *   down / up (appended to sbufl / sbufr with the edge-rank y wrap; counts[1],
 *   counts[2]; counts[3] = overflow of any of the three buffers, rows beyond the
 *   capacity are not written).  The caller zeroes counts[1..3].  nrecv = -(capacity + 1): header mode for the
 *   peer-memory exchange — the row count is the first double of rbuf and the rows
 *   follow a 5-double header.
 * skb_move_unpack: put `nin` incoming particles (AoS rows in `in`) into the
 *   holes listed in ihole[1..nh], append what is left at np, or — if holes remain
 *   — compact the tail into them (pplib2.c:883-952).  New count = np + nin - nh
 *   (computed by the caller).  scratch: >= 2*nh + 8 ints. */
int skb_move_pack(skb_particles_t p, const int *ihole, int nh, double *sbufl,
                  double *sbufr, int nbmax, int *counts, const skb_grid_t *grid,
                  int rank, int nvp, void *stream);
int skb_move_classify(const double *rbuf, int nrecv, double *keep, double *sbufl,
                      double *sbufr, int nbmax, int *counts, const skb_grid_t *grid,
                      int rank, int nvp, void *stream);
int skb_move_unpack(skb_particles_t p, long long np, const int *ihole, int nh,
                    const double *in, int nin, int *scratch, void *stream);

/* ---- deposit ---------------------------------------------------------------
 * deposit_cic/tsc(particles, current, grid, S)               deposit.pyx:6,21
 * Accumulates into `current` (Float4 [myp][mx]); the caller zeroes it when the
 * reference's erase=True semantics are wanted (sources.py:37-38). */
int skb_deposit(skb_particles_t p, long long np, double *current,
                const skb_grid_t *grid, int order, double S,
                const skb_tiling_t *tiling, void *stream);

/* Deterministic deposit (no atomics): needs an exact ordering covering all particles
 * (tiling->cell_end, n_sorted == np).  Per-cell stencil sums go to cellsums
 * [ncells][(order+1)^2 * 4] doubles (scratch) and are gathered into `current` in a fixed
 * order, so the result is a pure function of the stored particle array. */
int skb_deposit_deterministic(skb_particles_t p, long long np, double *current,
                              const skb_grid_t *grid, int order, double S,
                              const skb_tiling_t *tiling, double *cellsums,
                              void *stream);

/* push_and_deposit_cic/tsc(particles, E, B, qtmh, dt, grid, ihole, current, S,
 *                          update)                   push_and_deposit.pyx:10,91
 * ihole semantics as skb_epilogue_t (unordered list); ihole[0] = -1 flags a
 * particle that moved more than half a cell in the half step (:66-68).
 * next_cell_counts (may be NULL; update only): histogram [ncells+1] of the NEW cell
 * keys (tiles 2^key_tlx x 2^key_tly) of the particles that stay, cleared by the
 * caller, for skb_tile_sort_precounted; must not alias tiling->cell_end and needs an
 * exact ordering (tiling->cell_end, n_sorted == np). */
int skb_push_and_deposit(skb_particles_t p, long long np, const double *E,
                         const double *B, const skb_grid_t *grid, int order,
                         double qtmh, double dt, int *ihole, int ntmax,
                         double *current, double S, int update,
                         const skb_tiling_t *tiling, int *next_cell_counts,
                         int key_tlx, int key_tly, void *stream);

/* ---- tile sort (new component; reference's cppdsortp2yl is unused/broken) ----
 * Counting sort of the SoA particle arrays by cell key, out of place.
 * cell_counts: [ncells+1] ints with ncells = ntx*nty << (tlx+tly) (scratch);
 * tile_offsets [ntx*nty+1] and
 * chunk_first_tile [ceil(np/chunk)] are outputs; block_sums: >= 4100 ints.
 * Ties (particles of one cell) end up in claim order, which is not contractual
 * (neither is particle order in the reference, tests/test_skeletor.py:144-147);
 * `stable` / `perm` are reserved for a stable variant and must be 0 / NULL. */
int skb_tile_geometry(const skb_grid_t *grid, int tlx, int tly, int *ntx, int *nty);
int skb_cell_keys(skb_particles_t p, long long np, const skb_grid_t *grid,
                  int order, int tlx, int tly, int *keys, void *stream);
int skb_tile_sort(skb_particles_t in, skb_particles_t out, long long np,
                  const skb_grid_t *grid, int order, int tlx, int tly, int chunk,
                  int *cell_counts, int *block_sums, int *tile_offsets,
                  int *chunk_first_tile, int stable, int *perm, void *stream);

/* ---- fused push + tile sort (alternative to push + skb_tile_sort_precounted; correct
 * and tested, but not faster on B200 because the push is instruction-bound) ------
 * Two passes that both recompute the push from the OLD particle state:
 * skb_push_count:   pass 1 — push in registers; particles that leave the slab are
 *   packed into sbufl / sbufr exactly as skb_move_pack would (counts[0], counts[1],
 *   overflow flag counts[2]); all others add to the histogram of their NEW cell key
 *   (cell_counts is cleared by the call).  The particle arrays are not written.
 * [caller: exchange buffers; skb_sort_count_rows(arrivals); skb_sort_scan]
 * skb_push_scatter: pass 2 — push again (bit-identical) and write every staying
 *   particle to its sorted position in `out`.
 * [caller: skb_sort_scatter_rows(arrivals)]
 * epi: SKB_EPI_SHEAR / SKB_EPI_PERIODIC_X flags and S, t (SKB_EPI_HOLES is ignored).
 * 120 B/particle of HBM traffic instead of 176 B for push + key pass + move. */
int skb_push_count(skb_particles_t p, long long np, const double *E, const double *B,
                   const skb_grid_t *grid, int order, double qtmh, double dt,
                   int modified, double Omega, double S, const skb_tiling_t *tiling,
                   const skb_epilogue_t *epi, int tlx, int tly, int *cell_counts,
                   double *sbufl, double *sbufr, int nbmax, int *counts, int rank,
                   int nvp, void *stream);
int skb_push_scatter(skb_particles_t p, skb_particles_t out, long long np,
                     const double *E, const double *B, const skb_grid_t *grid, int order,
                     double qtmh, double dt, int modified, double Omega, double S,
                     const skb_tiling_t *tiling, const skb_epilogue_t *epi, int tlx,
                     int tly, int *cell_pos, void *stream);
/* pieces of the tile sort, for the fused path: clear the histogram; histogram /
 * scatter of AoS rows (migration arrivals); scan + tile offsets + chunk table */
int skb_sort_clear(int *cell_counts, const skb_grid_t *grid, int tlx, int tly,
                   void *stream);
int skb_sort_count_rows(const double *rows, int n, const skb_grid_t *grid, int order,
                        int tlx, int tly, int *cell_counts, void *stream);
int skb_sort_scan(int *cell_counts, const skb_grid_t *grid, int tlx, int tly, int chunk,
                  int *block_sums, int *tile_offsets, int *chunk_first_tile,
                  void *stream);
int skb_sort_scatter_rows(const double *rows, int n, skb_particles_t out,
                          const skb_grid_t *grid, int order, int tlx, int tly,
                          int *cell_pos, void *stream);

/* skb_tile_sort without its histogram pass: cell_counts already holds the histogram
 * of the keys of in[0..np) (SKB_EPI_COUNT epilogue + skb_sort_count_rows). */
int skb_tile_sort_precounted(skb_particles_t in, skb_particles_t out, long long np,
                             const skb_grid_t *grid, int order, int tlx, int tly,
                             int chunk, int *cell_counts, int *block_sums,
                             int *tile_offsets, int *chunk_first_tile, void *stream);

/* Optional: rewrite every cell's particle range (cell_end as left by the sort) in
 * lexicographic order of (x, y, vx, vy, vz), out of place.  The array then depends only
 * on the SET of particles: bitwise reproducible, and equal to
 * np.lexsort((vz, vy, vx, y, x, key)) of the same particles. */
int skb_canonical_cells(skb_particles_t in, skb_particles_t out, const int *cell_end,
                        const skb_grid_t *grid, int tlx, int tly, void *stream);

/* ---- gapped particle layout (new component): ordering without a move pass ---------
 * Every cell owns a slot range with slack; gap_start [ncells+1] (slot ranges),
 * gap_count [ncells] (live particles at the front of each range).
 * skb_gap_build:   dense exactly-ordered arrays (cell_end) -> gapped arrays; nothing is
 *   written when the slot ranges need more than `capacity` slots (gap_start[ncells]).
 * skb_push_gapped: push (+ shear boost / x wrap: epi_flags, epi_S, epi_t) of every cell's
 *   particles; stayers are written back compacted into their own range, particles that
 *   change cell go to `movers` (AoS rows; counts[0]), particles that leave the slab to
 *   sbufl / sbufr as skb_move_pack would (counts[1], counts[2]); counts[4] = movers
 *   re-inserted in place (statistics); counts[3] = flags
 *   (1: mover list full, some particles were parked in their old cell -> rebuild
 *       (generic kernel only);  2: exchange buffer overflow;  8: mover lists full,
 *       particles were lost (cell-stream kernel: size the lists for the step)).
 *   The nleft particles of `leftover` (below) are pushed first, with the generic
 *   kernel, and join the head of the mover list (nleft <= mover_cap).
 *   Movers whose new cell is handled by the same thread block never reach `movers`:
 *   the block parks them in one of npool scratch blocks ([npool][scratch_rows][5]
 *   doubles, pool_owner [npool] zero-initialised; npool >= the number of thread blocks
 *   that can be resident (checked at launch), or 0 to disable) and
 *   inserts them itself; full cells overflow into `leftover` (leftover_counts[0] rows,
 *   [1] overflow flag; both reset by this call after the old leftovers are consumed).
 * skb_gap_insert:  drop AoS rows (movers, arrivals) into the free slots of their cells;
 *   rows that do not fit go to `leftover`, a small SoA list [5][leftover_cap]
 *   (counts[0] rows; counts[1] = leftover overflow) that lives beside the cells until
 *   the next push.
 * skb_gap_densify: gapped -> dense ordered arrays + cell_end + tile_offsets, leftover
 *   rows appended behind the n_in_cells ordered particles.
 * skb_exclusive_scan / skb_chunk_table: helpers shared with the tile sort. */
int skb_gap_build(skb_particles_t in, skb_particles_t out, const int *cell_end,
                  const skb_grid_t *grid, int tlx, int tly, int *gap_start,
                  int *gap_count, int *block_sums, long long capacity, void *stream);
int skb_push_gapped(skb_particles_t p, const double *E, const double *B,
                    const skb_grid_t *grid, int order, double qtmh, double dt,
                    int modified, double Omega, double S, int epi_flags, double epi_S,
                    double epi_t, int tlx, int tly, const int *gap_start, int *gap_count,
                    double *movers, int mover_cap, double *sbufl, double *sbufr,
                    int nbmax, int *counts, int rank, int nvp, double *leftover,
                    int leftover_cap, int nleft, int *leftover_counts, double *scratch,
                    int scratch_rows, int npool, int *pool_owner, void *stream);
/* push_and_deposit_cic/tsc (push_and_deposit.pyx:10,91) on the gapped layout: update != 0
 * does what skb_push_gapped does after the half-step deposit (second half drift, x wrap,
 * routing; counts / leftover as there), update == 0 only deposits (predictor sweep,
 * nothing is written).  counts[3] bit 4 (and ihole[0] == -1 for a leftover particle;
 * ihole: >= nleft + 1 ints of scratch) flags a particle that moved more than half a cell
 * in the half step (:66-68). */
int skb_push_and_deposit_gapped(skb_particles_t p, const double *E, const double *B,
                                const skb_grid_t *grid, int order, double qtmh, double dt,
                                double *current, double S, int update, int *ihole,
                                int ntmax, int tlx, int tly, const int *gap_start,
                                int *gap_count, double *movers, int mover_cap,
                                double *sbufl, double *sbufr, int nbmax, int *counts,
                                int rank, int nvp, double *leftover, int leftover_cap,
                                int nleft, int *leftover_counts, double *scratch,
                                int scratch_rows, int npool, int *pool_owner,
                                void *stream);
/* push / push_modified (particle_push.pyx:4,40,81,117) + boundary epilogue as
 * skb_push_gapped, fused with the full-step deposit that follows the push in the time
 * loop (Sources.deposit, sources.py:27-50 -> deposit_cic/tsc, deposit.pyx:6,21): `current`
 * (zeroed by the caller) receives the RAW stencil sums (dep_S: rate of shear of the
 * deposit, deposit.pxd:24) of every particle that is in the slab after the push; the
 * arrivals from the neighbour ranks are added with skb_deposit_rows.  counts[3] bit 8:
 * mover lists full, particles were lost (fatal); bit 16: some rows bypassed the fused
 * deposit (scratch block full) - discard `current` and call skb_deposit. */
int skb_push_deposit_gapped(skb_particles_t p, const double *E, const double *B,
                            const skb_grid_t *grid, int order, double qtmh, double dt,
                            int modified, double Omega, double S, int epi_flags,
                            double epi_S, double epi_t, double *current, double dep_S,
                            int tlx, int tly, const int *gap_start, int *gap_count,
                            double *movers, int mover_cap, double *sbufl, double *sbufr,
                            int nbmax, int *counts, int rank, int nvp, double *leftover,
                            int leftover_cap, int nleft, int *leftover_counts,
                            double *scratch, int scratch_rows, int npool, int *pool_owner,
                            void *stream);
/* deposit_cic/tsc (deposit.pyx:6,21) of n AoS rows {x, y, vx, vy, vz} into `current` */
int skb_deposit_rows(const double *rows, int n, double *current, const skb_grid_t *grid,
                     int order, double S, void *stream);
/* drift (particle_push.pyx:159-169) + periodic_x + the routing of skb_push_gapped on the
 * gapped layout: Particles.drift (particles.py:259-265) without leaving the layout. */
int skb_drift_gapped(skb_particles_t p, const skb_grid_t *grid, int order, double dt,
                     int tlx, int tly, const int *gap_start, int *gap_count, double *movers,
                     int mover_cap, double *sbufl, double *sbufr, int nbmax, int *counts,
                     int rank, int nvp, double *leftover, int leftover_cap, int nleft,
                     int *leftover_counts, double *scratch, int scratch_rows, int npool,
                     int *pool_owner, void *stream);
int skb_gap_insert(const double *rows, int n, skb_particles_t p, const int *gap_start,
                   int *gap_count, const skb_grid_t *grid, int order, int tlx, int tly,
                   double *leftover, int leftover_cap, int *counts, void *stream);
/* skb_gap_insert with the row count read on the device: min(*n_dev, nmax) rows */
int skb_gap_insert_counted(const double *rows, const int *n_dev, int nmax, skb_particles_t p,
                           const int *gap_start, int *gap_count, const skb_grid_t *grid,
                           int order, int tlx, int tly, double *leftover, int leftover_cap,
                           int *counts, void *stream);
/* One migration message of cppmove2 (pplib2.c:741-753: count + particle rows to a
 * neighbour) written straight into the neighbour's receive slot `dst` (peer memory):
 * dst[0] = n = min(*count, max_rows), dst[1..4] = 0, dst[5..] = the n AoS rows; the
 * count is read on the device, so no host synchronisation precedes the send. */
int skb_peer_send(const double *rows, const int *count, int max_rows, double *dst,
                  void *stream);
int skb_gap_densify(skb_particles_t in, skb_particles_t out, const int *gap_start,
                    const int *gap_count, const skb_grid_t *grid, int tlx, int tly,
                    int *dense_start, int *cell_end, int *tile_offsets, int *block_sums,
                    const double *leftover, int leftover_cap, int nleft,
                    long long n_in_cells, void *stream);
int skb_exclusive_scan(int *a, int n, int *block_sums, void *stream);
int skb_chunk_table(const int *tile_offsets, const skb_grid_t *grid, int tlx, int tly,
                    int chunk, int *chunk_first_tile, void *stream);

/* ---- guard cells (NumPy slicing in the reference) ----------------------------
 * nc = doubles per cell (1, 3 or 4).
 * skb_copy_guards: Field.copy_guards_y + copy_guards_x, field.py:73-98.
 *   from_below / from_above: packed [lby][nx][nc] rows received from the
 *   neighbours (the rank below's last active rows / the rank above's first active
 *   rows); NULL = periodic wrap inside this slab (single rank).
 * skb_add_guards: Sources.add_guards_x + add_guards_y + zeroing, sources.py:91-150.
 *   phase 0: x fold over all rows (run before the shear remap / y exchange);
 *   phase 1: y fold of `from_below` (the rank below's UPPER guard rows) and
 *            `from_above` (the rank above's LOWER guard rows), or of this slab's
 *            own guards when NULL, then zero every guard cell.
 * skb_pack_rows: copy rows [iy0, iy0+nrows) x active columns into a packed buffer. */
int skb_copy_guards(double *f, int nc, const skb_grid_t *grid,
                    const double *from_below, const double *from_above,
                    void *stream);
int skb_add_guards(double *f, int nc, const skb_grid_t *grid, int phase,
                   const double *from_below, const double *from_above,
                   void *stream);
int skb_pack_rows(const double *f, int nc, const skb_grid_t *grid, int iy0,
                  int nrows, double *out, void *stream);
/* refresh the x guards of one row after the spectral shear remap, field.py:169-171 */
int skb_copy_guards_x_rows(double *f, int nc, const skb_grid_t *grid, int iy0,
                           int nrows, void *stream);
/* whole-array scale of every component: Sources.normalize, sources.py:61-63 */
int skb_scale(double *f, long long n, double fac, void *stream);

/* ---- finite differences: finite_difference.pyx:5-85 --------------------------
 * f* point at the first element of a scalar plane with element stride `es`
 * doubles (1 for a scalar field, 3/4 for a component of an interleaved field). */
int skb_gradient(const double *f, int es, double *grad, const skb_grid_t *grid,
                 void *stream);
int skb_curl(const double *fx, const double *fy, const double *fz, int es,
             double *curl, const skb_grid_t *grid, int down, void *stream);
int skb_divergence(const double *fx, const double *fy, int es, double *div,
                   const skb_grid_t *grid, void *stream);
int skb_interp(const double *fx, const double *fy, const double *fz, int es,
               double *out, const skb_grid_t *grid, int up, void *stream);

/* Ohm.__call__ (ohm.py:35-75) fused into one pass over the active cells:
 * E = -alpha grad(log rho) + eta curl_down(B) + ((curl_down(B) - J)/rho) x unstagger(B).
 * Je_out / Bc_out (Float3, may be NULL) receive the reference's self.Je / self.B. */
int skb_ohm(const double *sources, const double *B, double *E, double *Je_out,
            double *Bc_out, const skb_grid_t *grid, double alpha, double eta,
            void *stream);
/* Faraday.__call__ (faraday.py:16-30): B -= dt * curl_up(E) on active cells */
int skb_faraday(const double *E, double *B, double *dB_out,
                const skb_grid_t *grid, double dt, void *stream);
/* ---- time-stepper field algebra on the device (horowitz.py:138-169) ------------------
 * Every call takes `skip` (device int, may be NULL): the kernel does nothing when *skip
 * != 0, so the iterations of the Horowitz loop can be queued ahead of the convergence
 * test.
 * skb_field_combine: out = c*(x + y) (mode 0: E2 = 0.5*(E3 + E), B2 = 0.5*(B3 + B)) or
 *   out = -x + c*y (mode 1), n doubles.
 * skb_faraday_to:    Bout = Bin - dt*curl_up(E) on the active cells ("B3 = B; faraday(E2,
 *   B3, dt)", horowitz.py:147-148, faraday.py:16-30).
 * skb_ohm_if:        skb_ohm with the skip flag.
 * skb_horowitz_update: E3 <- -E + 2*E2 (whole array) and acc[0] += sum over active cells
 *   and components of (E3_new - E3_old)^2 (calculate_diff, horowitz.py:112-121).
 * skb_converged:     state[0] = 1, state[1] = iter when sqrt(acc[0]*scale) < tol (and not
 *   already set). */
int skb_field_combine(double *out, const double *x, const double *y, long long n, double c,
                      int mode, const int *skip, void *stream);
int skb_faraday_to(const double *E, const double *Bin, double *Bout, double *dB_out,
                   const skb_grid_t *grid, double dt, const int *skip, void *stream);
int skb_ohm_if(const double *sources, const double *B, double *E, double *Je_out,
               double *Bc_out, const skb_grid_t *grid, double alpha, double eta,
               const int *skip, void *stream);
int skb_horowitz_update(double *E3, const double *E, const double *E2, const skb_grid_t *grid,
                        double *acc, const int *skip, void *stream);
int skb_converged(const double *acc, double scale, double tol, int iter, int *state,
                  void *stream);

/* ---- electrostatic field solve, k-space part -------------------------------------------
 * calc_form_factors + grad_inv_del (operators.pyx:13-135; ppic2's cppois22,
 * ppic2_wrapper.pyx:100-116) on the spectrum of a cuFFT real-to-complex transform:
 * q, fx, fy = [ny][nx/2+1] complex128.  fx, fy = -i k S(k)/k^2 q with the modes the
 * reference zeroes; *we += field energy (operators.pyx:135; zeroed by the caller, may be
 * NULL).  float32_quirk != 0 reproduces the reference's crealf / cimagf truncation
 * (SURVEY.md Q3).  The transforms themselves (cwpfft2rinit / cwppfft2r / cwppfft2r2,
 * ppic2_wrapper.pyx:69-142) are cuFFT calls of the host side. */
int skb_poisson_kspace(const double *q, double *fx, double *fy, int nx, int ny, double Lx,
                       double Ly, double ax, double ay, double affp, int float32_quirk,
                       double *we, void *stream);


#ifdef __cplusplus
}
#endif
#endif

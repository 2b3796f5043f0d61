// gapped.cu — gapped particle layout: sort maintenance without moving every particle.
//
// The tile sort's move pass costs 80 B/particle per step although only ~15 % of the
// particles change cell.  In the gapped layout every cell owns a slot range with some
// slack ([gap_start[k], gap_start[k+1]), live particles at the front, gap_count[k] of
// them).  The push then
//   * writes the particles that stay in their cell back into the same range, compacted
//     to the front (it has to write them anyway: 80 B/particle, as before),
//   * packs the ones that leave the slab into the exchange buffers (cppmove2's pack,
//     pplib2.c:666-707),
//   * parks the ones that move to another cell of the same thread block in a scratch
//     block (L2 resident) and, once all cells of the block are compacted, drops them
//     into the free slots of their new cells itself,
//   * appends the rest (they cross a block boundary) to a global mover list (AoS rows)
// and a small insertion kernel drops those and the arrivals into their new cells.  No
// histogram, no scan, no move pass: ~100 B/particle for push + ordering instead of
// 176 B.  When a cell's slack or the mover list overflows, nothing is lost: the
// particle goes to a small leftover list (pushed / deposited by the generic kernels,
// re-inserted every step) or stays parked in its old cell with a flag raised, and the
// caller rebuilds the layout through the dense path (densify -> tile sort -> build).
//
// The same kernel, instantiated with the half-step deposit of push_and_deposit
// (push_and_deposit.pyx:10-170) between two half drifts, serves the time steppers
// (skb_push_and_deposit_gapped).
//
// New component (no reference counterpart); ordering stays a performance property.
#include <cstdlib>
#include "gapped.cuh"

#ifndef GAP_MINB
#define GAP_MINB 3        // resident CTAs/SM aimed at for CIC (80 registers, no spills)
#endif
#define GAP_BLOCK 64      // particles per ring stage
__device__ __forceinline__ void gap_cp_async16(double *smem_dst, const double *gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gap_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gap_cp_wait_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void gap_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- build / densify ----------------------------------------------------------------
// capacity of a cell with n particles: 25 % slack, rounded up
__global__ void __launch_bounds__(256)
gap_caps_kernel(const int *__restrict__ cell_end, int ncells, int *gap_start) {
  int c = blockIdx.x * 256 + threadIdx.x;
  if (c > ncells) return;
  if (c == ncells) { gap_start[c] = 0; return; }
  const int n = cell_end[c] - (c ? cell_end[c - 1] : 0);
  // 25 % slack; big cells start on 128-byte lines, small ones only on even slots (the
  // 16-byte cp.async of the push needs that much)
  gap_start[c] = n < 64 ? (n + max(4, n >> 2) + 3) & ~3 : (n + (n >> 2) + 15) & ~15;
}

// dense (cell_end) -> gapped (gap_start already scanned); one warp per cell
__global__ void __launch_bounds__(256)
gap_fill_kernel(skb_particles_t in, skb_particles_t out, const int *__restrict__ cell_end,
                const int *__restrict__ gap_start, int *gap_count, int ncells,
                long long capacity) {
  // slot ranges do not fit the arrays: write nothing (the host reads gap_start[ncells])
  if (gap_start[ncells] > capacity || gap_start[ncells] < 0) return;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * 256) >> 5;
  for (int c = (blockIdx.x * 256 + threadIdx.x) >> 5; c < ncells; c += warps) {
    const int s = c ? cell_end[c - 1] : 0, n = cell_end[c] - s, d = gap_start[c];
    for (int i = lane; i < n; i += 32) {
      out.x[d + i] = in.x[s + i]; out.y[d + i] = in.y[s + i]; out.vx[d + i] = in.vx[s + i];
      out.vy[d + i] = in.vy[s + i]; out.vz[d + i] = in.vz[s + i];
    }
    if (lane == 0) gap_count[c] = n;
  }
}

// gapped -> dense: dense_start = exclusive scan of gap_count (array of ncells + 1 ints);
// writes cell_end and the tile offsets of the dense ordering
__global__ void __launch_bounds__(256)
gap_densify_kernel(skb_particles_t in, skb_particles_t out, const int *__restrict__ gap_start,
                   const int *__restrict__ gap_count, const int *__restrict__ dense_start,
                   int *cell_end, int *tile_offsets, int ncells, int cells_log2) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * 256) >> 5;
  for (int c = (blockIdx.x * 256 + threadIdx.x) >> 5; c < ncells; c += warps) {
    const int s = gap_start[c], n = gap_count[c], d = dense_start[c];
    for (int i = lane; i < n; i += 32) {
      out.x[d + i] = in.x[s + i]; out.y[d + i] = in.y[s + i]; out.vx[d + i] = in.vx[s + i];
      out.vy[d + i] = in.vy[s + i]; out.vz[d + i] = in.vz[s + i];
    }
    if (lane == 0) {
      cell_end[c] = d + n;
      if ((c & ((1 << cells_log2) - 1)) == 0) tile_offsets[c >> cells_log2] = d;
      if (c == ncells - 1) tile_offsets[ncells >> cells_log2] = d + n;
    }
  }
}

// leftover list (SoA, stride cap) -> tail of the dense arrays
__global__ void __launch_bounds__(256)
gap_append_rows_kernel(const double *__restrict__ lo, int cap, int n, skb_particles_t out,
                       long long at) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  out.x[at + i] = lo[i]; out.y[at + i] = lo[(size_t)cap + i];
  out.vx[at + i] = lo[2 * (size_t)cap + i]; out.vy[at + i] = lo[3 * (size_t)cap + i];
  out.vz[at + i] = lo[4 * (size_t)cap + i];
}

// pushed leftover particles (SoA) -> exchange buffers or the head of the mover list;
// cur != NULL (fused full-step deposit): the rows that stay in the slab are deposited
// here, one by one (the leavers are deposited by the rank that receives them)
template <int ORDER>
__global__ void __launch_bounds__(256)
gap_route_kernel(const double *__restrict__ lo, int cap, int n, double e0, double e1,
                 double ny, int rank, int nvp, double *movers, double *sbufl,
                 double *sbufr, int nbmax, int *counts, double *cur, DepParams dp,
                 DevGrid g) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double x = lo[i], vx = lo[2 * (size_t)cap + i], vy = lo[3 * (size_t)cap + i],
               vz = lo[4 * (size_t)cap + i];
  double y = lo[(size_t)cap + i];
  double *r;
  if (y < e0 || y >= e1) {
    int slot; double *buf;
    if (y < e0) {
      if (rank == 0) y += ny;
      slot = atomicAdd(counts + 1, 1); buf = sbufl;
    } else {
      if (rank == nvp - 1) y -= ny;
      slot = atomicAdd(counts + 2, 1); buf = sbufr;
    }
    if (slot >= nbmax) { atomicOr(counts + 3, 2); return; }
    r = buf + (size_t)slot * 5;
  } else {
    r = movers + (size_t)atomicAdd(counts + 0, 1) * 5;   // n <= mover_cap: always fits
    if (cur) {
      constexpr int NS = ORDER + 1;
      double xs = x + dp.offx, ys = y + dp.offy;
      if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
      int ix, iy;
      double wx[NS], wy[NS];
      particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
      single_particle_emit<NS>(wx, wy, ix, iy, vx + dp.S * (y * g.dy + g.y0), vy, vz, cur, g);
    }
  }
  r[0] = x; r[1] = y; r[2] = vx; r[3] = vy; r[4] = vz;
}

// deposit of AoS rows (arrivals from the neighbour ranks) with HBM atomics:
// deposit_particle_cic/tsc, deposit.pxd:3-118
template <int ORDER>
__global__ void __launch_bounds__(256)
gap_deposit_rows_kernel(const double *__restrict__ rows, int n, double *cur, DepParams dp,
                        DevGrid g) {
  constexpr int NS = ORDER + 1;
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double *r = rows + (size_t)i * 5;
  if (__double_as_longlong(r[0]) == GAP_PAD_BITS) return;
  double xs = r[0] + dp.offx, ys = r[1] + dp.offy;
  if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
  int ix, iy;
  double wx[NS], wy[NS];
  particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
  single_particle_emit<NS>(wx, wy, ix, iy, r[2] + dp.S * (r[1] * g.dy + g.y0), r[3], r[4],
                           cur, g);
}

__global__ void __launch_bounds__(256)
gap_insert_kernel(const double *__restrict__ rows, int n, skb_particles_t P,
                  const int *__restrict__ gap_start, int *gap_count, KeyParams kp,
                  double *leftover, int leftover_cap, int *counts) {
  gap_insert_rows(rows, n, blockIdx.x * (256 * GAP_INS_ITEMS) + threadIdx.x, P, gap_start,
                  gap_count, kp, leftover, leftover_cap, counts);
}

// ---- push on the gapped layout -----------------------------------------------------------
// PD = 0: push (push / push_modified + boundary epilogue)
// PD = 1: push_and_deposit, update = True   (gather at the old position, kick, half
//         drift, deposit, second half drift, x wrap, then the same routing as PD = 0)
// PD = 2: push_and_deposit, update = False  (deposit only: nothing is written back)
template <int ORDER, bool MODIFIED, int PD>
__global__ void __launch_bounds__(GAP_THREADS, (ORDER == 1 && PD == 0) ? GAP_MINB
                                               : (ORDER == 2 && PD != 0) ? 1 : 2)
push_gapped_kernel(skb_particles_t P, const double *__restrict__ E,
                   const double *__restrict__ B, DevGrid g, DevTiling tl, GapPush q,
                   GapDeposit dq, int parts, int wstride, int wrows) {
  constexpr int NS = ORDER + 1;
  extern __shared__ double smem[];
  double *sE = smem;
  double *sB = smem + (size_t)wstride * wrows * 3;
  // (PD) window of the sources grid, behind the E, B windows and the particle rings
  double *sw = sB + (size_t)wstride * wrows * 3 + (GAP_THREADS / 32) * (2 * 5 * GAP_BLOCK);
  if constexpr (PD != 0) zero_window(sw, wstride * wrows * 4);
  const int cells_log2 = tl.tlx + tl.tly;
  const int cpp = (1 << cells_log2) / parts;
  const int tile = blockIdx.x / parts;
  const int c0 = (tile << cells_log2) + (blockIdx.x % parts) * cpp;
  const Window w = tile_window(tile, tl, g);
  __shared__ int s_blk, s_nrows;
  if (threadIdx.x == 0) {
    int b = -1;
    if (q.npool > 0) {
      b = blockIdx.x % q.npool;
      while (atomicCAS(q.pool_owner + b, 0, 1) != 0) b = (b + 1 == q.npool) ? 0 : b + 1;
    }
    s_blk = b; s_nrows = 0;
  }
  stage_window(sE, E, w, wstride, g);
  stage_window(sB, B, w, wstride, g);
  __syncthreads();
  double *const scr = s_blk >= 0 ? q.scratch + (size_t)s_blk * q.scratch_rows * 5 : nullptr;
  const int scr_rows = s_blk >= 0 ? q.scratch_rows : 0;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int cpw = cpp / (GAP_THREADS / 32);
  const int wc0 = c0 + wv * cpw;
  const double nxd = (double)g.nx;
  const int mxm = (1 << tl.tlx) - 1, mym = (1 << tl.tly) - 1;
  const int tile_x0 = (tile % q.key.ntx) << tl.tlx, tile_y0 = (tile / q.key.ntx) << tl.tly;
  const int part_off = (blockIdx.x % parts) * cpp;
  // mover slots are reserved GAP_MCHUNK at a time per warp (one global atomic per chunk
  // instead of one per 32 particles on a single address)
  int mbase = 0, mused = GAP_MCHUNK;
  // per-warp ring of two 64-particle stages filled with cp.async (16 B per lane and
  // array: slot ranges start on multiples of 16 slots), so the loads of the next two
  // blocks are in flight, without holding registers, while a block is pushed
  double *ring = sB + (size_t)wstride * wrows * 3 + (size_t)wv * (2 * 5 * GAP_BLOCK);
  for (int cb = 0; cb < cpw; cb += 32) {
    const bool mine = cb + lane < cpw;
    const int my_cnt = mine ? q.gap_count[wc0 + cb + lane] : 0;
    const int my_start = mine ? tl.gap_start[wc0 + cb + lane] : 0;
    const int ncell = min(32, cpw - cb);
#define GAP_CNT(j) __shfl_sync(SKB_FULL, my_cnt, (j) & 31)
#define GAP_ADVANCE(j, base)                                   \
  do {                                                         \
    base += GAP_BLOCK;                                         \
    if (base >= GAP_CNT(j)) {                                  \
      base = 0; j++;                                           \
      while (j < ncell && GAP_CNT(j) == 0) j++;                \
    }                                                          \
  } while (0)
#define GAP_FETCH(stage, j, base)                                                       \
  do {                                                                                  \
    const int fs_ = __shfl_sync(SKB_FULL, my_start, (j) & 31);                          \
    const int fi_ = (base) + 2 * lane;                                                  \
    if (fi_ < GAP_CNT(j)) {                                                             \
      double *d_ = ring + (stage) * (5 * GAP_BLOCK) + 2 * lane;                         \
      const long long o_ = (long long)fs_ + fi_;                                        \
      gap_cp_async16(d_, P.x + o_); gap_cp_async16(d_ + GAP_BLOCK, P.y + o_);           \
      gap_cp_async16(d_ + 2 * GAP_BLOCK, P.vx + o_);                                    \
      gap_cp_async16(d_ + 3 * GAP_BLOCK, P.vy + o_);                                    \
      gap_cp_async16(d_ + 4 * GAP_BLOCK, P.vz + o_);                                    \
    }                                                                                   \
  } while (0)
    int fj = 0, fbase = 0;
    while (fj < ncell && GAP_CNT(fj) == 0) fj++;
    int cj = fj, cbase = 0;
#pragma unroll
    for (int st = 0; st < 2; st++) {
      if (fj < ncell) { GAP_FETCH(st, fj, fbase); GAP_ADVANCE(fj, fbase); }
      gap_cp_commit();
    }
    int stage = 0, wcur = 0;                         // wcur: stayers written so far
    Acc<NS> acc;                                     // (PD) stencil sums of the current cell
    while (cj < ncell) {
      const int cell = wc0 + cb + cj;
      const int cix = tile_x0 + ((cell - (tile << cells_log2)) & mxm);
      const int ciy = tile_y0 + ((cell - (tile << cells_log2)) >> tl.tlx);
      const int s = __shfl_sync(SKB_FULL, my_start, cj);
      const int n = GAP_CNT(cj);
      if constexpr (PD != 0) {
        if (cbase == 0) {                            // first block of a cell
#pragma unroll
          for (int i = 0; i < NS * NS * 4; i++) acc.v[i] = 0.0;
          acc.ix = cix; acc.iy = ciy;
        }
      }
      gap_cp_wait_one();
      __syncwarp();
      const double *pb = ring + stage * (5 * GAP_BLOCK);
#pragma unroll 1
      for (int u = 0; u < GAP_BLOCK / 32; u++) {
        if (cbase + u * 32 >= n) break;
        const int i = cbase + u * 32 + lane;
        const bool act = i < n;
        bool stay = false, mover = false, local = false;
        double x = 0, y = 0, vx = 0, vy = 0, vz = 0;
        if (act) {
          const double *pp = pb + u * 32 + lane;
          x = pp[0]; y = pp[GAP_BLOCK]; vx = pp[2 * GAP_BLOCK]; vy = pp[3 * GAP_BLOCK];
          vz = pp[4 * GAP_BLOCK];
          if constexpr (PD == 0) {
            fields_and_kick<ORDER, MODIFIED>(sE, sB, w, wstride, E, B, g, q.k, x, y, vx, vy, vz);
            x = x + vx * q.dtdsx;                    // drift_particle, particle_push.pxd:88-91
            y = y + vy * q.dtdsy;
            if (q.flags & SKB_EPI_SHEAR) {           // particle_boundary.pyx:41-49
              if (y < 0.0) { x = x - q.x_boost; vx = vx - q.vx_boost; }
              if (y >= (double)g.ny) { x = x + q.x_boost; vx = vx + q.vx_boost; }
            }
            if (q.flags & SKB_EPI_PERIODIC_X) x = wrap_x(x, nxd);
          } else {
            const double xold = x, yold = y;
            fields_and_kick<ORDER, false>(sE, sB, w, wstride, E, B, g, q.k, x, y, vx, vy, vz);
            x = x + vx * dq.d2x;                     // first half of the drift
            y = y + vy * dq.d2y;
            // more than half a cell in half a step: push_and_deposit.pyx:66-68
            if (fabs(x - xold) > 0.5 || fabs(y - yold) > 0.5) atomicOr(q.counts + 3, 4);
            double xs = x + dq.dp.offx, ys = y + dq.dp.offy;
            if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
            int ix, iy;
            double wx[NS], wy[NS];
            particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
            const double vxr = vx + dq.dp.S * (y * g.dy + g.y0);
            if (ix == acc.ix && iy == acc.iy) accumulate<ORDER>(acc, wx, wy, vxr, vy, vz);
            else stray_particle_emit<NS>(wx, wy, ix, iy, vxr, vy, vz, sw, w, wstride, dq.cur, g);
            if constexpr (PD == 1) {
              x = x + vx * dq.d2x;                   // second half of the drift
              y = y + vy * dq.d2y;
              x = wrap_x(x, nxd);
            }
          }
          if constexpr (PD == 2) {
            // predictor sweep: particles stay untouched
          } else if (y < g.e0 || y >= g.e1) {        // leaves the slab: cppmove2's pack
            double *buf; int slot; double yy = y;
            if (yy < g.e0) {
              if (q.rank == 0) yy += (double)g.ny;
              slot = atomicAdd(q.counts + 1, 1); buf = q.sbufl;
            } else {
              if (q.rank == q.nvp - 1) yy -= (double)g.ny;
              slot = atomicAdd(q.counts + 2, 1); buf = q.sbufr;
            }
            if (slot < q.nbmax) {
              double *r = buf + (size_t)slot * 5;
              r[0] = x; r[1] = yy; r[2] = vx; r[3] = vy; r[4] = vz;
            } else {
              atomicOr(q.counts + 3, 2);
            }
          } else {
            // cell_key() == cell, spelled out on the cell coordinates (cheaper than
            // the tile-major key itself; phase B / skb_gap_insert compute that)
            double xs = x + q.key.offx, ys = y + q.key.offy;
            if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
            const int ix = min(max((int)xs, 0), q.key.mx - 1);
            const int iy = min(max((int)ys, 0), q.key.myp - 1);
            stay = ix == cix && iy == ciy;
            mover = !stay;
            const int loc = (((iy - tile_y0) << tl.tlx) | (ix - tile_x0)) - part_off;
            local = mover && (unsigned)(ix - tile_x0) <= (unsigned)mxm &&
                    (unsigned)(iy - tile_y0) <= (unsigned)mym && (unsigned)loc < (unsigned)cpp;
          }
        }
        if constexpr (PD == 2) continue;
        // movers inside this CTA's cell range: scratch block
        const unsigned lm = __ballot_sync(SKB_FULL, local && scr_rows > 0);
        if (lm) {
          int b0 = 0;
          if (lane == __ffs(lm) - 1) b0 = atomicAdd(&s_nrows, __popc(lm));
          b0 = __shfl_sync(SKB_FULL, b0, __ffs(lm) - 1);
          if (local) {
            const int slot = b0 + __popc(lm & lt);
            if (slot < scr_rows) {
              double *o = scr + (size_t)slot * 5;
              o[0] = x; o[1] = y; o[2] = vx; o[3] = vy; o[4] = vz;
              mover = false;
            }
          }
        }
        // the other movers: global list, slots from the warp's reservation
        const unsigned mm = __ballot_sync(SKB_FULL, mover);
        if (mm) {
          const int k = __popc(mm), room = GAP_MCHUNK - mused;
          int nb = mbase;
          if (k > room) {
            if (lane == 0) nb = atomicAdd(q.counts + 0, GAP_MCHUNK);
            nb = __shfl_sync(SKB_FULL, nb, 0);
          }
          if (mover) {
            const int r = __popc(mm & lt);
            const int slot = r < room ? mbase + mused + r : nb + (r - room);
            if (slot < q.mover_cap) {
              double *o = q.movers + (size_t)slot * 5;
              o[0] = x; o[1] = y; o[2] = vx; o[3] = vy; o[4] = vz;
            } else {
              stay = true;                           // list full: park it here, rebuild later
              atomicOr(q.counts + 3, 1);
            }
          }
          if (k > room) { mbase = nb; mused = k - room; } else mused += k;
        }
        // stayers: compacted to the front of the cell's range (always behind the reads:
        // wcur <= cbase + u*32, and the blocks in flight start at cbase + 64 or later)
        const unsigned sm = __ballot_sync(SKB_FULL, stay);
        if (stay) {
          const long long d = (long long)s + wcur + __popc(sm & lt);
          P.x[d] = x; P.y[d] = y; P.vx[d] = vx; P.vy[d] = vy; P.vz[d] = vz;
        }
        wcur += __popc(sm);
      }
      __syncwarp();                                  // stage fully read: refill it
      int nj = cj, nbase = cbase;
      GAP_ADVANCE(nj, nbase);
      if (nj != cj) {                                // the cell is finished
        if constexpr (PD != 2) {
          if (lane == 0) q.gap_count[cell] = wcur;
          wcur = 0;
        }
        if constexpr (PD != 0) {
          // one warp reduction and one emit per cell (see deposit_cells_kernel)
          const int lo = (NS == 3) ? 1 : 0;
          const bool in_window = (acc.ix - lo >= w.x0) && (acc.ix - lo + NS <= w.x1) &&
                                 (acc.iy - lo >= w.y0) && (acc.iy - lo + NS <= w.y1);
          if constexpr (NS == 2) {
            warp_reduce_scatter<16>(acc.v, lane);
            if (lane < 16)
              emit_one<NS>(acc.v[0], scatter_index<16>(lane), in_window, acc.ix, acc.iy, sw, w,
                           wstride, dq.cur, g);
          } else {
            warp_reduce_scatter<32>(acc.v, lane);
            warp_reduce_scatter<4>(acc.v + 32, lane);
            emit_one<NS>(acc.v[0], scatter_index<32>(lane), in_window, acc.ix, acc.iy, sw, w,
                         wstride, dq.cur, g);
            if (lane < 4)
              emit_one<NS>(acc.v[32], 32 + scatter_index<4>(lane), in_window, acc.ix, acc.iy, sw,
                           w, wstride, dq.cur, g);
          }
        }
      }
      cj = nj; cbase = nbase;
      if (fj < ncell) { GAP_FETCH(stage, fj, fbase); GAP_ADVANCE(fj, fbase); }
      gap_cp_commit();
      stage ^= 1;
    }
    gap_cp_wait_all();
#undef GAP_CNT
#undef GAP_ADVANCE
#undef GAP_FETCH
  }
  // phase B: all cells of this CTA are compacted; drop the parked movers into them
  __syncthreads();
  if constexpr (PD != 0) flush_window(sw, w, wstride, dq.cur, g);
  if (scr_rows > 0) {
    const int nrows = min(s_nrows, scr_rows);
    if (threadIdx.x == 0 && nrows) atomicAdd(q.counts + 4, nrows);   // statistics
    for (int r0 = 0; r0 < nrows; r0 += GAP_THREADS * GAP_INS_ITEMS)
      gap_insert_rows(scr, nrows, r0 + (int)threadIdx.x, P, q.gap_start, q.gap_count, q.key,
                      q.leftover, q.leftover_cap, q.lcounts);
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(q.pool_owner + s_blk, 0);
  }
  // unused tail of this warp's last reservation: padding rows
  for (int r = mused + lane; r < GAP_MCHUNK; r += 32)
    if (mbase + r < q.mover_cap)
      q.movers[(size_t)(mbase + r) * 5] = __longlong_as_double(GAP_PAD_BITS);
}

// ---- C ABI ---------------------------------------------------------------------------------
extern "C" int skb_exclusive_scan(int *a, int n, int *block_sums, void *stream);
extern "C" int skb_tile_geometry(const skb_grid_t *grid, int tlx, int tly, int *ntx, int *nty);

static int gap_ncells(const skb_grid_t *grid, int tlx, int tly) {
  int ntx, nty;
  skb_tile_geometry(grid, tlx, tly, &ntx, &nty);
  return (ntx * nty) << (tlx + tly);
}

extern "C" int skb_gap_build(skb_particles_t in, skb_particles_t out, const int *cell_end,
                             const skb_grid_t *grid, int tlx, int tly, int *gap_start,
                             int *gap_count, int *block_sums, long long capacity,
                             void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int ncells = gap_ncells(grid, tlx, tly);
  gap_caps_kernel<<<(ncells + 256) / 256, 256, 0, st>>>(cell_end, ncells, gap_start);
  SKB_CHECK_LAUNCH();
  int rc = skb_exclusive_scan(gap_start, ncells + 1, block_sums, stream);
  if (rc) return rc;
  int blocks = min((ncells + 7) / 8, 148 * 16);
  gap_fill_kernel<<<blocks, 256, 0, st>>>(in, out, cell_end, gap_start, gap_count, ncells,
                                          capacity);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_gap_densify(skb_particles_t in, skb_particles_t out, const int *gap_start,
                               const int *gap_count, const skb_grid_t *grid, int tlx,
                               int tly, int *dense_start, int *cell_end, int *tile_offsets,
                               int *block_sums, const double *leftover, int leftover_cap,
                               int nleft, long long n_in_cells, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int ncells = gap_ncells(grid, tlx, tly);
  cudaError_t e = cudaMemcpyAsync(dense_start, gap_count, sizeof(int) * (size_t)ncells,
                                  cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(dense_start + ncells, 0, sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  int rc = skb_exclusive_scan(dense_start, ncells + 1, block_sums, stream);
  if (rc) return rc;
  int blocks = min((ncells + 7) / 8, 148 * 16);
  gap_densify_kernel<<<blocks, 256, 0, st>>>(in, out, gap_start, gap_count, dense_start,
                                             cell_end, tile_offsets, ncells, tlx + tly);
  SKB_CHECK_LAUNCH();
  if (nleft > 0) {
    gap_append_rows_kernel<<<(nleft + 255) / 256, 256, 0, st>>>(leftover, leftover_cap,
                                                                nleft, out, n_in_cells);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}

// the same with the row count read on the device (no host round trip between the kernel
// that produced the rows and their insertion): a fixed grid strides over the rows
#ifdef GAP_INS_MINB                  // (tuning knob of tools/build_variants.py)
__global__ void __launch_bounds__(256, GAP_INS_MINB)
#else
__global__ void __launch_bounds__(256)
#endif
gap_insert_counted_kernel(const double *__restrict__ rows, const int *__restrict__ n_dev,
                          int nmax, skb_particles_t P, const int *__restrict__ gap_start,
                          int *gap_count, KeyParams kp, double *leftover, int leftover_cap,
                          int *counts) {
  const int n = min(*n_dev, nmax);
  const int per = 256 * GAP_INS_ITEMS;
  for (int b = blockIdx.x; (long long)b * per < n; b += gridDim.x)
    gap_insert_rows(rows, n, b * per + threadIdx.x, P, gap_start, gap_count, kp, leftover,
                    leftover_cap, counts);
}

extern "C" int skb_gap_insert_counted(const double *rows, const int *n_dev, int nmax,
                                      skb_particles_t p, const int *gap_start,
                                      int *gap_count, const skb_grid_t *grid, int order,
                                      int tlx, int tly, double *leftover, int leftover_cap,
                                      int *counts, void *stream) {
  if (nmax <= 0) return 0;
  DevGrid g = make_grid(grid);
  KeyParams kp = make_keyparams(g, order, tlx, tly);
  const int per = 256 * GAP_INS_ITEMS;
  const int nblk = (int)min((long long)(nmax + per - 1) / per, 148LL * 16);
  gap_insert_counted_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(
      rows, n_dev, nmax, p, gap_start, gap_count, kp, leftover, leftover_cap, counts);
  SKB_CHECK_LAUNCH();
  return 0;
}

// Migration message into a neighbour's receive slot (peer memory over NVLink, or local):
// header row {count, 0, 0, 0, 0} + count rows, the count read on the device
// (pplib2.c:741-753 sends the count in a message of its own).
__global__ void __launch_bounds__(256)
peer_send_kernel(const double *__restrict__ rows, const int *__restrict__ count,
                 int max_rows, double *__restrict__ dst) {
  const int n = min(max(*count, 0), max_rows);
  const long long nd = (long long)n * 5;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 < 5) dst[i0] = i0 == 0 ? (double)n : 0.0;
  for (long long i = i0; i < nd; i += (long long)gridDim.x * blockDim.x) dst[5 + i] = rows[i];
}

extern "C" int skb_peer_send(const double *rows, const int *count, int max_rows, double *dst,
                             void *stream) {
  peer_send_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(rows, count, max_rows, dst);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_gap_insert(const double *rows, int n, skb_particles_t p,
                              const int *gap_start, int *gap_count, const skb_grid_t *grid,
                              int order, int tlx, int tly, double *leftover,
                              int leftover_cap, int *counts, void *stream) {
  if (n <= 0) return 0;
  DevGrid g = make_grid(grid);
  KeyParams kp = make_keyparams(g, order, tlx, tly);
  const int per = 256 * GAP_INS_ITEMS;
  gap_insert_kernel<<<(n + per - 1) / per, 256, 0, (cudaStream_t)stream>>>(
      rows, n, p, gap_start, gap_count, kp, leftover, leftover_cap, counts);
  SKB_CHECK_LAUNCH();
  return 0;
}

size_t cell_stream_smem(int pd);
int cell_stream_launch(int pd, int order, int modified, skb_particles_t p, const double *E,
                       const double *B, const DevGrid &g, const GapPush &q,
                       const GapDeposit &dq, int ntiles, cudaStream_t st);

// pd: 0 = push, 1 = push_and_deposit with update, 2 = push_and_deposit without update,
// 3 = push + the full-step deposit that follows it (sources.py:27-50), raw sums
static int gapped_sweep(int pd, skb_particles_t p, const double *E, const double *B,
                        const skb_grid_t *grid, int order, double qtmh, double dt,
                        int modified, double Omega, double S, int epi_flags, double epi_S,
                        double epi_t, double *current, double dep_S, int *ihole, int ntmax,
                        int tlx, int tly, const int *gap_start, int *gap_count,
                        double *movers, int mover_cap, double *sbufl, double *sbufr,
                        int nbmax, int *counts, int rank, int nvp, double *leftover,
                        int leftover_cap, int nleft, int *leftover_counts, double *scratch,
                        int scratch_rows, int npool, int *pool_owner, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (order != 1 && order != 2) return (int)cudaErrorInvalidValue;
  DevGrid g = make_grid(grid);
  skb_tiling_t t = {};
  skb_tile_geometry(grid, tlx, tly, &t.ntx, &t.nty);
  t.tlx = tlx; t.tly = tly; t.chunk = 2048;
  t.gap_start = gap_start; t.gap_count = gap_count;
  DevTiling tl = make_tiling(&t);
  GapPush q;
  q.k = make_kick(g, qtmh, dt, Omega, S);
  q.dtdsx = dt / g.dx; q.dtdsy = dt / g.dy;
  q.vx_boost = epi_S * g.Ly;                       // particle_boundary.pyx:37-38
  q.x_boost = q.vx_boost * epi_t / g.dx;
  q.flags = epi_flags;
  q.key = make_keyparams(g, order, tlx, tly);
  q.gap_count = gap_count; q.movers = movers; q.mover_cap = mover_cap;
  q.sbufl = sbufl; q.sbufr = sbufr; q.nbmax = nbmax; q.rank = rank; q.nvp = nvp;
  q.counts = counts;
  q.scratch = scratch; q.scratch_rows = scratch_rows & ~1;   // even: 16-byte rows for TMA
  q.npool = (pd != 2 && scratch && pool_owner && scratch_rows > 1) ? npool : 0;
  q.pool_owner = pool_owner; q.gap_start = gap_start;
  q.leftover = leftover; q.leftover_cap = leftover_cap; q.lcounts = leftover_counts;
  GapDeposit dq = {};
  dq.cur = current;
  dq.dp.offx = q.k.offEx;                          // offsetE reused, push_and_deposit.pyx:71
  dq.dp.offy = q.k.offEy;
  if (pd == 3) {                                   // deposit.pyx:14-15
    dq.dp.offx = q.key.offx;
    dq.dp.offy = q.key.offy;
  }
  dq.dp.S = dep_S;
  dq.d2x = 0.5 * dt / g.dx;
  dq.d2y = 0.5 * dt / g.dy;
  // the cell-stream kernel (cellstream.cu) serves push and push + deposit on 16 x 16
  // tiles; SKB_GAP_GENERIC=1 keeps the generic kernel below (A/B measurements)
  static const bool force_generic = getenv("SKB_GAP_GENERIC") && atoi(getenv("SKB_GAP_GENERIC"));
  const bool stream_kernel = (pd == 0 || pd == 3) && tlx == 4 && tly == 4 && !force_generic;
  if ((pd == 3 || (epi_flags & SKB_EPI_DRIFT_ONLY)) && !stream_kernel)
    return (int)cudaErrorNotSupported;
  cudaError_t e = cudaMemsetAsync(counts, 0, 5 * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  if (q.npool > 0) {       // (all blocks free: a failed earlier launch cannot leave claims)
    e = cudaMemsetAsync(pool_owner, 0, sizeof(int) * (size_t)q.npool, st);
    if (e != cudaSuccess) return (int)e;
  }
  const int ntiles = t.ntx * t.nty;
  const int cells = 1 << (tlx + tly);
  // at least two blocks per tile: the parked movers of the resident blocks then fit L2
  int parts = cells >= 32 ? 2 : 1;
  while (parts < cells / 8 && (long long)ntiles * parts < 8 * 148) parts <<= 1;
  if (const char *ev = getenv("SKB_GAP_PARTS")) {          // tuning aid
    const int pv = atoi(ev);
    if (pv >= 1 && pv <= cells / 8 && (pv & (pv - 1)) == 0) parts = pv;
  }
  const int ws = window_stride(tl), wr = window_rows(tl);
  // E and B windows + the warps' particle rings (+ the window of the sources grid)
  const size_t smem = ((size_t)ws * wr * (3 * 2 + (pd ? 4 : 0)) +
                       (GAP_THREADS / 32) * 2 * 5 * GAP_BLOCK) * sizeof(double);
  // 16-byte cp.async: the five arrays must be 16-byte aligned (and every gap_start even,
  // as skb_gap_build guarantees)
  if ((((uintptr_t)p.x | (uintptr_t)p.y | (uintptr_t)p.vx | (uintptr_t)p.vy |
        (uintptr_t)p.vz) & 15) != 0)
    return (int)cudaErrorMisalignedAddress;
  void (*k)(skb_particles_t, const double *, const double *, DevGrid, DevTiling, GapPush,
            GapDeposit, int, int, int) = nullptr;
  if (stream_kernel) {
  } else if (pd == 0) {
    if (order == 1) k = modified ? push_gapped_kernel<1, true, 0> : push_gapped_kernel<1, false, 0>;
    else k = modified ? push_gapped_kernel<2, true, 0> : push_gapped_kernel<2, false, 0>;
  } else if (pd == 1) {
    k = order == 1 ? push_gapped_kernel<1, false, 1> : push_gapped_kernel<2, false, 1>;
  } else {
    k = order == 1 ? push_gapped_kernel<1, false, 2> : push_gapped_kernel<2, false, 2>;
  }
  if (!stream_kernel) {
    // the scratch-pool claim spins until a block is free: the pool must hold at least as
    // many blocks as CTAs can be resident
    if (q.npool > 0) {
      int dev = 0, sms = 0, per = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
      }
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k, GAP_THREADS, smem);
      if (e != cudaSuccess) return (int)e;
      if (q.npool < max(per, 1) * max(sms, 1)) return (int)cudaErrorInvalidValue;
    } else if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
  }
  if (pd == 1 || pd == 2) {  // in-band flag of the generic kernel below (ihole[0] = -1)
    e = cudaMemsetAsync(ihole, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  if (nleft > 0) {
    // the particles that found no slot last step: generic kernel in place, then onto
    // the head of the mover list (or out to the neighbours) like everybody else
    if (nleft > leftover_cap || nleft > mover_cap) return (int)cudaErrorInvalidValue;
    const size_t lc = (size_t)leftover_cap;
    skb_particles_t lp = {leftover, leftover + lc, leftover + 2 * lc, leftover + 3 * lc,
                          leftover + 4 * lc};
    int rc;
    if (pd == 1 || pd == 2) {
      if (nleft > ntmax) return (int)cudaErrorInvalidValue;
      rc = skb_push_and_deposit(lp, nleft, E, B, grid, order, qtmh, dt, ihole, ntmax, current,
                                dep_S, pd == 1, nullptr, nullptr, tlx, tly, stream);
    } else {
      skb_epilogue_t ep = {};
      ep.flags = epi_flags & (SKB_EPI_SHEAR | SKB_EPI_PERIODIC_X);
      ep.S = epi_S; ep.t = epi_t;
      if (epi_flags & SKB_EPI_DRIFT_ONLY)
        rc = skb_drift(lp, nleft, dt, grid, &ep, stream);
      else
        rc = skb_boris_push(lp, nleft, E, B, grid, order, qtmh, dt, modified, Omega, S,
                            nullptr, &ep, stream);
    }
    if (rc) return rc;
    if (pd != 2) {
      double *rc_cur = pd == 3 ? current : nullptr;
      if (order == 1)
        gap_route_kernel<1><<<(nleft + 255) / 256, 256, 0, st>>>(
            leftover, leftover_cap, nleft, g.e0, g.e1, (double)g.ny, rank, nvp, movers, sbufl,
            sbufr, nbmax, counts, rc_cur, dq.dp, g);
      else
        gap_route_kernel<2><<<(nleft + 255) / 256, 256, 0, st>>>(
            leftover, leftover_cap, nleft, g.e0, g.e1, (double)g.ny, rank, nvp, movers, sbufl,
            sbufr, nbmax, counts, rc_cur, dq.dp, g);
      SKB_CHECK_LAUNCH();
    }
  }
  if (pd != 2) {
    // the old leftovers are consumed: the list restarts (this kernel and the
    // skb_gap_insert calls that follow append to it)
    e = cudaMemsetAsync(leftover_counts, 0, 2 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  if (stream_kernel)
    return cell_stream_launch(pd, order, modified, p, E, B, g, q, dq, ntiles, st);
  k<<<ntiles * parts, GAP_THREADS, smem, st>>>(p, E, B, g, tl, q, dq, parts, ws, wr);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_push_gapped(skb_particles_t p, const double *E, const double *B,
                               const skb_grid_t *grid, int order, double qtmh, double dt,
                               int modified, double Omega, double S, int epi_flags,
                               double epi_S, double epi_t, int tlx, int tly,
                               const int *gap_start, int *gap_count, double *movers,
                               int mover_cap, double *sbufl, double *sbufr, int nbmax,
                               int *counts, int rank, int nvp, double *leftover,
                               int leftover_cap, int nleft, int *leftover_counts,
                               double *scratch, int scratch_rows, int npool,
                               int *pool_owner, void *stream) {
  return gapped_sweep(0, p, E, B, grid, order, qtmh, dt, modified, Omega, S, epi_flags, epi_S,
                      epi_t, nullptr, 0.0, nullptr, 0, tlx, tly, gap_start, gap_count, movers,
                      mover_cap, sbufl, sbufr, nbmax, counts, rank, nvp, leftover,
                      leftover_cap, nleft, leftover_counts, scratch, scratch_rows, npool,
                      pool_owner, stream);
}

extern "C" int skb_push_and_deposit_gapped(
    skb_particles_t p, const double *E, const double *B, const skb_grid_t *grid, int order,
    double qtmh, double dt, double *current, double S, int update, int *ihole, int ntmax,
    int tlx, int tly, const int *gap_start, int *gap_count, double *movers, int mover_cap,
    double *sbufl, double *sbufr, int nbmax, int *counts, int rank, int nvp, double *leftover,
    int leftover_cap, int nleft, int *leftover_counts, double *scratch, int scratch_rows,
    int npool, int *pool_owner, void *stream) {
  return gapped_sweep(update ? 1 : 2, p, E, B, grid, order, qtmh, dt, 0, 0.0, 0.0,
                      SKB_EPI_PERIODIC_X, 0.0, 0.0, current, S, ihole, ntmax, tlx, tly,
                      gap_start, gap_count, movers, mover_cap, sbufl, sbufr, nbmax, counts,
                      rank, nvp, leftover, leftover_cap, nleft, leftover_counts, scratch,
                      scratch_rows, npool, pool_owner, stream);
}

// push / push_modified + the full-step deposit that follows it in the time loop
// (particles.py:159-188 then sources.py:27-50, deposit.pyx:6-34), one sweep: `current`
// receives the raw stencil sums of every particle that is in the slab after the push
// (arrivals from the neighbour ranks: skb_deposit_rows).  counts[3] & 16: some rows
// bypassed the fused deposit (scratch block full) - redo it with skb_deposit.
extern "C" int skb_push_deposit_gapped(
    skb_particles_t p, const double *E, const double *B, const skb_grid_t *grid, int order,
    double qtmh, double dt, int modified, double Omega, double S, int epi_flags, double epi_S,
    double epi_t, double *current, double dep_S, int tlx, int tly, const int *gap_start,
    int *gap_count, double *movers, int mover_cap, double *sbufl, double *sbufr, int nbmax,
    int *counts, int rank, int nvp, double *leftover, int leftover_cap, int nleft,
    int *leftover_counts, double *scratch, int scratch_rows, int npool, int *pool_owner,
    void *stream) {
  return gapped_sweep(3, p, E, B, grid, order, qtmh, dt, modified, Omega, S, epi_flags, epi_S,
                      epi_t, current, dep_S, nullptr, 0, tlx, tly, gap_start, gap_count,
                      movers, mover_cap, sbufl, sbufr, nbmax, counts, rank, nvp, leftover,
                      leftover_cap, nleft, leftover_counts, scratch, scratch_rows, npool,
                      pool_owner, stream);
}

// deposit_cic/tsc (deposit.pyx:6,21) of n AoS rows {x, y, vx, vy, vz} (e.g. the arrivals
// of a migration step) into `current`, HBM atomics
extern "C" int skb_deposit_rows(const double *rows, int n, double *current,
                                const skb_grid_t *grid, int order, double S, void *stream) {
  if (n <= 0) return 0;
  if (order != 1 && order != 2) return (int)cudaErrorInvalidValue;
  DevGrid g = make_grid(grid);
  DepParams dp;
  dp.offx = g.lbx - 0.5;              // deposit.pyx:14-15
  dp.offy = g.lby - 0.5 - g.noff;
  dp.S = S;
  if (order == 1)
    gap_deposit_rows_kernel<1><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rows, n, current, dp, g);
  else
    gap_deposit_rows_kernel<2><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rows, n, current, dp, g);
  SKB_CHECK_LAUNCH();
  return 0;
}

extern "C" int skb_drift_gapped(skb_particles_t p, const skb_grid_t *grid, int order, double dt,
                                int tlx, int tly, const int *gap_start, int *gap_count,
                                double *movers, int mover_cap, double *sbufl, double *sbufr,
                                int nbmax, int *counts, int rank, int nvp, double *leftover,
                                int leftover_cap, int nleft, int *leftover_counts,
                                double *scratch, int scratch_rows, int npool, int *pool_owner,
                                void *stream) {
  // (E, B are never read with SKB_EPI_DRIFT_ONLY)
  return gapped_sweep(0, p, nullptr, nullptr, grid, order, 0.0, dt, 0, 0.0, 0.0,
                      SKB_EPI_PERIODIC_X | SKB_EPI_DRIFT_ONLY, 0.0, 0.0, nullptr, 0.0, nullptr, 0,
                      tlx, tly, gap_start, gap_count, movers, mover_cap, sbufl, sbufr, nbmax,
                      counts, rank, nvp, leftover, leftover_cap, nleft, leftover_counts, scratch,
                      scratch_rows, npool, pool_owner, stream);
}

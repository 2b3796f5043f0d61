// deposit.cu — CIC / TSC charge and current deposition, and the fused
// push_and_deposit.
//
// Replaces deposit_cic/tsc (reference skeletor/cython/deposit.pyx:6-34,
// deposit.pxd:3-118) and push_and_deposit_cic/tsc (push_and_deposit.pyx:10-170).
//
// Design (no per-particle atomics): particles are ordered by the cell of their
// deposit stencil base (skb_tile_sort), so consecutive particles hit the same
// 2x2 (CIC) / 3x3 (TSC) cells.  Each lane streams particles with coalesced loads
// (lane-interleaved: lane l takes particles l, l+32, ... of its warp's range) and
// accumulates the NS*NS*4 stencil sums of its CURRENT cell in registers.  When
// any lane of the warp moves on to a new cell, the warp does one segmented
// reduction over lanes (shuffles; segments = runs of lanes with equal cell) and
// only the last lane of each run adds its totals to the CTA's shared-memory
// window of the source grid.  At the end of a tile segment the window is added to
// HBM once (coalesced rows, zero entries skipped).  Stencils that fall outside
// the window (stale ordering, unsorted tail) go to HBM directly.  Per-particle
// weights are formed exactly as the reference does (ty*tx, then *vx, ...), so the
// only difference to the reference is the order of the summation.
#include <cstdlib>
#include "deposit.cuh"


template <int ORDER, int MODE>
__global__ void __launch_bounds__(DEP_THREADS)
deposit_kernel(skb_particles_t P, long long np, const double *__restrict__ E,
               const double *__restrict__ B, double *__restrict__ cur, DevGrid g,
               DevTiling tl, DepParams q, FusedParams fq, int span, int wstride,
               int wrows) {
  constexpr int NS = ORDER + 1;
  extern __shared__ double smem[];
  double *sw = smem;                         // source window, 4 doubles per cell
  const int wcells = wstride * wrows * 4;
  double *sE = smem + wcells;                // MODE > 0: E and B windows
  double *sB = sE + wstride * wrows * 3;

  Acc<NS> a;
#pragma unroll
  for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
  a.iy = ACC_EMPTY; a.ix = 0;

  SegmentIter it;
  it.init(tl, np, span);
  long long s0, s1;
  int tile;
  while (it.next(tl, s0, s1, tile)) {
    Window w = tile_window(tile, tl, g);
    if (tile >= 0) {
      zero_window(sw, wcells);
      if (MODE > 0) {
        stage_window(sE, E, w, wstride, g);
        stage_window(sB, B, w, wstride, g);
      }
      __syncthreads();
    }
    // Uniform trip count (every lane takes part in the warp collectives) and a
    // warp-contiguous mapping: warp wv owns particles [wbeg, wend) and walks
    // through them 32 at a time, i.e. through consecutive cells.
    const long long n = s1 - s0;
    const int nwarps = DEP_THREADS / 32;
    const long long per_warp = ((n + nwarps - 1) / nwarps + 31) & ~31LL;
    const int wv = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wbeg = s0 + (long long)wv * per_warp;
    const long long wend = min(s1, wbeg + per_warp);
    const long long witers = per_warp >> 5;
    for (long long k = 0; k < witers; k++) {
      const long long i = wbeg + k * 32 + lane;
      const bool act = i < wend;
      double x = 0, y = 0, vx = 0, vy = 0, vz = 0;
      if (act) { x = P.x[i]; y = P.y[i]; vx = P.vx[i]; vy = P.vy[i]; vz = P.vz[i]; }
      if (MODE > 0 && act) {
        const double xold = x, yold = y;
        fields_and_kick<ORDER, false>(sE, sB, w, wstride, E, B, g, fq.k, x, y, vx, vy, vz);
        x = x + vx * fq.d2x;   // first half of the drift
        y = y + vy * fq.d2y;
        // more than half a cell in half a step: push_and_deposit.pyx:66-68
        if (fabs(x - xold) > 0.5 || fabs(y - yold) > 0.5) {
          if (MODE == 2) atomicOr(fq.ihole, SKB_CFL_BIT);
          else fq.ihole[0] = -1;
        }
      }
      double xs = x + q.offx, ys = y + q.offy;
      if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
      int ix, iy;
      double wx[NS], wy[NS];
      particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
      const bool changed = (a.iy != ACC_EMPTY) && (!act || iy != a.iy || ix != a.ix);
      if (__any_sync(SKB_FULL, changed)) warp_flush<NS>(a, sw, w, wstride, cur, g);
      if (act) {
        // particle velocity relative to the background shear, deposit.pxd:24
        const double vxr = vx + q.S * (y * g.dy + g.y0);
        a.ix = ix; a.iy = iy;
        accumulate<ORDER>(a, wx, wy, vxr, vy, vz);
        if (MODE == 2) {
          x = x + vx * fq.d2x;   // second half of the drift
          y = y + vy * fq.d2y;
          x = wrap_x(x, (double)g.nx);
          if (y < g.e0 || y >= g.e1) {      // calculate_ihole_cdef
            int slot = atomicAdd(fq.ihole, 1) & (SKB_CFL_BIT - 1);
            if (slot < fq.ntmax) fq.ihole[slot + 1] = (int)i + 1;
          }
          P.x[i] = x; P.y[i] = y; P.vx[i] = vx; P.vy[i] = vy; P.vz[i] = vz;
        }
      }
    }
    warp_flush<NS>(a, sw, w, wstride, cur, g);
    if (tile >= 0) {
      __syncthreads();
      flush_window(sw, w, wstride, cur, g);
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------
// Cell-aligned deposit (the fast path when the ordering is exact and cell_end is
// known): every warp is handed WHOLE cells, so all 32 lanes accumulate the same
// stencil in registers for the whole cell and there is exactly one warp reduction
// (reduce-scatter over lanes) and one emit per cell.  Loads are issued UNR iterations
// ahead of their use to keep enough bytes in flight with only 16 resident warps.
// A particle whose stencil base is not the cell it is filed under (caller passed a
// stale ordering) is deposited on its own through HBM atomics: correct, just slow.
// DET: deterministic variant — no atomics anywhere.  The NS*NS*4 stencil sums of every
// cell are written to cellsums[cell][NS*NS*4] (zeros for empty cells) and a second
// kernel (gather_cellsums_kernel) adds them into the source grid in a fixed order, so
// the result depends only on the (canonically ordered) particle array.
template <int ORDER, bool DET>
__global__ void __launch_bounds__(DEP_THREADS, 2)   // 2 CTAs/SM: at most 128 registers
deposit_cells_kernel(skb_particles_t P, double *__restrict__ cur, DevGrid g, DevTiling tl,
                     DepParams q, int parts, int wstride, int wrows,
                     double *__restrict__ cellsums) {
  constexpr int NS = ORDER + 1;
  constexpr int NV = NS * NS * 4;
  constexpr int UNR = DepUnroll<ORDER>::value;
  extern __shared__ double sw[];
  const int cells_log2 = tl.tlx + tl.tly;
  const int cpp = (1 << cells_log2) / parts;          // cells per CTA
  const int tile = blockIdx.x / parts;
  const int c0 = (tile << cells_log2) + (blockIdx.x % parts) * cpp;
  // dense layout: cell k = [cell_end[k-1], cell_end[k]); gapped layout: cell k =
  // [gap_start[k], gap_start[k] + gap_count[k])
  const bool gapped = tl.gap_start != nullptr;
  const int pbeg = gapped ? tl.gap_start[c0] : (c0 ? tl.cell_end[c0 - 1] : 0);
  const int pend = gapped ? tl.gap_start[c0 + cpp] : tl.cell_end[c0 + cpp - 1];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  if (pbeg == pend) {                                  // uniform: no particles here
    if (DET)
      for (int i = threadIdx.x; i < cpp * NV; i += DEP_THREADS) cellsums[(size_t)c0 * NV + i] = 0.0;
    return;
  }
  const Window w = tile_window(tile, tl, g);
  if (!DET) {
    zero_window(sw, wstride * wrows * 4);
    __syncthreads();
  }
  const int bx = (tile % tl.ntx) << tl.tlx, by = (tile / tl.ntx) << tl.tly;
  // each warp owns a contiguous block of cells (=> a contiguous particle range); the
  // cell boundaries are fetched 32 at a time, one per lane
  const int cpw = cpp / (DEP_THREADS / 32);           // cells per warp (>= 1)
  const int wc0 = c0 + wv * cpw;
  for (int cb = 0; cb < cpw; cb += 32) {
    const int mycell = wc0 + cb + lane;
    const bool mine = cb + lane < cpw;
    const int my_end = mine ? (gapped ? tl.gap_count[mycell] : tl.cell_end[mycell]) : 0;
    const int my_start = (mine && gapped) ? tl.gap_start[mycell] : 0;
    int prev_end = gapped ? 0 : ((wc0 + cb) ? tl.cell_end[wc0 + cb - 1] : 0);
    const int ncell = min(32, cpw - cb);
    for (int j = 0; j < ncell; j++) {
      int s = prev_end;
      int e = __shfl_sync(SKB_FULL, my_end, j);
      int next = e;                                    // where the following particles start
      if (gapped) {
        s = __shfl_sync(SKB_FULL, my_start, j);
        e += s;
        next = __shfl_sync(SKB_FULL, my_start, min(j + 1, 31));
        if (j + 1 >= ncell) next = pend;
      }
      prev_end = e;
      {
        // pull the particles that follow this cell (the warp's next cells) into L2
        // while this one is being reduced: 128-byte lines, one per lane and array
        const int ahead = next + lane * 16;
        if (ahead < min(next + 16 * DEP_PREFETCH_LINES, pend)) {
          dep_prefetch_l2(P.x + ahead); dep_prefetch_l2(P.y + ahead);
          dep_prefetch_l2(P.vx + ahead); dep_prefetch_l2(P.vy + ahead);
          dep_prefetch_l2(P.vz + ahead);
        }
      }
      if (s == e) {
        if (DET) for (int i = lane; i < NV; i += 32) cellsums[(size_t)(wc0 + cb + j) * NV + i] = 0.0;
        continue;
      }
      const int cell = wc0 + cb + j;
      const int local = cell & ((1 << cells_log2) - 1);
      Acc<NS> a;
#pragma unroll
      for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
      a.ix = bx + (local & ((1 << tl.tlx) - 1));
      a.iy = by + (local >> tl.tlx);
      for (int base = s; base < e; base += 32 * UNR) {
        double x[UNR], y[UNR], vx[UNR], vy[UNR], vz[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const int i = base + u * 32 + lane;
          if (i < e) { x[u] = P.x[i]; y[u] = P.y[i]; vx[u] = P.vx[i]; vy[u] = P.vy[i]; vz[u] = P.vz[i]; }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const int i = base + u * 32 + lane;
          if (i < e) {
            double xs = x[u] + q.offx, ys = y[u] + q.offy;
            if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
            int ix, iy;
            double wx[NS], wy[NS];
            particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
            // particle velocity relative to the background shear, deposit.pxd:24
            const double vxr = vx[u] + q.S * (y[u] * g.dy + g.y0);
            if (ix == a.ix && iy == a.iy) accumulate<ORDER>(a, wx, wy, vxr, vy[u], vz[u]);
            else single_particle_emit<NS>(wx, wy, ix, iy, vxr, vy[u], vz[u], cur, g);
          }
        }
      }
      // one reduce-scatter per cell; the NS*NS*4 sums land on as many lanes, which
      // add them to the window in parallel (or store them, deterministic variant)
      const int lo = (NS == 3) ? 1 : 0;
      const bool in_window = (a.ix - lo >= w.x0) && (a.ix - lo + NS <= w.x1) &&
                             (a.iy - lo >= w.y0) && (a.iy - lo + NS <= w.y1);
      double *cs = DET ? cellsums + (size_t)cell * NV : nullptr;
      if constexpr (NS == 2) {
        warp_reduce_scatter<16>(a.v, lane);
        if (lane < 16) {
          if (DET) cs[scatter_index<16>(lane)] = a.v[0];
          else emit_one<NS>(a.v[0], scatter_index<16>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
        }
      } else {
        warp_reduce_scatter<32>(a.v, lane);
        warp_reduce_scatter<4>(a.v + 32, lane);
        if (DET) {
          cs[scatter_index<32>(lane)] = a.v[0];
          if (lane < 4) cs[32 + scatter_index<4>(lane)] = a.v[32];
        } else {
          emit_one<NS>(a.v[0], scatter_index<32>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
          if (lane < 4)
            emit_one<NS>(a.v[32], 32 + scatter_index<4>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
        }
      }
    }
  }
  if (!DET) {
    __syncthreads();
    flush_window(sw, w, wstride, cur, g);
  }
}

// Cell-aligned deposit for FEW particles per cell (CIC): every half-warp takes a cell of
// its own, so one reduce-scatter (four exchange stages inside the half-warps) and one
// emit serve TWO cells and all 32 lanes end up holding one of the 2 x 16 sums.  At 64
// particles per cell the per-cell reduction + emit of deposit_cells_kernel is as much
// work as the accumulation itself; here it is halved.
__global__ void __launch_bounds__(DEP_THREADS, 2)
deposit_cells_half_kernel(skb_particles_t P, double *__restrict__ cur, DevGrid g, DevTiling tl,
                          DepParams q, int parts, int wstride, int wrows) {
  constexpr int NS = 2, UNR = 4;
  extern __shared__ double sw[];
  const int cells_log2 = tl.tlx + tl.tly;
  const int cpp = (1 << cells_log2) / parts;          // cells per CTA
  const int tile = blockIdx.x / parts;
  const int c0 = (tile << cells_log2) + (blockIdx.x % parts) * cpp;
  const bool gapped = tl.gap_start != nullptr;
  const int pbeg = gapped ? tl.gap_start[c0] : (c0 ? tl.cell_end[c0 - 1] : 0);
  const int pend = gapped ? tl.gap_start[c0 + cpp] : tl.cell_end[c0 + cpp - 1];
  if (pbeg == pend) return;
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int half = lane >> 4, hl = lane & 15;
  const Window w = tile_window(tile, tl, g);
  zero_window(sw, wstride * wrows * 4);
  __syncthreads();
  const int bx = (tile % tl.ntx) << tl.tlx, by = (tile / tl.ntx) << tl.tly;
  const int cpw = cpp / (DEP_THREADS / 32);           // cells per warp (>= 1)
  const int wc0 = c0 + wv * cpw;
  for (int j = 0; j < cpw; j += 2) {
    // my half's cell: [s, e)
    const int jc = j + half;
    int s = 0, e = 0;
    if (jc < cpw) {
      const int cell = wc0 + jc;
      if (gapped) { s = tl.gap_start[cell]; e = s + tl.gap_count[cell]; }
      else { s = cell ? tl.cell_end[cell - 1] : 0; e = tl.cell_end[cell]; }
    }
    const int n = e - s;
    const int nmax = max(n, __shfl_xor_sync(SKB_FULL, n, 16));
    if (nmax == 0) continue;
    const int local = (wc0 + min(jc, cpw - 1)) & ((1 << cells_log2) - 1);
    Acc<NS> a;
#pragma unroll
    for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
    a.ix = bx + (local & ((1 << tl.tlx) - 1));
    a.iy = by + (local >> tl.tlx);
    for (int base = 0; base < nmax; base += 16 * UNR) {
      double x[UNR], y[UNR], vx[UNR], vy[UNR], vz[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int k = base + u * 16 + hl;
        if (k < n) {
          const int i = s + k;
          x[u] = P.x[i]; y[u] = P.y[i]; vx[u] = P.vx[i]; vy[u] = P.vy[i]; vz[u] = P.vz[i];
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int k = base + u * 16 + hl;
        if (k < n) {
          const double xs = x[u] + q.offx, ys = y[u] + q.offy;
          int ix, iy;
          double wx[NS], wy[NS];
          particle_terms<1>(xs, ys, ix, iy, wx, wy);
          const double vxr = vx[u] + q.S * (y[u] * g.dy + g.y0);      // deposit.pxd:24
          if (ix == a.ix && iy == a.iy) accumulate<1>(a, wx, wy, vxr, vy[u], vz[u]);
          else single_particle_emit<NS>(wx, wy, ix, iy, vxr, vy[u], vz[u], cur, g);
        }
      }
    }
    const bool in_window = (a.ix >= w.x0) && (a.ix + NS <= w.x1) && (a.iy >= w.y0) &&
                           (a.iy + NS <= w.y1);
    warp_reduce_scatter<16, true>(a.v, lane);
    if (n > 0)
      emit_one<NS>(a.v[0], scatter_index<16>(hl), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
  }
  __syncthreads();
  flush_window(sw, w, wstride, cur, g);
}

// ---- cell-aligned deposit with a shared-memory particle ring ---------------------------
// deposit_cells_kernel holds a chunk of particle data in registers while its loads are in
// flight; with the 36 TSC accumulators that leaves 16 warps per SM and two chunks of
// loads per warp, and the kernel waits on memory (long scoreboard) at half the HBM rate.
// Here every warp streams its cells' particles through a ring of DEPR_NST stages of 64
// particles in shared memory, filled by 8-byte cp.async two stages ahead of the stage
// being accumulated (no registers held by loads in flight, every lane copies exactly the
// elements it will read, so completion is per-thread: cp.async.wait_group, no barrier).
// The stencil sums use fused multiply-adds (the deposit is compared at <= 1e-12, not bit
// for bit: summation order already differs from the reference).
#ifndef DEPR_NST
#define DEPR_NST 3
#endif
#define DEPR_STAGE 64
#define DEPR_STAGE_D (5 * DEPR_STAGE)

__device__ __forceinline__ void depr_cp8(double *smem_dst, const double *gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

template <int ORDER>
__device__ __forceinline__ void accumulate_fma(Acc<ORDER + 1> &a, const double (&wx)[ORDER + 1],
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
      v[1] = fma(wgt, vxr, v[1]);
      v[2] = fma(wgt, vy, v[2]);
      v[3] = fma(wgt, vz, v[3]);
    }
}

// position in the warp's chunk sequence: cell j (0..31, 32 = end), offset inside it
struct DeprIt { int j, off; };

template <int ORDER>
__global__ void __launch_bounds__(DEP_THREADS, 2)
deposit_cells_ring_kernel(skb_particles_t P, double *__restrict__ cur, DevGrid g, DevTiling tl,
                          DepParams q, int parts, int wstride, int wrows) {
  constexpr int NS = ORDER + 1;
  extern __shared__ double sw[];
  const int cells_log2 = tl.tlx + tl.tly;
  const int cpp = (1 << cells_log2) / parts;          // cells per CTA
  const int tile = blockIdx.x / parts;
  const int c0 = (tile << cells_log2) + (blockIdx.x % parts) * cpp;
  const bool gapped = tl.gap_start != nullptr;
  const int pbeg = gapped ? tl.gap_start[c0] : (c0 ? tl.cell_end[c0 - 1] : 0);
  const int pend = gapped ? tl.gap_start[c0 + cpp] : tl.cell_end[c0 + cpp - 1];
  if (pbeg == pend) return;                            // uniform: no particles here
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const Window w = tile_window(tile, tl, g);
  const int wsize = wstride * wrows * 4;
  zero_window(sw, wsize);
  __syncthreads();
  double *const ring = sw + ((wsize + 1) & ~1) + wv * (DEPR_NST * DEPR_STAGE_D);
  const int bx = (tile % tl.ntx) << tl.tlx, by = (tile / tl.ntx) << tl.tly;
  // the warp's cells (at most 32, one per lane): first slot and particle count
  const int cpw = cpp / (DEP_THREADS / 32);
  const int wc0 = c0 + wv * cpw;
  int my_start = 0, my_n = 0;
  if (lane < cpw) {
    const int cell = wc0 + lane;
    if (gapped) { my_start = tl.gap_start[cell]; my_n = tl.gap_count[cell]; }
    else { my_start = cell ? tl.cell_end[cell - 1] : 0; my_n = tl.cell_end[cell] - my_start; }
  }
  const unsigned nonempty = __ballot_sync(SKB_FULL, my_n > 0);
  auto after = [&](int j) {                            // next non-empty cell after j
    const unsigned m = j >= 31 ? 0u : (nonempty & ~((2u << j) - 1u));
    return m ? __ffs(m) - 1 : 32;
  };
  auto advance = [&](DeprIt &it) {
    if (it.j >= 32) return;
    it.off += DEPR_STAGE;
    if (it.off >= __shfl_sync(SKB_FULL, my_n, it.j)) { it.off = 0; it.j = after(it.j); }
  };
  auto fetch = [&](int stage, const DeprIt &it) {
    if (it.j < 32) {
      const int s = __shfl_sync(SKB_FULL, my_start, it.j) + it.off;
      const int n = __shfl_sync(SKB_FULL, my_n, it.j) - it.off;
      double *d = ring + stage * DEPR_STAGE_D;
#pragma unroll
      for (int u = 0; u < DEPR_STAGE / 32; u++) {
        const int k = u * 32 + lane;
        if (k < n) {
          depr_cp8(d + k, P.x + s + k);
          depr_cp8(d + DEPR_STAGE + k, P.y + s + k);
          depr_cp8(d + 2 * DEPR_STAGE + k, P.vx + s + k);
          depr_cp8(d + 3 * DEPR_STAGE + k, P.vy + s + k);
          depr_cp8(d + 4 * DEPR_STAGE + k, P.vz + s + k);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // (one group per call, even empty)
  };
  DeprIt f, c;
  f.j = c.j = nonempty ? __ffs(nonempty) - 1 : 32;
  f.off = c.off = 0;
#pragma unroll
  for (int s = 0; s < DEPR_NST - 1; s++) { fetch(s, f); advance(f); }
  int stage = 0;
  Acc<NS> a;
#pragma unroll
  for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
  while (c.j < 32) {
    {
      int fs = stage + DEPR_NST - 1;
      if (fs >= DEPR_NST) fs -= DEPR_NST;
      fetch(fs, f);                                    // the stage consumed one turn ago
      advance(f);
    }
    const int ncell = __shfl_sync(SKB_FULL, my_n, c.j);
    const int n = ncell - c.off;
    if (c.off == 0) {
      const int local = (wc0 + c.j) & ((1 << cells_log2) - 1);
      a.ix = bx + (local & ((1 << tl.tlx) - 1));
      a.iy = by + (local >> tl.tlx);
      // pull the particles of the cells that follow into L2, beyond the reach of the ring
      const int nj = after(c.j);
      const int next = nj < 32 ? __shfl_sync(SKB_FULL, my_start, nj & 31) : pend;
      const int ahead = next + lane * 16;
      if (ahead < min(next + 16 * DEP_PREFETCH_LINES, pend)) {
        dep_prefetch_l2(P.x + ahead); dep_prefetch_l2(P.y + ahead);
        dep_prefetch_l2(P.vx + ahead); dep_prefetch_l2(P.vy + ahead);
        dep_prefetch_l2(P.vz + ahead);
      }
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPR_NST - 1) : "memory");
    const double *d = ring + stage * DEPR_STAGE_D;
#pragma unroll
    for (int u = 0; u < DEPR_STAGE / 32; u++) {
      const int k = u * 32 + lane;
      if (k < n) {
        const double x = d[k], y = d[DEPR_STAGE + k], vx = d[2 * DEPR_STAGE + k];
        const double vy = d[3 * DEPR_STAGE + k], vz = d[4 * DEPR_STAGE + k];
        double xs = x + q.offx, ys = y + q.offy;
        if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
        int ix, iy;
        double wx[NS], wy[NS];
        particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
        // particle velocity relative to the background shear, deposit.pxd:24
        const double vxr = vx + q.S * (y * g.dy + g.y0);
        if (ix == a.ix && iy == a.iy) accumulate_fma<ORDER>(a, wx, wy, vxr, vy, vz);
        else single_particle_emit<NS>(wx, wy, ix, iy, vxr, vy, vz, cur, g);
      }
    }
    if (n <= DEPR_STAGE) {
      // last chunk of the cell: one reduce-scatter, the NS*NS*4 sums land on as many lanes
      const int lo = (NS == 3) ? 1 : 0;
      const bool in_window = (a.ix - lo >= w.x0) && (a.ix - lo + NS <= w.x1) &&
                             (a.iy - lo >= w.y0) && (a.iy - lo + NS <= w.y1);
      if constexpr (NS == 2) {
        warp_reduce_scatter<16>(a.v, lane);
        if (lane < 16)
          emit_one<NS>(a.v[0], scatter_index<16>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
      } else {
        warp_reduce_scatter<32>(a.v, lane);
        warp_reduce_scatter<4>(a.v + 32, lane);
        emit_one<NS>(a.v[0], scatter_index<32>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
        if (lane < 4)
          emit_one<NS>(a.v[32], 32 + scatter_index<4>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
      }
#pragma unroll
      for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
    }
    advance(c);
    stage = stage + 1 == DEPR_NST ? 0 : stage + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  flush_window(sw, w, wstride, cur, g);
}

// ---- two cells per warp + particle ring --------------------------------------------------
// deposit_cells_ring_kernel with every half-warp on a cell of its own (cells 2i and 2i+1 of
// the warp's block advance together): one reduce-scatter over 16 lanes and one emit serve
// TWO cells, which halves the per-cell reduction work — a quarter of the TSC kernel's
// instructions at 256 particles per cell, and as much as the accumulation itself below
// ~100 particles per cell.  Ring stage = 32 particles per half-warp.
#define DEPP_CHUNK 32
#define DEPP_STAGE_D (2 * 5 * DEPP_CHUNK)

// Reduce-scatter of V values over the 16 lanes of each half-warp.  V = 16: v[0] of lane hl
// = value half_scatter_index<16>(hl); V = 32: v[0], v[1] = values half_scatter_index<32>(hl)
// + {0, 1}; V = 4: v[0] = value half_scatter_index<4>(hl) (replicated over hl >> 2).
template <int V>
__device__ __forceinline__ int half_scatter_index(int hl) {
  int idx = 0;
#pragma unroll
  for (int s = 0, h = V / 2; h >= 1 && s < 4; s++, h >>= 1) idx += ((hl >> s) & 1) * h;
  return idx;
}
template <int V>
__device__ __forceinline__ void half_reduce_scatter(double *v, int lane) {
  int bit = 1;
#pragma unroll
  for (int h = V / 2; h >= 1 && bit < 16; h >>= 1, bit <<= 1) {
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
  for (; bit < 16; bit <<= 1) v[0] += __shfl_xor_sync(SKB_FULL, v[0], bit);
}

struct DeppIt { int i, c; };      // pair of cells (0..15, 16 = end), chunk inside it

template <int ORDER>
__global__ void __launch_bounds__(DEP_THREADS, 2)
deposit_cells_pair_kernel(skb_particles_t P, double *__restrict__ cur, DevGrid g, DevTiling tl,
                          DepParams q, int parts, int wstride, int wrows) {
  constexpr int NS = ORDER + 1;
  extern __shared__ double sw[];
  const int cells_log2 = tl.tlx + tl.tly;
  const int cpp = (1 << cells_log2) / parts;          // cells per CTA
  const int tile = blockIdx.x / parts;
  const int c0 = (tile << cells_log2) + (blockIdx.x % parts) * cpp;
  const bool gapped = tl.gap_start != nullptr;
  const int pbeg = gapped ? tl.gap_start[c0] : (c0 ? tl.cell_end[c0 - 1] : 0);
  const int pend = gapped ? tl.gap_start[c0 + cpp] : tl.cell_end[c0 + cpp - 1];
  if (pbeg == pend) return;                            // uniform: no particles here
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int half = lane >> 4, hl = lane & 15;
  const Window w = tile_window(tile, tl, g);
  const int wsize = wstride * wrows * 4;
  zero_window(sw, wsize);
  __syncthreads();
  double *const ring = sw + ((wsize + 1) & ~1) + wv * (DEPR_NST * DEPP_STAGE_D) +
                       half * (5 * DEPP_CHUNK);
  const int bx = (tile % tl.ntx) << tl.tlx, by = (tile / tl.ntx) << tl.tly;
  // the warp's cells (at most 32, one per lane): first slot and particle count
  const int cpw = cpp / (DEP_THREADS / 32);
  const int wc0 = c0 + wv * cpw;
  int my_start = 0, my_n = 0;
  if (lane < cpw) {
    const int cell = wc0 + lane;
    if (gapped) { my_start = tl.gap_start[cell]; my_n = tl.gap_count[cell]; }
    else { my_start = cell ? tl.cell_end[cell - 1] : 0; my_n = tl.cell_end[cell] - my_start; }
  }
  // lane i < 16: chunks of pair i = those of its bigger cell
  const int na = __shfl_sync(SKB_FULL, my_n, (2 * lane) & 31);
  const int nb = __shfl_sync(SKB_FULL, my_n, (2 * lane + 1) & 31);
  const int my_nch = lane < 16 ? (max(na, nb) + DEPP_CHUNK - 1) / DEPP_CHUNK : 0;
  const unsigned nonempty = __ballot_sync(SKB_FULL, my_nch > 0);
  auto after = [&](int i) {                            // next non-empty pair after i
    const unsigned m = nonempty & ~((2u << i) - 1u);
    return m ? __ffs(m) - 1 : 16;
  };
  auto advance = [&](DeppIt &it) {
    if (it.i >= 16) return;
    it.c += 1;
    if (it.c >= __shfl_sync(SKB_FULL, my_nch, it.i)) { it.c = 0; it.i = after(it.i); }
  };
  auto fetch = [&](int stage, const DeppIt &it) {
    if (it.i < 16) {
      const int cell = 2 * it.i + half;
      const int s = __shfl_sync(SKB_FULL, my_start, cell) + it.c * DEPP_CHUNK;
      const int n = __shfl_sync(SKB_FULL, my_n, cell) - it.c * DEPP_CHUNK;
      double *d = ring + stage * DEPP_STAGE_D;
#pragma unroll
      for (int u = 0; u < DEPP_CHUNK / 16; u++) {
        const int k = u * 16 + hl;
        if (k < n) {
          depr_cp8(d + k, P.x + s + k);
          depr_cp8(d + DEPP_CHUNK + k, P.y + s + k);
          depr_cp8(d + 2 * DEPP_CHUNK + k, P.vx + s + k);
          depr_cp8(d + 3 * DEPP_CHUNK + k, P.vy + s + k);
          depr_cp8(d + 4 * DEPP_CHUNK + k, P.vz + s + k);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // (one group per call, even empty)
  };
  DeppIt f, c;
  f.i = c.i = nonempty ? __ffs(nonempty) - 1 : 16;
  f.c = c.c = 0;
#pragma unroll
  for (int s = 0; s < DEPR_NST - 1; s++) { fetch(s, f); advance(f); }
  int stage = 0;
  Acc<NS> a;
#pragma unroll
  for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
  a.ix = a.iy = 0;
  int ncell = 0;
  while (c.i < 16) {
    {
      int fs = stage + DEPR_NST - 1;
      if (fs >= DEPR_NST) fs -= DEPR_NST;
      fetch(fs, f);                                    // the stage consumed one turn ago
      advance(f);
    }
    const int nch = __shfl_sync(SKB_FULL, my_nch, c.i);
    if (c.c == 0) {
      const int cl = 2 * c.i + half;                   // my half's cell of this pair
      ncell = __shfl_sync(SKB_FULL, my_n, cl);
      const int local = (wc0 + min(cl, cpw - 1)) & ((1 << cells_log2) - 1);
      a.ix = bx + (local & ((1 << tl.tlx) - 1));
      a.iy = by + (local >> tl.tlx);
      // pull my half's next cell into L2, beyond the reach of the ring
      const int nx2 = min(cl + 2, 31);
      const int s2 = __shfl_sync(SKB_FULL, my_start, nx2);
      const int n2 = __shfl_sync(SKB_FULL, my_n, nx2);
      if (cl + 2 < cpw && hl * 16 < n2) {
        const int ahead = s2 + hl * 16;
        dep_prefetch_l2(P.x + ahead); dep_prefetch_l2(P.y + ahead);
        dep_prefetch_l2(P.vx + ahead); dep_prefetch_l2(P.vy + ahead);
        dep_prefetch_l2(P.vz + ahead);
      }
    }
    const int n = ncell - c.c * DEPP_CHUNK;
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPR_NST - 1) : "memory");
    const double *d = ring + stage * DEPP_STAGE_D;
#pragma unroll
    for (int u = 0; u < DEPP_CHUNK / 16; u++) {
      const int k = u * 16 + hl;
      if (k < n) {
        const double x = d[k], y = d[DEPP_CHUNK + k], vx = d[2 * DEPP_CHUNK + k];
        const double vy = d[3 * DEPP_CHUNK + k], vz = d[4 * DEPP_CHUNK + k];
        double xs = x + q.offx, ys = y + q.offy;
        if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
        int ix, iy;
        double wx[NS], wy[NS];
        particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
        // particle velocity relative to the background shear, deposit.pxd:24
        const double vxr = vx + q.S * (y * g.dy + g.y0);
        if (ix == a.ix && iy == a.iy) accumulate_fma<ORDER>(a, wx, wy, vxr, vy, vz);
        else single_particle_emit<NS>(wx, wy, ix, iy, vxr, vy, vz, cur, g);
      }
    }
    if (c.c + 1 >= nch) {
      // last chunk of the pair: one reduce-scatter for both cells
      const int lo = (NS == 3) ? 1 : 0;
      const bool in_window = (a.ix - lo >= w.x0) && (a.ix - lo + NS <= w.x1) &&
                             (a.iy - lo >= w.y0) && (a.iy - lo + NS <= w.y1);
      if constexpr (NS == 2) {
        half_reduce_scatter<16>(a.v, lane);
        if (ncell > 0)
          emit_one<NS>(a.v[0], half_scatter_index<16>(hl), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
      } else {
        half_reduce_scatter<32>(a.v, lane);
        half_reduce_scatter<4>(a.v + 32, lane);
        if (ncell > 0) {
          const int i0 = half_scatter_index<32>(hl);
          emit_one<NS>(a.v[0], i0, in_window, a.ix, a.iy, sw, w, wstride, cur, g);
          emit_one<NS>(a.v[1], i0 + 1, in_window, a.ix, a.iy, sw, w, wstride, cur, g);
          if (hl < 4)
            emit_one<NS>(a.v[32], 32 + half_scatter_index<4>(hl), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
        }
      }
#pragma unroll
      for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
    }
    advance(c);
    stage = stage + 1 == DEPR_NST ? 0 : stage + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  flush_window(sw, w, wstride, cur, g);
}

// Deterministic second phase: every grid cell adds the contributions of the (up to
// NS*NS) stencil-base cells that reach it, in a fixed order.  One thread per (cell, k).
template <int NS>
__global__ void __launch_bounds__(256)
gather_cellsums_kernel(const double *__restrict__ cellsums, double *__restrict__ cur,
                       DevGrid g, KeyParams kp) {
  constexpr int NV = NS * NS * 4;
  const int lo = (NS == 3) ? 1 : 0;
  long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)g.mx * g.myp * 4) return;
  const int k = (int)(idx & 3);
  const long long cellidx = idx >> 2;
  const int gy = (int)(cellidx / g.mx), gx = (int)(cellidx - (long long)gy * g.mx);
  double sum = 0.0;
#pragma unroll
  for (int r = 0; r < NS; r++)
#pragma unroll
    for (int c = 0; c < NS; c++) {
      const int bx = gx + lo - c, by = gy + lo - r;     // stencil base that reaches (gx, gy)
      if (bx < 0 || bx >= g.mx || by < 0 || by >= g.myp) continue;
      const int mxm = (1 << kp.tlx) - 1, mym = (1 << kp.tly) - 1;
      const int key = ((((by >> kp.tly) * kp.ntx + (bx >> kp.tlx)) << (kp.tlx + kp.tly)) |
                       ((by & mym) << kp.tlx) | (bx & mxm));
      sum += cellsums[(size_t)key * NV + (r * NS + c) * 4 + k];
    }
  cur[idx] += sum;
}

// ---------------------------------------------------------------------------------
// Cell-aligned fused push_and_deposit (push_and_deposit.pyx:10-170) for exactly
// ordered input: as deposit_cells_kernel, plus the E,B windows; every particle is
// gathered at its OLD position (the one it is filed under), kicked, half-drifted and
// deposited.  Particles whose half-step stencil base is still their cell (the vast
// majority at CFL-limited steps) accumulate in registers; the others are deposited on
// their own (shared-memory window or HBM atomics).  MODE 2 (update): second half
// drift, x wrap, leaver list, write-back, and the histogram of the new cell keys for
// the following tile sort (into next_counts, a different array than cell_end).
struct PdCellsParams {
  KeyParams key;
  int *next_counts;   // MODE 2: histogram of the new keys (may be NULL)
};

template <int ORDER, int MODE>
__global__ void __launch_bounds__(DEP_THREADS)
pd_cells_kernel(skb_particles_t P, const double *__restrict__ E,
                const double *__restrict__ B, double *__restrict__ cur, DevGrid g,
                DevTiling tl, DepParams q, FusedParams fq, PdCellsParams pc, int parts,
                int wstride, int wrows) {
  constexpr int NS = ORDER + 1;
  constexpr int UNR = 2;
  extern __shared__ double smem[];
  double *sw = smem;
  double *sE = smem + wstride * wrows * 4;
  double *sB = sE + wstride * wrows * 3;
  const int cells_log2 = tl.tlx + tl.tly;
  const int cpp = (1 << cells_log2) / parts;
  const int tile = blockIdx.x / parts;
  const int c0 = (tile << cells_log2) + (blockIdx.x % parts) * cpp;
  const int pbeg = c0 ? tl.cell_end[c0 - 1] : 0;
  const int pend = tl.cell_end[c0 + cpp - 1];
  if (pbeg == pend) return;
  const Window w = tile_window(tile, tl, g);
  zero_window(sw, wstride * wrows * 4);
  stage_window(sE, E, w, wstride, g);
  stage_window(sB, B, w, wstride, g);
  __syncthreads();
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int bx = (tile % tl.ntx) << tl.tlx, by = (tile / tl.ntx) << tl.tly;
  const int cpw = cpp / (DEP_THREADS / 32);
  const int wc0 = c0 + wv * cpw;
  for (int cb = 0; cb < cpw; cb += 32) {
    const int my_end = (cb + lane < cpw) ? tl.cell_end[wc0 + cb + lane] : 0;
    int prev_end = (wc0 + cb) ? tl.cell_end[wc0 + cb - 1] : 0;
    const int ncell = min(32, cpw - cb);
    for (int j = 0; j < ncell; j++) {
      const int s = prev_end;
      const int e = __shfl_sync(SKB_FULL, my_end, j);
      prev_end = e;
      {
        // pull the particles that follow this cell into L2 (see deposit_cells_kernel)
        const int ahead = e + lane * 16;
        if (ahead < min(e + 16 * DEP_PREFETCH_LINES, pend)) {
          dep_prefetch_l2(P.x + ahead); dep_prefetch_l2(P.y + ahead);
          dep_prefetch_l2(P.vx + ahead); dep_prefetch_l2(P.vy + ahead);
          dep_prefetch_l2(P.vz + ahead);
        }
      }
      if (s == e) continue;
      const int local = (wc0 + cb + j) & ((1 << cells_log2) - 1);
      Acc<NS> a;
#pragma unroll
      for (int i = 0; i < NS * NS * 4; i++) a.v[i] = 0.0;
      a.ix = bx + (local & ((1 << tl.tlx) - 1));
      a.iy = by + (local >> tl.tlx);
      for (int base = s; base < e; base += 32 * UNR) {
        double x[UNR], y[UNR], vx[UNR], vy[UNR], vz[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const int i = base + u * 32 + lane;
          if (i < e) { x[u] = P.x[i]; y[u] = P.y[i]; vx[u] = P.vx[i]; vy[u] = P.vy[i]; vz[u] = P.vz[i]; }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const int i = base + u * 32 + lane;
          int key = -1;
          if (i < e) {
            const double xold = x[u], yold = y[u];
            fields_and_kick<ORDER, false>(sE, sB, w, wstride, E, B, g, fq.k, x[u], y[u],
                                          vx[u], vy[u], vz[u]);
            x[u] = x[u] + vx[u] * fq.d2x;   // first half of the drift
            y[u] = y[u] + vy[u] * fq.d2y;
            // more than half a cell in half a step: push_and_deposit.pyx:66-68
            if (fabs(x[u] - xold) > 0.5 || fabs(y[u] - yold) > 0.5) {
              if (MODE == 2) atomicOr(fq.ihole, SKB_CFL_BIT);
              else fq.ihole[0] = -1;
            }
            double xs = x[u] + q.offx, ys = y[u] + q.offy;
            if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
            int ix, iy;
            double wx[NS], wy[NS];
            particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
            const double vxr = vx[u] + q.S * (y[u] * g.dy + g.y0);
            if (ix == a.ix && iy == a.iy) accumulate<ORDER>(a, wx, wy, vxr, vy[u], vz[u]);
            else stray_particle_emit<NS>(wx, wy, ix, iy, vxr, vy[u], vz[u], sw, w, wstride, cur, g);
            if (MODE == 2) {
              x[u] = x[u] + vx[u] * fq.d2x;   // second half of the drift
              y[u] = y[u] + vy[u] * fq.d2y;
              x[u] = wrap_x(x[u], (double)g.nx);
              if (y[u] < g.e0 || y[u] >= g.e1) {      // calculate_ihole_cdef
                int slot = atomicAdd(fq.ihole, 1) & (SKB_CFL_BIT - 1);
                if (slot < fq.ntmax) fq.ihole[slot + 1] = i + 1;
              } else if (pc.next_counts) {
                key = cell_key(x[u], y[u], pc.key);
              }
              P.x[i] = x[u]; P.y[i] = y[u]; P.vx[i] = vx[u]; P.vy[i] = vy[u]; P.vz[i] = vz[u];
            }
          }
          if (MODE == 2 && pc.next_counts) {
            const unsigned peers = __match_any_sync(SKB_FULL, key);
            if (key >= 0 && lane == __ffs(peers) - 1)
              atomicAdd(pc.next_counts + key, __popc(peers));
          }
        }
      }
      const int lo = (NS == 3) ? 1 : 0;
      const bool in_window = (a.ix - lo >= w.x0) && (a.ix - lo + NS <= w.x1) &&
                             (a.iy - lo >= w.y0) && (a.iy - lo + NS <= w.y1);
      if constexpr (NS == 2) {
        warp_reduce_scatter<16>(a.v, lane);
        if (lane < 16)
          emit_one<NS>(a.v[0], scatter_index<16>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
      } else {
        warp_reduce_scatter<32>(a.v, lane);
        warp_reduce_scatter<4>(a.v + 32, lane);
        emit_one<NS>(a.v[0], scatter_index<32>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
        if (lane < 4)
          emit_one<NS>(a.v[32], 32 + scatter_index<4>(lane), in_window, a.ix, a.iy, sw, w, wstride, cur, g);
      }
    }
  }
  __syncthreads();
  flush_window(sw, w, wstride, cur, g);
}

template <int ORDER, int MODE>
static int launch_pd_cells(skb_particles_t p, const double *E, const double *B,
                           double *current, const DevGrid &g, const DevTiling &tl,
                           const DepParams &q, const FusedParams &fq, const PdCellsParams &pc,
                           cudaStream_t st) {
  const int ntiles = tl.ntx * tl.nty;
  const int cells = 1 << (tl.tlx + tl.tly);
  int parts = 1;
  while (parts < cells / 8 && (long long)ntiles * parts < 8 * 148) parts <<= 1;
  const int ws = window_stride(tl), wr = window_rows(tl);
  const size_t smem = (size_t)ws * wr * 10 * sizeof(double);
  auto k = pd_cells_kernel<ORDER, MODE>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  k<<<ntiles * parts, DEP_THREADS, smem, st>>>(p, E, B, current, g, tl, q, fq, pc, parts, ws, wr);
  SKB_CHECK_LAUNCH();
  return 0;
}

// decode ihole[0] = count | CFL bit into the reference's in-band convention
__global__ void finalize_fused_ihole_kernel(int *ihole, int ntmax) {
  int v = ihole[0];
  int n = v & (SKB_CFL_BIT - 1);
  if (v & SKB_CFL_BIT) ihole[0] = -1;
  else if (n > ntmax) ihole[0] = -(n - 1);
  else ihole[0] = n;
}

template <int ORDER, int MODE>
static int launch_deposit(skb_particles_t p, long long np, const double *E,
                          const double *B, double *current, const DevGrid &g,
                          const DevTiling &tl, const DepParams &q,
                          const FusedParams &fq, cudaStream_t st) {
  // CTA work item: up to 8 chunks, fewer when that would leave SMs idle
  int mult = 8;
  while (mult > 1 && (np / ((long long)tl.chunk * mult)) < 4 * 148) mult >>= 1;
  const int span = tl.chunk * mult;
  const int ws = window_stride(tl), wr = window_rows(tl);
  size_t smem = (size_t)ws * wr * (MODE > 0 ? 10 : 4) * sizeof(double);
  long long nblk = (np + span - 1) / span;
  auto k = deposit_kernel<ORDER, MODE>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  k<<<(unsigned)nblk, DEP_THREADS, smem, st>>>(p, np, E, B, current, g, tl, q, fq, span, ws, wr);
  SKB_CHECK_LAUNCH();
  return 0;
}

static int deposit_impl(skb_particles_t p, long long np, double *current,
                        const skb_grid_t *grid, int order, double S,
                        const skb_tiling_t *tiling, double *cellsums, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (np <= 0) return 0;
  DevGrid g = make_grid(grid);
  DevTiling tl = make_tiling(tiling);
  DepParams q;
  q.offx = g.lbx - 0.5;              // deposit.pyx:14-15
  q.offy = g.lby - 0.5 - g.noff;
  q.S = S;
  FusedParams fq = {};
  if (order != 1 && order != 2) return (int)cudaErrorInvalidValue;
  if ((tl.gap_start && tl.gap_count) || (tl.tile_offsets && tl.cell_end && tl.n_sorted > 0)) {
    // exact ordering with per-cell ranges (dense or gapped): whole cells per warp
    const int ntiles = tl.ntx * tl.nty;
    const int cells = 1 << (tl.tlx + tl.tly);
    int parts = 1;
    while (parts < cells / 8 && (long long)ntiles * parts < 8 * 148) parts <<= 1;
    const int ws = window_stride(tl), wr = window_rows(tl);
    const size_t smem = (size_t)ws * wr * 4 * sizeof(double);
    if (cellsums) {
      if (tl.gap_start || np != tl.n_sorted) return (int)cudaErrorInvalidValue;  // needs a full exact dense order
      if (order == 1)
        deposit_cells_kernel<1, true><<<ntiles * parts, DEP_THREADS, 0, st>>>(p, current, g, tl, q, parts, ws, wr, cellsums);
      else
        deposit_cells_kernel<2, true><<<ntiles * parts, DEP_THREADS, 0, st>>>(p, current, g, tl, q, parts, ws, wr, cellsums);
      SKB_CHECK_LAUNCH();
      KeyParams kp = make_keyparams(g, order, tl.tlx, tl.tly);
      const unsigned gb = (unsigned)(((long long)g.mx * g.myp * 4 + 255) / 256);
      if (order == 1) gather_cellsums_kernel<2><<<gb, 256, 0, st>>>(cellsums, current, g, kp);
      else gather_cellsums_kernel<3><<<gb, 256, 0, st>>>(cellsums, current, g, kp);
      SKB_CHECK_LAUNCH();
      return 0;
    }
    // few particles per cell (CIC): two cells per warp (SKB_DEP_HALF=0/1 overrides)
    static const int half_env = getenv("SKB_DEP_HALF") ? atoi(getenv("SKB_DEP_HALF")) : -1;
    const double ppc = (double)np / ((double)g.nx * (double)g.nyp);
    const bool half = half_env >= 0 ? half_env != 0 : ppc < 100.0;
    // particle ring in shared memory (SKB_DEP_RING=0 falls back to register staging); the
    // two-cells-per-warp kernel keeps the low particle counts
    static const int ring_env = getenv("SKB_DEP_RING") ? atoi(getenv("SKB_DEP_RING")) : -1;
    const bool ring = cells / parts <= 32 * (DEP_THREADS / 32) && ring_env != 0;
    // two cells per warp: TSC and few particles per cell (SKB_DEP_PAIR=0/1 overrides)
    static const int pair_env = getenv("SKB_DEP_PAIR") ? atoi(getenv("SKB_DEP_PAIR")) : -1;
    const bool pair = ring && (pair_env >= 0 ? pair_env != 0 : (order == 2 || half));
    if (pair) {
      const size_t rsmem = ((size_t)((ws * wr * 4 + 1) & ~1) +
                            (size_t)(DEP_THREADS / 32) * DEPR_NST * DEPP_STAGE_D) * sizeof(double);
      auto k = order == 1 ? deposit_cells_pair_kernel<1> : deposit_cells_pair_kernel<2>;
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
      if (e != cudaSuccess) return (int)e;
      k<<<ntiles * parts, DEP_THREADS, rsmem, st>>>(p, current, g, tl, q, parts, ws, wr);
    } else if (ring && !(order == 1 && half)) {
      const size_t rsmem = ((size_t)((ws * wr * 4 + 1) & ~1) +
                            (size_t)(DEP_THREADS / 32) * DEPR_NST * DEPR_STAGE_D) * sizeof(double);
      auto k = order == 1 ? deposit_cells_ring_kernel<1> : deposit_cells_ring_kernel<2>;
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
      if (e != cudaSuccess) return (int)e;
      k<<<ntiles * parts, DEP_THREADS, rsmem, st>>>(p, current, g, tl, q, parts, ws, wr);
    } else if (order == 1 && half)
      deposit_cells_half_kernel<<<ntiles * parts, DEP_THREADS, smem, st>>>(p, current, g, tl, q, parts, ws, wr);
    else if (order == 1)
      deposit_cells_kernel<1, false><<<ntiles * parts, DEP_THREADS, smem, st>>>(p, current, g, tl, q, parts, ws, wr, nullptr);
    else
      deposit_cells_kernel<2, false><<<ntiles * parts, DEP_THREADS, smem, st>>>(p, current, g, tl, q, parts, ws, wr, nullptr);
    SKB_CHECK_LAUNCH();
    if (tl.gap_start || np <= tl.n_sorted) return 0;
    // unsorted tail [n_sorted, np): generic kernel without ordering
    const long long n0 = tl.n_sorted;
    p.x += n0; p.y += n0; p.vx += n0; p.vy += n0; p.vz += n0;
    np -= n0;
    tl = make_tiling(nullptr);
  } else if (cellsums) {
    return (int)cudaErrorInvalidValue;
  }
  if (order == 1) return launch_deposit<1, 0>(p, np, nullptr, nullptr, current, g, tl, q, fq, st);
  return launch_deposit<2, 0>(p, np, nullptr, nullptr, current, g, tl, q, fq, st);
}

extern "C" int skb_deposit(skb_particles_t p, long long np, double *current,
                           const skb_grid_t *grid, int order, double S,
                           const skb_tiling_t *tiling, void *stream) {
  return deposit_impl(p, np, current, grid, order, S, tiling, nullptr, stream);
}

extern "C" int skb_deposit_deterministic(skb_particles_t p, long long np, double *current,
                                         const skb_grid_t *grid, int order, double S,
                                         const skb_tiling_t *tiling, double *cellsums,
                                         void *stream) {
  if (!cellsums) return (int)cudaErrorInvalidValue;
  return deposit_impl(p, np, current, grid, order, S, tiling, cellsums, stream);
}

extern "C" int skb_push_and_deposit(skb_particles_t p, long long np, const double *E,
                                    const double *B, const skb_grid_t *grid, int order,
                                    double qtmh, double dt, int *ihole, int ntmax,
                                    double *current, double S, int update,
                                    const skb_tiling_t *tiling, int *next_cell_counts,
                                    int key_tlx, int key_tly, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DevGrid g = make_grid(grid);
  DevTiling tl = make_tiling(tiling);
  FusedParams fq;
  fq.k = make_kick(g, qtmh, dt, 0.0, 0.0);
  fq.d2x = 0.5 * dt / g.dx;
  fq.d2y = 0.5 * dt / g.dy;
  fq.ihole = ihole;
  fq.ntmax = ntmax;
  DepParams q;
  q.offx = fq.k.offEx;               // offsetE reused, push_and_deposit.pyx:71
  q.offy = fq.k.offEy;
  q.S = S;
  if (order != 1 && order != 2) return (int)cudaErrorInvalidValue;
  if (update) {
    cudaError_t e = cudaMemsetAsync(ihole, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  // whole-cells-per-warp path: exact ordering with per-cell ranges covering ALL particles
  const bool cells_path = tl.tile_offsets && tl.cell_end && tl.n_sorted == np && np > 0 &&
                          (!next_cell_counts || next_cell_counts != tl.cell_end);
  if (cells_path) {
    PdCellsParams pc;
    pc.key = make_keyparams(g, order, key_tlx, key_tly);
    pc.next_counts = update ? next_cell_counts : nullptr;
    int rc;
    if (order == 1)
      rc = update ? launch_pd_cells<1, 2>(p, E, B, current, g, tl, q, fq, pc, st)
                  : launch_pd_cells<1, 1>(p, E, B, current, g, tl, q, fq, pc, st);
    else
      rc = update ? launch_pd_cells<2, 2>(p, E, B, current, g, tl, q, fq, pc, st)
                  : launch_pd_cells<2, 1>(p, E, B, current, g, tl, q, fq, pc, st);
    if (rc) return rc;
  } else if (np > 0) {
    if (next_cell_counts) return (int)cudaErrorInvalidValue;  // histogram needs the cells path
    int rc;
    if (order == 1)
      rc = update ? launch_deposit<1, 2>(p, np, E, B, current, g, tl, q, fq, st)
                  : launch_deposit<1, 1>(p, np, E, B, current, g, tl, q, fq, st);
    else
      rc = update ? launch_deposit<2, 2>(p, np, E, B, current, g, tl, q, fq, st)
                  : launch_deposit<2, 1>(p, np, E, B, current, g, tl, q, fq, st);
    if (rc) return rc;
  }
  if (update) {
    finalize_fused_ihole_kernel<<<1, 1, 0, st>>>(ihole, ntmax);
    SKB_CHECK_LAUNCH();
  }
  return 0;
}

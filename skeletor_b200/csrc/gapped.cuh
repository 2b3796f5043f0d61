// gapped.cuh — definitions shared by the gapped-layout kernels (gapped.cu: build, densify,
// insertion, the generic push; cellstream.cu: the TMA cell-stream push).
#pragma once
#include "common.cuh"
#include "gather.cuh"
#include "deposit.cuh"

#define GAP_THREADS 256

// ---- insertion of movers / arrivals ---------------------------------------------------
// rows whose x carries this bit pattern are padding of the mover list (unused tail of a
// warp's slot reservation) and are skipped
#define GAP_PAD_BITS 0x7ff8dead0badf00dLL
#define GAP_MCHUNK 64

// One slot claim per warp and destination cell (rows arrive roughly ordered by source
// tile, so a warp sees few distinct cells and its writes into one cell are contiguous).
#ifndef GAP_INS_ITEMS
#define GAP_INS_ITEMS 4
#endif
// Insert rows i0 + t*256 (t < GAP_INS_ITEMS, < n) into their cells; a warp-collective:
// all 32 lanes of a warp call it together.  GAP_INS_ITEMS independent rows per thread:
// the row load -> slot claim -> store chains overlap instead of adding up.
__device__ __forceinline__ void gap_insert_rows(const double *__restrict__ rows, int n, int i0,
                                                skb_particles_t P,
                                                const int *__restrict__ gap_start,
                                                int *gap_count, const KeyParams &kp,
                                                double *leftover, int leftover_cap,
                                                int *counts) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  double r0[GAP_INS_ITEMS], r1[GAP_INS_ITEMS], r2[GAP_INS_ITEMS], r3[GAP_INS_ITEMS],
      r4[GAP_INS_ITEMS];
  bool valid[GAP_INS_ITEMS];
#pragma unroll
  for (int t = 0; t < GAP_INS_ITEMS; t++) {
    const int i = i0 + t * 256;
    valid[t] = i < n;
    r0[t] = r1[t] = r2[t] = r3[t] = r4[t] = 0.0;
    if (valid[t]) {
      const double *r = rows + (size_t)i * 5;
      r0[t] = r[0]; r1[t] = r[1]; r2[t] = r[2]; r3[t] = r[3]; r4[t] = r[4];
    }
  }
  int s[GAP_INS_ITEMS], cap[GAP_INS_ITEMS], base[GAP_INS_ITEMS], rank[GAP_INS_ITEMS],
      leader[GAP_INS_ITEMS];
#pragma unroll
  for (int t = 0; t < GAP_INS_ITEMS; t++) {
    valid[t] = valid[t] && __double_as_longlong(r0[t]) != GAP_PAD_BITS;
    const int key = valid[t] ? cell_key(r0[t], r1[t], kp) : -1 - lane;
    const unsigned peers = __match_any_sync(SKB_FULL, key);
    const int cnt = __popc(peers);
    leader[t] = __ffs(peers) - 1; rank[t] = __popc(peers & lt);
    s[t] = cap[t] = base[t] = 0;
    if (valid[t] && lane == leader[t]) {
      base[t] = atomicAdd(gap_count + key, cnt);
      s[t] = gap_start[key]; cap[t] = gap_start[key + 1] - s[t];
      const int over = min(max(base[t] + cnt - cap[t], 0), cnt);
      if (over) atomicSub(gap_count + key, over);  // cell full: those go to the leftovers
    }
  }
#pragma unroll
  for (int t = 0; t < GAP_INS_ITEMS; t++) {
    const int ss = __shfl_sync(SKB_FULL, s[t], leader[t]);
    const int cc = __shfl_sync(SKB_FULL, cap[t], leader[t]);
    const int slot = __shfl_sync(SKB_FULL, base[t], leader[t]) + rank[t];
    if (!valid[t]) continue;
    if (slot < cc) {
      const long long d = (long long)ss + slot;
      P.x[d] = r0[t]; P.y[d] = r1[t]; P.vx[d] = r2[t]; P.vy[d] = r3[t]; P.vz[d] = r4[t];
    } else {
      const int l = atomicAdd(counts + 0, 1);
      if (l < leftover_cap) {
        const size_t lc = (size_t)leftover_cap;
        leftover[l] = r0[t]; leftover[lc + l] = r1[t]; leftover[2 * lc + l] = r2[t];
        leftover[3 * lc + l] = r3[t]; leftover[4 * lc + l] = r4[t];
      } else {
        counts[1] = 1;                            // even the leftover list is full
      }
    }
  }
}

// ---- parameters of the gapped push kernels ------------------------------------------------
struct GapPush {
  KickParams k;
  double dtdsx, dtdsy, vx_boost, x_boost;
  int flags;                 // SKB_EPI_SHEAR | SKB_EPI_PERIODIC_X
  KeyParams key;
  int *gap_count;            // updated in place
  double *movers;            // AoS rows
  int mover_cap;
  double *sbufl, *sbufr;
  int nbmax, rank, nvp;
  int *counts;               // [0] movers, [1] sbufl, [2] sbufr, [3] flags (1: mover list
                             // full -> some particles sit in the wrong cell, 2: nbmax),
                             // [4] movers re-inserted by their own block
  // movers whose new cell belongs to the same CTA are parked in a scratch block (claimed
  // from a pool for the lifetime of the CTA, L2 resident) and dropped into their cells
  // by the CTA itself once all its cells are compacted
  double *scratch;           // [npool][scratch_rows][5]
  int scratch_rows, npool;
  int *pool_owner;           // [npool] 0 = free
  const int *gap_start;
  double *leftover;          // SoA [5][leftover_cap]: rows whose cell is full
  int leftover_cap;
  int *lcounts;              // [0] leftover rows, [1] leftover overflow
};

// fused push_and_deposit on the gapped layout (push_and_deposit.pyx:10-170): sources
// grid, deposit offsets / shear, half-step drift factors
struct GapDeposit {
  double *cur;
  DepParams dp;
  double d2x, d2y;           // 0.5*dt/dx, 0.5*dt/dy, push_and_deposit.pyx:37-38
};


// cellstream.cu — the push on the gapped particle layout as a TMA-fed cell stream (sm_100a).
//
// Same work as push_gapped_kernel (gapped.cu): gather E,B -> Boris kick -> drift -> shear
// boost / x wrap -> route (stay in the cell, move to another cell, leave the slab), i.e.
// reference particles.py:159-188 / 233-257 in one pass, and optionally the full-step
// deposit that normally follows it (sources.py:27-50, deposit.pyx:6-34) fused in.  It is
// organised around what ncu showed about that kernel (514 warp instructions per 32
// particles, two thirds of the issue slots busy, 145 of them fp64 arithmetic):
//
//  * one warp streams the particles of one ROW of 16 cells of a 16 x 16 tile; a CTA of 4
//    warps owns a quarter of a tile and stages its own (4+5) x (16+5) window of E and B
//    (small CTAs: four are resident per SM, so the block-wide barrier in front of the
//    re-insertion phase of one of them is covered by the streams of the other three);
//  * particles arrive through a per-warp ring of 64-particle stages, each filled by ONE
//    TMA tensor copy (cp.async.bulk.tensor.2d + mbarrier complete_tx) of a [5 rows x 64
//    slots] box of the [5][Nmax] particle tensor - one elected lane issues it, nobody
//    computes per-lane load addresses; the tail of a cell (<= 32 particles) comes as a
//    [5 x 32] box with the same shared-memory row pitch (3-D view, see cs_particle_map32); (an optional mode processes a stage as two particles
//    per lane; with 128 registers per thread the compiler cannot interleave the two
//    chains and it is slower, see CS_NP2);
//  * all particles of a cell share the E stencil (the sort key IS the E-gather / deposit
//    base cell), so for CIC the 2 x 2 x 3 E values live in registers for the whole cell;
//    B is read per lane from the shared window (its base cell differs by the half-cell
//    offset, particle_push.pyx:15-19);
//  * stayers are written back compacted to the front of the cell's slot range;
//  * every particle that changes cell is staged as an AoS row in a per-warp shared-memory
//    buffer and flushed 32 rows (1280 B) at a time with ONE TMA bulk store
//    (cp.async.bulk.global.shared::cta) into the CTA's scratch block; after the stream the
//    CTA classifies those rows with all lanes busy: rows for its own cells go into their
//    free slots (warp-aggregated claims), the rest onto the global mover list;
//  * (PD = 3) stayers accumulate their full-step stencil sums in registers, one warp
//    reduce-scatter + emit per cell into a shared window of the sources grid; the rows
//    the CTA re-inserted are accumulated the same way in a second sweep over the cells'
//    new tails, rows that leave the CTA are deposited one by one.
//
// Per-particle arithmetic keeps the reference's operation order (-fmad=false), so the
// results are bit-identical to push_gapped_kernel / the oracle.
#include <cstdlib>
#include <cuda.h>
#include "gapped.cuh"

#ifndef CS_WARPS
#define CS_WARPS 4           // warps per CTA (8: half a tile, 4: a quarter)
#endif
#define CS_THREADS (32 * CS_WARPS)
#define CS_PARTS (16 / CS_WARPS)    // CTAs per tile
#define CS_CPW 16            // cells per warp: one row of the 16 x 16 tile
#define CS_CELLS (16 * CS_WARPS)    // cells per CTA
#ifndef CS_STAGE
#define CS_STAGE 64          // particles per ring stage (64 or 32)
#endif
#ifndef CS_NP2
#define CS_NP2 0              // 1: blocks of 64 particles run as two particles per lane (measured: no gain)
#endif
#ifndef CS_ABLATE
#define CS_ABLATE 0          // timing experiments only (results are wrong): 1 drop movers,
#endif                       // 2 no stayer stores, 4 no gather + kick, 8 no phase B
#ifndef CS_L2HINTS
#define CS_L2HINTS 0         // particle stream: L2 evict_first; parked rows: evict_last
#endif
#ifndef CS_BOX32
#define CS_BOX32 1           // stages with <= 32 particles left use the [5 x 32] box
#endif
#ifndef CS_HOIST_E
#define CS_HOIST_E 1         // (CIC) the cell's E stencil lives in registers
#endif
// Measured on config 5 (one launch, 1.07e9 particles; tools/build_variants.py +
// tools/run_ablate.sh, gpurun_out r2d-r2j): 4 warps x 4 CTAs/SM x 2 stages 22.0 ms;
// 8 warps x 2 CTAs x 3 stages 23.3; 4 x 4 x 3 stages 24.8; 3 CTAs of 8 warps at 80
// registers (32-particle stages) 24.0; 5 CTAs of 4 warps at 102 registers 27.3 (spills);
// two particles per lane 25.3; L2 evict_first / evict_last hints 23.2.
#ifndef CS_NST
#define CS_NST 2             // ring stages per warp
#endif
#ifndef CS_MINB
#define CS_MINB 4            // resident CTAs per SM aimed at
#endif
#define CS_MROWS 64          // mover rows buffered per warp: two halves of 32
#define CS_WS 21             // window stride in cells: 16 + SKB_HALO_LO + SKB_HALO_HI
#define CS_WR (CS_WARPS + 5)  // window rows: the CTA's rows + SKB_HALO_LO + SKB_HALO_HI
#define CS_WIN3 ((CS_WR * CS_WS * 3 + 15) & ~15)   // doubles reserved per Float3 window
#define CS_WIN4 ((CS_WR * CS_WS * 4 + 15) & ~15)   // doubles reserved for the Float4 window
#define CS_STAGE_D (5 * CS_STAGE)   // doubles per ring stage

// ---- PTX wrappers: mbarrier + TMA bulk copies ----------------------------------------
__device__ __forceinline__ unsigned cs_smem(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cs_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cs_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool cs_mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// L2 eviction policies: the particle stream is touched once per step (evict_first) while
// the parked mover rows are read back by the same CTA ~100 us later (evict_last)
__device__ __forceinline__ unsigned long long cs_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long cs_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// box of a 2-D tensor (global) -> shared, completion signalled on the mbarrier; c0 = slot
// (inner coordinate), c1 = row
__device__ __forceinline__ void cs_tma_load_2d(unsigned dst, const CUtensorMap *tm, int c0,
                                               int c1, unsigned bar) {
#if CS_L2HINTS
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0),
      "r"(c1), "l"(cs_policy_evict_first())
      : "memory");
#else
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
#endif
}
// the same for a 3-D tensor (see cs_particle_map32: a [5 x 32] box laid out with the row
// pitch of the [5 x 64] one)
__device__ __forceinline__ void cs_tma_load_3d(unsigned dst, const CUtensorMap *tm, int c0,
                                               int c1, int c2, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void cs_store_stream(double *p, double v) {
#if CS_L2HINTS
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v),
               "l"(cs_policy_evict_first())
               : "memory");
#else
  *p = v;
#endif
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void cs_bulk_store(void *dst, unsigned src, unsigned bytes) {
#if CS_L2HINTS
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(dst), "r"(src), "r"(bytes), "l"(cs_policy_evict_last())
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src), "r"(bytes)
               : "memory");
#endif
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cs_bulk_wait_read() {      // sources may be overwritten
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void cs_bulk_wait_all() {       // writes are complete
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void cs_fence_async_smem() {    // generic-proxy STS -> TMA reads
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Shared-memory accesses of the hot loop through 32-bit shared::cta addresses with
// immediate offsets: a generic pointer into shared memory costs a window-base computation
// (S2UR SR_CgaCtaId / ULEA) at every use.
template <int OFF>
__device__ __forceinline__ double cs_lds(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void cs_sts(unsigned a, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
// component K of gather_cic (particle_push.pxd:21-27) from the window cell at address a
template <int K>
__device__ __forceinline__ double cs_cic(unsigned a, double dx, double tx, double dy, double ty) {
  return dy * (dx * cs_lds<((CS_WS + 1) * 3 + K) * 8>(a) + tx * cs_lds<(CS_WS * 3 + K) * 8>(a)) +
         ty * (dx * cs_lds<(3 + K) * 8>(a) + tx * cs_lds<K * 8>(a));
}
// component K of gather_tsc (particle_push.pxd:58-67); a = cell (iy - 1, ix - 1)
template <int K>
__device__ __forceinline__ double cs_tsc(unsigned a, double wmx, double w0x, double wpx,
                                         double wmy, double w0y, double wpy) {
  return wmy * (wmx * cs_lds<K * 8>(a) + w0x * cs_lds<(3 + K) * 8>(a) + wpx * cs_lds<(6 + K) * 8>(a)) +
         w0y * (wmx * cs_lds<(CS_WS * 3 + K) * 8>(a) + w0x * cs_lds<(CS_WS * 3 + 3 + K) * 8>(a) +
                wpx * cs_lds<(CS_WS * 3 + 6 + K) * 8>(a)) +
         wpy * (wmx * cs_lds<(2 * CS_WS * 3 + K) * 8>(a) + w0x * cs_lds<(2 * CS_WS * 3 + 3 + K) * 8>(a) +
                wpx * cs_lds<(2 * CS_WS * 3 + 6 + K) * 8>(a));
}

// ---- phase B: the rows the CTA parked in its scratch block ---------------------------
// Rows i0 + t*256 (t < GAP_INS_ITEMS, < n) of `rows` (AoS, global): rows whose cell is one
// of this CTA's cells [c0, c0 + CS_CELLS) are dropped into the free slots of that cell
// (one slot claim per warp and destination cell, as gap_insert_rows); the others are
// appended to the global mover list for skb_gap_insert.  PD: rows that do not end up in
// one of the CTA's cells are deposited here, one by one (shared window or HBM atomics);
// the re-inserted ones are deposited from the cells' tails afterwards.  Warp-collective.
template <int ORDER, int PD>
__device__ __forceinline__ void cs_place_rows(const double *__restrict__ rows, int n, int i0,
                                              skb_particles_t P, const GapPush &q,
                                              const GapDeposit &dq, const DevGrid &g,
                                              const Window &w, double *sS, int c0,
                                              int &mbase, int &mused, const int *s_gs,
                                              int *s_cnt) {
  constexpr int NS = ORDER + 1;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  double r0[GAP_INS_ITEMS], r1[GAP_INS_ITEMS], r2[GAP_INS_ITEMS], r3[GAP_INS_ITEMS],
      r4[GAP_INS_ITEMS];
  bool valid[GAP_INS_ITEMS];
#pragma unroll
  for (int t = 0; t < GAP_INS_ITEMS; t++) {
    const int i = i0 + t * CS_THREADS;
    valid[t] = i < n;
    r0[t] = r1[t] = r2[t] = r3[t] = r4[t] = 0.0;
    if (valid[t]) {
      const double *r = rows + (size_t)i * 5;
      r0[t] = __ldcg(r); r1[t] = __ldcg(r + 1); r2[t] = __ldcg(r + 2);
      r3[t] = __ldcg(r + 3); r4[t] = __ldcg(r + 4);
    }
  }
  int key[GAP_INS_ITEMS], s[GAP_INS_ITEMS], cap[GAP_INS_ITEMS], base[GAP_INS_ITEMS],
      rank[GAP_INS_ITEMS], leader[GAP_INS_ITEMS];
  bool local[GAP_INS_ITEMS];
#pragma unroll
  for (int t = 0; t < GAP_INS_ITEMS; t++) {
    valid[t] = valid[t] && __double_as_longlong(r0[t]) != GAP_PAD_BITS;
    key[t] = valid[t] ? cell_key(r0[t], r1[t], q.key) : -1;
    local[t] = valid[t] && (unsigned)(key[t] - c0) < (unsigned)CS_CELLS;
    const unsigned peers = __match_any_sync(SKB_FULL, local[t] ? key[t] : -1 - lane);
    const int cnt = __popc(peers);
    leader[t] = __ffs(peers) - 1; rank[t] = __popc(peers & lt);
    s[t] = cap[t] = base[t] = 0;
    if (local[t] && lane == leader[t]) {
      // slot ranges and live counts of the CTA's cells are in shared memory: no trip to
      // L2 / HBM on this path
      const int lc = key[t] - c0;
      base[t] = atomicAdd(s_cnt + lc, cnt);
      s[t] = s_gs[lc]; cap[t] = s_gs[lc + 1] - s[t];
      const int over = min(max(base[t] + cnt - cap[t], 0), cnt);
      if (over) atomicSub(s_cnt + lc, over);      // cell full: those go to the leftovers
    }
  }
#pragma unroll
  for (int t = 0; t < GAP_INS_ITEMS; t++) {
    const int ss = __shfl_sync(SKB_FULL, s[t], leader[t]);
    const int cc = __shfl_sync(SKB_FULL, cap[t], leader[t]);
    const int slot = __shfl_sync(SKB_FULL, base[t], leader[t]) + rank[t];
    bool placed = false;
    if (local[t]) {
      if (slot < cc) {
        const long long d = (long long)ss + slot;
        P.x[d] = r0[t]; P.y[d] = r1[t]; P.vx[d] = r2[t]; P.vy[d] = r3[t]; P.vz[d] = r4[t];
        placed = true;
      } else {
        const int l = atomicAdd(q.lcounts + 0, 1);
        if (l < q.leftover_cap) {
          const size_t lc = (size_t)q.leftover_cap;
          q.leftover[l] = r0[t]; q.leftover[lc + l] = r1[t]; q.leftover[2 * lc + l] = r2[t];
          q.leftover[3 * lc + l] = r3[t]; q.leftover[4 * lc + l] = r4[t];
        } else {
          q.lcounts[1] = 1;                         // even the leftover list is full
        }
      }
    }
    // rows for other CTAs' cells: global mover list; slots are reserved GAP_MCHUNK at a
    // time per warp (the list head is ONE address for the whole grid: an atomic per
    // ballot would serialise in L2)
    const bool fwd = valid[t] && !local[t];
    const unsigned fm = __ballot_sync(SKB_FULL, fwd);
    if (fm) {
      const int k = __popc(fm), room = GAP_MCHUNK - mused;
      int nb = mbase;
      if (k > room) {
        if (lane == 0) nb = atomicAdd(q.counts + 0, GAP_MCHUNK);
        nb = __shfl_sync(SKB_FULL, nb, 0);
      }
      if (fwd) {
        const int r = __popc(fm & lt);
        const int ms = r < room ? mbase + mused + r : nb + (r - room);
        if (ms < q.mover_cap) {
          double *o = q.movers + (size_t)ms * 5;
          o[0] = r0[t]; o[1] = r1[t]; o[2] = r2[t]; o[3] = r3[t]; o[4] = r4[t];
        } else {
          // list full: park the row in whichever of this CTA's cells has a free slot
          // (flag 1: some particles sit in a wrong cell, the layout gets rebuilt)
          for (int a = 0; a < CS_CELLS && !placed; a++) {
            const int c = ((unsigned)key[t] + (unsigned)a) & (CS_CELLS - 1);
            const int cs = s_gs[c], cc2 = s_gs[c + 1] - cs;
            const int pos = atomicAdd(s_cnt + c, 1);
            if (pos < cc2) {
              const long long d = (long long)cs + pos;
              P.x[d] = r0[t]; P.y[d] = r1[t]; P.vx[d] = r2[t]; P.vy[d] = r3[t]; P.vz[d] = r4[t];
              placed = true;
            } else {
              atomicSub(s_cnt + c, 1);
            }
          }
          atomicOr(q.counts + 3, placed ? 1 : 8);   // 8: no room anywhere, particle lost
        }
      }
      if (k > room) {
        // the unused tail of the old chunk stays what it was made when it was reserved:
        // padding (see below)
        mbase = nb; mused = k - room;
        for (int r = mused + lane; r < GAP_MCHUNK; r += 32)
          if (mbase + r < q.mover_cap)
            q.movers[(size_t)(mbase + r) * 5] = __longlong_as_double(GAP_PAD_BITS);
      } else {
        mused += k;
      }
    }
    if constexpr (PD != 0) {
      if (valid[t] && !placed) {
        double xs = r0[t] + dq.dp.offx, ys = r1[t] + dq.dp.offy;
        if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
        int ix, iy;
        double wx[NS], wy[NS];
        particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
        const double vxr = r2[t] + dq.dp.S * (r1[t] * g.dy + g.y0);
        stray_particle_emit<NS>(wx, wy, ix, iy, vxr, r3[t], r4[t], sS, w, CS_WS, dq.cur, g);
      }
    }
  }
}

// Reserve room for 32 mover rows (one half of a warp's buffer): in the CTA's scratch block
// if it has room (returns the row), else on the global mover list (returns CS_GLOBAL | row;
// rows there bypass phase B: with PD the caller redoes the deposit, flag 16).  CS_NOSLOT
// when both are full: the warp then parks its movers in their old cells (flag 1: the
// layout is rebuilt).  Every reservation that succeeds is later written in full (rows or
// padding).  Warp-collective.
#define CS_NOSLOT (-1)
#define CS_GLOBAL (1 << 30)
template <int PD>
__device__ __forceinline__ int cs_reserve(int *s_nrows, int scr_rows, int *counts, double *movers,
                                       int mover_cap) {
  const int lane = threadIdx.x & 31;
  int slot = 0;
  if (lane == 0) slot = atomicAdd(s_nrows, 32);
  slot = __shfl_sync(SKB_FULL, slot, 0);
  if (slot + 32 <= scr_rows) return slot;
  // 33 rows: the list head may be odd (leftover rows), a TMA store needs 16-byte rows
  int gs = 0;
  if (lane == 0) gs = atomicAdd(counts + 0, 33);
  gs = __shfl_sync(SKB_FULL, gs, 0);
  if (gs + 33 <= mover_cap) {
    const int ge = (gs + 1) & ~1;
    if (lane == 0) {
      movers[(size_t)(gs == ge ? gs + 32 : gs) * 5] = __longlong_as_double(GAP_PAD_BITS);
      if (PD != 0) atomicOr(counts + 3, 16);
    }
    return CS_GLOBAL | ge;
  }
  // no room: whatever part of the reservation lies inside the list becomes padding
  for (int r = gs + lane; r < min(gs + 33, mover_cap); r += 32)
    movers[(size_t)r * 5] = __longlong_as_double(GAP_PAD_BITS);
  if (lane == 0) atomicOr(counts + 3, 1);
  return CS_NOSLOT;
}

// per-warp state of the mover staging
struct CsMovers {
  int count;      // rows staged so far by this warp
  int slot;       // reservation of the half being filled (CS_NOSLOT: lists full, park)
};

// Stage the movers of one ballot (<= 32 rows) as AoS rows in the warp's buffer; a half (32
// rows, 1280 B) that fills up goes out with one TMA bulk store to the destination that
// was reserved BEFORE its first row was staged.  Returns true for the lanes whose
// particle could not be staged (lists full): it stays in its old cell.  Warp-collective.
template <int PD>
__device__ __forceinline__ bool cs_stage_movers(bool mover, double x, double y, double vx,
                                                double vy, double vz, CsMovers &mv,
                                                unsigned mbuf_s, int *s_nrows, double *scr,
                                                int scr_rows, const GapPush &q) {
  const unsigned mm = __ballot_sync(SKB_FULL, mover);
  if (mm == 0) return false;
  if (mv.slot == CS_NOSLOT) return mover;            // parking
  const int lane = threadIdx.x & 31;
  const int pos = mv.count + __popc(mm & ((1u << lane) - 1u));
  const int after = mv.count + __popc(mm);
  const int boundary = (mv.count | 31) + 1;
  bool parked = false;
  if (after < boundary) {                             // the common case: no half completes
    if (mover) {
      const unsigned r = mbuf_s + (unsigned)(pos & (CS_MROWS - 1)) * 40u;
      cs_sts<0>(r, x); cs_sts<8>(r, y); cs_sts<16>(r, vx); cs_sts<24>(r, vy); cs_sts<32>(r, vz);
    }
    mv.count = after;
    return false;
  }
  // this ballot completes a half: the half about to be entered must have been read out
  if (lane == 0) cs_bulk_wait_read();
  __syncwarp();
  const int nxt = cs_reserve<PD>(s_nrows, scr_rows, q.counts, q.movers, q.mover_cap);
  if (mover) {
    if (nxt == CS_NOSLOT && pos >= boundary) {
      parked = true;
    } else {
      const unsigned r = mbuf_s + (unsigned)(pos & (CS_MROWS - 1)) * 40u;
      cs_sts<0>(r, x); cs_sts<8>(r, y); cs_sts<16>(r, vx); cs_sts<24>(r, vy); cs_sts<32>(r, vz);
    }
  }
  const int h = (mv.count >> 5) & 1;                  // the half that is complete now
  cs_fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    double *dst = (mv.slot & CS_GLOBAL) ? q.movers + (size_t)(mv.slot & ~CS_GLOBAL) * 5
                                        : scr + (size_t)mv.slot * 5;
    cs_bulk_store(dst, mbuf_s + (unsigned)h * 1280u, 1280u);
  }
  mv.slot = nxt;
  mv.count = nxt == CS_NOSLOT ? boundary : after;
  return parked;
}

// ---- per-particle work of one block of NP x 32 particles of a cell ------------------------
template <int ORDER, int PD>
struct CsCell {
  int cix, ciy, s, wcur;          // cell coordinates, first slot, stayers written so far
  bool fast;                      // every stencil of this cell lies inside the window
  double eC[4][3];                // (CIC) the E stencil of the cell
  double xsafe, ysafe;            // a position inside the cell for idle lanes
  Acc<ORDER + 1> acc;             // (PD) stencil sums of the cell
};

template <int ORDER, bool MODIFIED, int PD, int NP>
__device__ __forceinline__ void cs_block(unsigned pp, int nact,
                                         CsCell<ORDER, PD> &c, CsMovers &mv,
                                         skb_particles_t P, long long pstride,
                                         const double *sE, const double *sB, double *sS,
                                         unsigned sE_s, unsigned sB_s,
                                         const Window &w, const double *E, const double *B,
                                         const DevGrid &g, const GapPush &q,
                                         const GapDeposit &dq, unsigned mbuf_s, int *s_nrows,
                                         double *scr, int scr_rows) {
  // pp: shared address of this lane's first particle in the stage ([5][CS_STAGE] doubles);
  // sE_s, sB_s: shared addresses of the field windows sE, sB
  constexpr int NS = ORDER + 1;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const double nxd = (double)g.nx, nyd = (double)g.ny;
  bool act[NP];
  double x[NP], y[NP], vx[NP], vy[NP], vz[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    act[p] = p * 32 + lane < nact;
    // idle lanes compute on a harmless particle at rest inside the cell
    x[p] = c.xsafe; y[p] = c.ysafe; vx[p] = vy[p] = vz[p] = 0.0;
    if (act[p]) {
      const unsigned r = pp + (unsigned)p * 256u;
      x[p] = cs_lds<0>(r); y[p] = cs_lds<CS_STAGE * 8>(r); vx[p] = cs_lds<2 * CS_STAGE * 8>(r);
      vy[p] = cs_lds<3 * CS_STAGE * 8>(r); vz[p] = cs_lds<4 * CS_STAGE * 8>(r);
    }
  }
  // ---- gather + kick.  The cell promises the E base cell of its particles (the sort key)
  // and with it the B base cell up to + 1 (xb = xe + 1/2 before rounding, rounding is
  // monotonic); a lane that breaks the promise sends the warp down the generic path.
  double xe[NP], ye[NP], xb[NP], yb[NP];
  int ixe[NP], iye[NP], ixb[NP], iyb[NP];
  bool ok = true;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    xe[p] = x[p] + q.k.offEx; ye[p] = y[p] + q.k.offEy;
    xb[p] = x[p] + q.k.offBx; yb[p] = y[p] + q.k.offBy;
    if (ORDER == 2) {
      xe[p] = xe[p] + 0.5; ye[p] = ye[p] + 0.5; xb[p] = xb[p] + 0.5; yb[p] = yb[p] + 0.5;
    }
    ixe[p] = (int)xe[p]; iye[p] = (int)ye[p]; ixb[p] = (int)xb[p]; iyb[p] = (int)yb[p];
    ok = ok && ixe[p] == c.cix && iye[p] == c.ciy;
  }
  if ((CS_ABLATE & 4) || (q.flags & SKB_EPI_DRIFT_ONLY)) {
    // drift alone (particle_push.pyx:159-169): velocities stay as they are
  } else if (c.fast && __all_sync(SKB_FULL, ok)) {
#pragma unroll
    for (int p = 0; p < NP; p++) {
      double e[3], b[3];
      if constexpr (ORDER == 1) {
        const double dx = xe[p] - (double)ixe[p], tx = 1.0 - dx;
        const double dy = ye[p] - (double)iye[p], ty = 1.0 - dy;
#if CS_HOIST_E
#pragma unroll
        for (int k = 0; k < 3; k++)
          e[k] = dy * (dx * c.eC[3][k] + tx * c.eC[2][k]) + ty * (dx * c.eC[1][k] + tx * c.eC[0][k]);
#else
        const unsigned ep = sE_s + (unsigned)(((c.ciy - w.y0) * CS_WS + (c.cix - w.x0)) * 24);
        e[0] = cs_cic<0>(ep, dx, tx, dy, ty); e[1] = cs_cic<1>(ep, dx, tx, dy, ty);
        e[2] = cs_cic<2>(ep, dx, tx, dy, ty);
#endif
        const unsigned bp = sB_s + (unsigned)(((iyb[p] - w.y0) * CS_WS + (ixb[p] - w.x0)) * 24);
        const double dxb = xb[p] - (double)ixb[p], txb = 1.0 - dxb;
        const double dyb = yb[p] - (double)iyb[p], tyb = 1.0 - dyb;
        b[0] = cs_cic<0>(bp, dxb, txb, dyb, tyb); b[1] = cs_cic<1>(bp, dxb, txb, dyb, tyb);
        b[2] = cs_cic<2>(bp, dxb, txb, dyb, tyb);
      } else {
        // tsc_weights, particle_push.pxd:37-57 (xe, xb already carry the + 0.5)
        double wmx, w0x, wpx, wmy, w0y, wpy;
        {
          const double d = xe[p] - (double)ixe[p] - 0.5; w0x = 0.75 - d * d;
          const double h = 0.5 + d; wpx = 0.5 * (h * h); wmx = 1.0 - (w0x + wpx);
        }
        {
          const double d = ye[p] - (double)iye[p] - 0.5; w0y = 0.75 - d * d;
          const double h = 0.5 + d; wpy = 0.5 * (h * h); wmy = 1.0 - (w0y + wpy);
        }
        const unsigned ep = sE_s + (unsigned)(((c.ciy - 1 - w.y0) * CS_WS + (c.cix - 1 - w.x0)) * 24);
        e[0] = cs_tsc<0>(ep, wmx, w0x, wpx, wmy, w0y, wpy);
        e[1] = cs_tsc<1>(ep, wmx, w0x, wpx, wmy, w0y, wpy);
        e[2] = cs_tsc<2>(ep, wmx, w0x, wpx, wmy, w0y, wpy);
        {
          const double d = xb[p] - (double)ixb[p] - 0.5; w0x = 0.75 - d * d;
          const double h = 0.5 + d; wpx = 0.5 * (h * h); wmx = 1.0 - (w0x + wpx);
        }
        {
          const double d = yb[p] - (double)iyb[p] - 0.5; w0y = 0.75 - d * d;
          const double h = 0.5 + d; wpy = 0.5 * (h * h); wmy = 1.0 - (w0y + wpy);
        }
        const unsigned bp = sB_s + (unsigned)(((iyb[p] - 1 - w.y0) * CS_WS + (ixb[p] - 1 - w.x0)) * 24);
        b[0] = cs_tsc<0>(bp, wmx, w0x, wpx, wmy, w0y, wpy);
        b[1] = cs_tsc<1>(bp, wmx, w0x, wpx, wmy, w0y, wpy);
        b[2] = cs_tsc<2>(bp, wmx, w0x, wpx, wmy, w0y, wpy);
      }
      rescale_and_kick<MODIFIED>(e, b, g, q.k, y[p], vx[p], vy[p], vz[p]);
    }
  } else {
#pragma unroll                     // (static indices: x[], y[], ... must stay in registers)
    for (int p = 0; p < NP; p++)
      if (act[p])
        fields_and_kick<ORDER, MODIFIED>(sE, sB, w, CS_WS, E, B, g, q.k, x[p], y[p], vx[p],
                                         vy[p], vz[p]);
  }
  // ---- drift, boundary epilogue, new stencil-base cell --------------------------------
  bool odd = false;                                   // x wrap needed or leaver: rare
#pragma unroll
  for (int p = 0; p < NP; p++) {
    x[p] = x[p] + vx[p] * q.dtdsx;                    // drift_particle, particle_push.pxd:88-91
    y[p] = y[p] + vy[p] * q.dtdsy;
    if (q.flags & SKB_EPI_SHEAR) {                    // particle_boundary.pyx:41-49
      if (y[p] < 0.0) { x[p] = x[p] - q.x_boost; vx[p] = vx[p] - q.vx_boost; }
      if (y[p] >= nyd) { x[p] = x[p] + q.x_boost; vx[p] = vx[p] + q.vx_boost; }
    }
    odd = odd || (act[p] && (!(x[p] >= 0.0 && x[p] < nxd) || y[p] < g.e0 || y[p] >= g.e1));
  }
  bool leaver[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) leaver[p] = false;
  if (__any_sync(SKB_FULL, odd)) {
#pragma unroll
    for (int p = 0; p < NP; p++) {
      if (!act[p]) continue;
      if ((q.flags & SKB_EPI_PERIODIC_X) && !(x[p] >= 0.0 && x[p] < nxd)) x[p] = wrap_x(x[p], nxd);
      if (y[p] < g.e0 || y[p] >= g.e1) {              // leaves the slab: cppmove2's pack
        leaver[p] = true;
        double *buf; int slot; double yy = y[p];
        if (yy < g.e0) {
          if (q.rank == 0) yy += nyd;
          slot = atomicAdd(q.counts + 1, 1); buf = q.sbufl;
        } else {
          if (q.rank == q.nvp - 1) yy -= nyd;
          slot = atomicAdd(q.counts + 2, 1); buf = q.sbufr;
        }
        if (slot < q.nbmax) {
          double *r = buf + (size_t)slot * 5;
          r[0] = x[p]; r[1] = yy; r[2] = vx[p]; r[3] = vy[p]; r[4] = vz[p];
        } else {
          atomicOr(q.counts + 3, 2);
        }
      }
    }
  }
  // new stencil-base cell (== cell_key without the clamp: a particle outside the array is
  // a mover and gets clamped by the insertion) and, for PD, the deposit weights
  int nix[NP], niy[NP];
  double wx[NP][NS], wy[NP][NS];
  bool stay[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    double xs = x[p] + q.key.offx, ys = y[p] + q.key.offy;
    if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
    particle_terms<ORDER>(xs, ys, nix[p], niy[p], wx[p], wy[p]);
    stay[p] = act[p] && !leaver[p] && nix[p] == c.cix && niy[p] == c.ciy;
  }
  // ---- movers: staged in shared memory; stayers: compacted write-back ----------------
  bool parked[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    const bool mover = !(CS_ABLATE & 1) && act[p] && !leaver[p] && !stay[p];
    parked[p] = cs_stage_movers<PD>(mover, x[p], y[p], vx[p], vy[p], vz[p], mv, mbuf_s,
                                    s_nrows, scr, scr_rows, q);
    stay[p] = stay[p] || parked[p];
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    // compacted to the front of the cell's range: always behind the reads (everything up to
    // the end of this block is already in the ring)
    const unsigned sm = __ballot_sync(SKB_FULL, stay[p]);
    if (stay[p] && !(CS_ABLATE & 2)) {
      double *px = P.x + ((long long)c.s + c.wcur + __popc(sm & lt));
      cs_store_stream(px, x[p]); cs_store_stream(px + pstride, y[p]);
      cs_store_stream(px + 2 * pstride, vx[p]); cs_store_stream(px + 3 * pstride, vy[p]);
      cs_store_stream(px + 4 * pstride, vz[p]);
      if constexpr (PD != 0) {
        const double vxr = vx[p] + dq.dp.S * (y[p] * g.dy + g.y0);     // deposit.pxd:24
        if (!parked[p]) accumulate<ORDER>(c.acc, wx[p], wy[p], vxr, vy[p], vz[p]);
        else stray_particle_emit<NS>(wx[p], wy[p], nix[p], niy[p], vxr, vy[p], vz[p], sS, w,
                                     CS_WS, dq.cur, g);
      }
    }
    c.wcur += __popc(sm);
  }
}

// one warp reduction and one emit per cell (see deposit_cells_kernel)
template <int NS>
__device__ __forceinline__ void cs_emit_cell(Acc<NS> &acc, bool in_window, double *sS,
                                             const Window &w, double *cur, const DevGrid &g) {
  const int lane = threadIdx.x & 31;
  if constexpr (NS == 2) {
    warp_reduce_scatter<16>(acc.v, lane);
    if (lane < 16)
      emit_one<NS>(acc.v[0], scatter_index<16>(lane), in_window, acc.ix, acc.iy, sS, w, CS_WS,
                   cur, g);
  } else {
    warp_reduce_scatter<32>(acc.v, lane);
    warp_reduce_scatter<4>(acc.v + 32, lane);
    emit_one<NS>(acc.v[0], scatter_index<32>(lane), in_window, acc.ix, acc.iy, sS, w, CS_WS, cur,
                 g);
    if (lane < 4)
      emit_one<NS>(acc.v[32], 32 + scatter_index<4>(lane), in_window, acc.ix, acc.iy, sS, w,
                   CS_WS, cur, g);
  }
}

// ---- the kernel --------------------------------------------------------------------------
// PD = 0: push;  PD = 3: push + full-step deposit into dq.cur (raw sums, not normalised)
// tm64 / tm32: tensor maps of the [5][pstride] particle tensor with boxes of 5 x 64 and
// 5 x 32 slots
template <int ORDER, bool MODIFIED, int PD>
__global__ void __launch_bounds__(CS_THREADS, (ORDER == 2 && PD != 0) ? 1 : CS_MINB)
cell_stream_kernel(const __grid_constant__ CUtensorMap tm64,
                   const __grid_constant__ CUtensorMap tm32, skb_particles_t P,
                   long long pstride, const double *__restrict__ E,
                   const double *__restrict__ B, DevGrid g, GapPush q, GapDeposit dq) {
  constexpr int NS = ORDER + 1;
  constexpr int LO = (ORDER == 2) ? 1 : 0;
  extern __shared__ __align__(128) double smem[];
  double *rings = smem;                                       // [warps][CS_NST][5][CS_STAGE]
  double *mbufs = rings + CS_WARPS * CS_NST * CS_STAGE_D;     // [warps][CS_MROWS][5]
  double *sE = mbufs + CS_WARPS * CS_MROWS * 5;
  double *sB = sE + CS_WIN3;
  double *sS = sB + CS_WIN3;                                  // (PD) window of the sources
  unsigned long long *bars = (unsigned long long *)(sS + (PD ? CS_WIN4 : 0));
  __shared__ int s_blk, s_nrows;
  __shared__ int s_nstay[CS_CELLS];                // stayers of every cell (after the stream)
  __shared__ int s_cnt[CS_CELLS];                  // live particles incl. the re-inserted rows
  __shared__ int s_gs[CS_CELLS + 1];               // slot ranges of the CTA's cells

  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int tile = blockIdx.x / CS_PARTS, part = blockIdx.x % CS_PARTS;
  const int c0 = (tile << 8) + part * CS_CELLS;
  const int wc0 = c0 + wv * CS_CPW;
  // my cells: lane j < 16 holds the slot range and the count of cell wc0 + j
  const int my_cnt = lane < CS_CPW ? q.gap_count[wc0 + lane] : 0;
  const int my_start = lane < CS_CPW ? q.gap_start[wc0 + lane] : 0;
  if (!__syncthreads_or(my_cnt)) return;                     // nothing lives here
  const int bx = (tile % q.key.ntx) << 4;
  const int by = ((tile / q.key.ntx) << 4) + part * CS_WARPS;
  Window w;
  w.x0 = max(bx - SKB_HALO_LO, 0); w.y0 = max(by - SKB_HALO_LO, 0);
  w.x1 = min(bx + 16 + SKB_HALO_HI, g.mx); w.y1 = min(by + CS_WARPS + SKB_HALO_HI, g.myp);
  if (threadIdx.x == 0) {
    int b = -1;
    if (q.npool > 0) {
      b = blockIdx.x % q.npool;
      while (atomicCAS(q.pool_owner + b, 0, 1) != 0) b = (b + 1 == q.npool) ? 0 : b + 1;
    }
    s_blk = b; s_nrows = 0;
  }
  if (threadIdx.x < CS_CELLS) s_nstay[threadIdx.x] = 0;
  if (lane < CS_CPW) s_gs[wv * CS_CPW + lane] = my_start;
  if (threadIdx.x == 0) s_gs[CS_CELLS] = q.gap_start[c0 + CS_CELLS];
  const unsigned bar0 = cs_smem(bars + wv * CS_NST);
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < CS_NST; st++) cs_mbar_init(bar0 + 8 * st, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (PD != 0) zero_window(sS, CS_WIN4);
  if (!(q.flags & SKB_EPI_DRIFT_ONLY)) {
    stage_window(sE, E, w, CS_WS, g);
    stage_window(sB, B, w, CS_WS, g);
  }
  __syncthreads();
  double *const scr = s_blk >= 0 ? q.scratch + (size_t)s_blk * q.scratch_rows * 5 : nullptr;
  const int scr_rows = s_blk >= 0 ? (q.scratch_rows & ~31) : 0;    // whole halves only
  double *const ring = rings + (size_t)wv * (CS_NST * CS_STAGE_D);
  double *const mbuf = mbufs + (size_t)wv * (CS_MROWS * 5);
  // The shared-window base, made opaque so that it lives in ONE register: the compiler
  // otherwise recomputes it (S2R SR_CgaCtaId + shifts) in front of every shared access.
  unsigned sbase;
  asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(cs_smem(smem)));
  const unsigned ring_s = sbase + (unsigned)((ring - smem) * sizeof(double));
  const unsigned mbuf_s = sbase + (unsigned)((mbuf - smem) * sizeof(double));
  const unsigned sE_s = sbase + (unsigned)((sE - smem) * sizeof(double));
  const unsigned sB_s = sbase + (unsigned)((sB - smem) * sizeof(double));
  CsMovers mv;
  mv.count = 0;
  mv.slot = cs_reserve<PD>(&s_nrows, scr_rows, q.counts, q.movers, q.mover_cap);

#define CS_CNT(j) __shfl_sync(SKB_FULL, my_cnt, (j))
#define CS_ADVANCE(j, base)                                    \
  do {                                                         \
    base += CS_STAGE;                                          \
    if (base >= CS_CNT(j)) {                                   \
      base = 0; j++;                                           \
      while (j < CS_CPW && CS_CNT(j) == 0) j++;                \
    }                                                          \
  } while (0)
  // one stage = up to 64 particles of ONE cell: a [5 x 64] box, or [5 x 32] when no more
  // than 32 particles are left (the box may reach into the cell's free slots or beyond:
  // those columns are never looked at)
#define CS_FETCH(stage, j, base)                                                        \
  do {                                                                                  \
    const int fs_ = __shfl_sync(SKB_FULL, my_start, (j)) + (base);                      \
    const bool big_ = CS_STAGE == 64 && (!CS_BOX32 || CS_CNT(j) - (base) > 32);         \
    if (lane == 0) {                                                                    \
      const unsigned bar_ = bar0 + 8 * (stage);                                         \
      const unsigned dst_ = ring_s + (unsigned)((stage) * CS_STAGE_D) * 8u;             \
      /* both boxes fill a whole [5][CS_STAGE] stage (the short one pads with zeros) */ \
      cs_mbar_expect_tx(bar_, 5u * 8u * (unsigned)CS_STAGE);                            \
      if (big_ || CS_STAGE == 32) cs_tma_load_2d(dst_, CS_STAGE == 64 ? &tm64 : &tm32, fs_, 0, bar_); \
      else cs_tma_load_3d(dst_, &tm32, fs_, 0, 0, bar_);                                \
    }                                                                                   \
  } while (0)

  int fj = 0, fbase = 0;
  while (fj < CS_CPW && CS_CNT(fj) == 0) fj++;
  int cj = fj, cbase = 0;
#pragma unroll
  for (int st = 0; st < CS_NST; st++)
    if (fj < CS_CPW) { CS_FETCH(st, fj, fbase); CS_ADVANCE(fj, fbase); }
  int stage = 0, n = 0;
  unsigned phases = 0;                             // parity bit of every ring stage
  CsCell<ORDER, PD> c;
  c.ciy = by + wv; c.cix = 0; c.s = 0; c.wcur = 0; c.fast = false;
  c.ysafe = (double)c.ciy + 0.25 - q.k.offEy - (ORDER == 2 ? 0.5 : 0.0);
  c.xsafe = 0.0;
  while (cj < CS_CPW) {
    if (cbase == 0) {                              // first stage of a cell
      c.cix = bx + cj;
      c.s = __shfl_sync(SKB_FULL, my_start, cj);
      n = CS_CNT(cj);
      c.xsafe = (double)c.cix + 0.25 - q.k.offEx - (ORDER == 2 ? 0.5 : 0.0);
      // every stencil a particle filed under this cell can touch lies inside the window
      c.fast = c.cix - LO >= w.x0 && c.cix + 2 < w.x1 && c.ciy - LO >= w.y0 && c.ciy + 2 < w.y1;
      if (CS_HOIST_E && ORDER == 1 && c.fast) {
        const double *eb = sE + ((c.ciy - w.y0) * CS_WS + (c.cix - w.x0)) * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          c.eC[0][k] = eb[k]; c.eC[1][k] = eb[3 + k];
          c.eC[2][k] = eb[CS_WS * 3 + k]; c.eC[3][k] = eb[CS_WS * 3 + 3 + k];
        }
      }
      if constexpr (PD != 0) {
#pragma unroll
        for (int i = 0; i < NS * NS * 4; i++) c.acc.v[i] = 0.0;
        c.acc.ix = c.cix; c.acc.iy = c.ciy;
      }
    }
    while (!cs_mbar_try_wait(bar0 + 8 * stage, (phases >> stage) & 1u)) {}
    phases ^= 1u << stage;
    const unsigned pp = ring_s + (unsigned)(stage * CS_STAGE_D + lane) * 8u;
    const int nrem = n - cbase;
#if CS_NP2
    if (nrem > 32)
      cs_block<ORDER, MODIFIED, PD, 2>(pp, nrem, c, mv, P, pstride, sE, sB, sS, sE_s, sB_s, w, E, B,
                                       g, q, dq, mbuf_s, &s_nrows, scr, scr_rows);
    else
      cs_block<ORDER, MODIFIED, PD, 1>(pp, nrem, c, mv, P, pstride, sE, sB, sS, sE_s, sB_s, w, E, B,
                                       g, q, dq, mbuf_s, &s_nrows, scr, scr_rows);
#else
    {
#pragma unroll 1
      for (int u = 0; u < CS_STAGE / 32; u++) {
        if (u * 32 >= nrem) break;
        cs_block<ORDER, MODIFIED, PD, 1>(pp + (unsigned)u * 256u, nrem - u * 32, c, mv, P, pstride,
                                         sE, sB, sS, sE_s, sB_s, w, E, B, g, q, dq, mbuf_s,
                                         &s_nrows, scr, scr_rows);
      }
    }
#endif
    __syncwarp();                                  // stage fully read: refill it
    int nj = cj, nbase = cbase;
    CS_ADVANCE(nj, nbase);
    if (nj != cj) {                                // the cell is finished
      if (lane == 0) s_nstay[wv * CS_CPW + cj] = c.wcur;
      c.wcur = 0;
      if constexpr (PD != 0) cs_emit_cell<NS>(c.acc, c.fast, sS, w, dq.cur, g);
    }
    cj = nj; cbase = nbase;
    if (fj < CS_CPW) { CS_FETCH(stage, fj, fbase); CS_ADVANCE(fj, fbase); }
    stage = (stage + 1 == CS_NST) ? 0 : stage + 1;
  }
#undef CS_CNT
#undef CS_ADVANCE
#undef CS_FETCH
  // the rows still in the warp's buffer: the unused part of the half becomes padding (a
  // reservation is always written in full)
  if (mv.slot != CS_NOSLOT) {
    const int rem = mv.count & 31;
    const int h = (mv.count >> 5) & 1;
    if (lane >= rem) mbuf[(h * 32 + lane) * 5] = __longlong_as_double(GAP_PAD_BITS);
    cs_fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      double *dst = (mv.slot & CS_GLOBAL) ? q.movers + (size_t)(mv.slot & ~CS_GLOBAL) * 5
                                          : scr + (size_t)mv.slot * 5;
      cs_bulk_store(dst, cs_smem(mbuf + h * 160), 1280u);
    }
  }
  if (lane == 0) cs_bulk_wait_all();               // my scratch rows are in memory
  // phase B: all cells of this CTA are compacted; place the parked rows
  __syncthreads();
  if (threadIdx.x < CS_CELLS) s_cnt[threadIdx.x] = s_nstay[threadIdx.x];
  __syncthreads();
  const int nrows = min(s_nrows, scr_rows);        // (reservations are whole halves)
  if (nrows > 0 && !(CS_ABLATE & 8)) {
    if (threadIdx.x == 0) atomicAdd(q.counts + 4, nrows);            // statistics
    int mbase = 0, mused = GAP_MCHUNK;             // this warp's chunk of the global list
    // the rows were written ~100 us ago and have mostly left L2: pull them all back at
    // once instead of paying the HBM latency once per batch
    for (int i = (int)threadIdx.x * 16; i < nrows * 5; i += CS_THREADS * 16)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(scr + i));
    for (int r0 = 0; r0 < nrows; r0 += CS_THREADS * GAP_INS_ITEMS)
      cs_place_rows<ORDER, PD>(scr, nrows, r0 + (int)threadIdx.x, P, q, dq, g, w, sS, c0, mbase,
                               mused, s_gs, s_cnt);
  }
  __syncthreads();
  if (threadIdx.x < CS_CELLS) q.gap_count[c0 + threadIdx.x] = s_cnt[threadIdx.x];
  if (threadIdx.x == 0 && s_blk >= 0) atomicExch(q.pool_owner + s_blk, 0);
  if constexpr (PD != 0) {
    // the rows re-inserted above sit behind the stayers of their cells: accumulate them
    // like the stayers (same cell => same stencil), one reduction + emit per cell
    for (int j = 0; j < CS_CPW; j++) {
      const int n0 = s_nstay[wv * CS_CPW + j];
      const int n1 = s_cnt[wv * CS_CPW + j];
      if (n1 <= n0) continue;
      const int st = __shfl_sync(SKB_FULL, my_start, j);
      const int ax = bx + j;
      Acc<NS> &acc = c.acc;
#pragma unroll
      for (int i = 0; i < NS * NS * 4; i++) acc.v[i] = 0.0;
      acc.ix = ax; acc.iy = c.ciy;
      for (int i = n0 + lane; i < n1; i += 32) {
        const long long d = (long long)st + i;
        const double x = __ldcg(P.x + d), y = __ldcg(P.y + d), vx = __ldcg(P.vx + d),
                     vy = __ldcg(P.vy + d), vz = __ldcg(P.vz + d);
        double xs = x + dq.dp.offx, ys = y + dq.dp.offy;
        if (ORDER == 2) { xs = xs + 0.5; ys = ys + 0.5; }
        int ix, iy;
        double wx[NS], wy[NS];
        particle_terms<ORDER>(xs, ys, ix, iy, wx, wy);
        const double vxr = vx + dq.dp.S * (y * g.dy + g.y0);
        if (ix == acc.ix && iy == acc.iy) accumulate<ORDER>(acc, wx, wy, vxr, vy, vz);
        else stray_particle_emit<NS>(wx, wy, ix, iy, vxr, vy, vz, sS, w, CS_WS, dq.cur, g);
      }
      const bool in_window = ax - LO >= w.x0 && ax + 2 < w.x1 && c.ciy - LO >= w.y0 &&
                             c.ciy + 2 < w.y1;
      cs_emit_cell<NS>(acc, in_window, sS, w, dq.cur, g);
    }
    __syncthreads();
    flush_window(sS, w, CS_WS, dq.cur, g);
  }
}

// ---- launch ---------------------------------------------------------------------------------
size_t cell_stream_smem(int pd) {
  return (size_t)(2 * CS_WIN3 + (pd ? CS_WIN4 : 0) + CS_WARPS * CS_NST * CS_STAGE_D +
                  CS_WARPS * CS_MROWS * 5) * sizeof(double) +
         CS_WARPS * CS_NST * sizeof(unsigned long long);
}

typedef CUresult (*cs_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                 const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map of the [5][stride] float64 particle tensor at `base` with a [5 x box] box
static int cs_particle_map(CUtensorMap *tm, void *base, long long stride, int box) {
  static cs_encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return (int)e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
    encode = (cs_encode_fn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)stride, 5};
  const cuuint64_t strides[1] = {(cuuint64_t)stride * sizeof(double)};
  const cuuint32_t boxd[2] = {(cuuint32_t)box, 5};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, boxd, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// The SHORT box: [5 x 32] slots, but laid out in shared memory with the row pitch of the
// long one, so the kernel reads both with the same immediates.  A 3-D view {slots, 1, 5} of
// the same tensor with a box of {32, 2, 5}: the second slice of the middle dimension lies
// outside its extent of 1 and is zero-filled by the TMA unit without touching HBM.
static int cs_particle_map32(CUtensorMap *tm, void *base, long long stride) {
  CUtensorMap probe;
  int rc = cs_particle_map(&probe, base, stride, 32);       // (also resolves the entry point)
  if (rc) return rc;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) !=
          cudaSuccess || !fn)
    return (int)cudaErrorNotSupported;
  const cuuint64_t dims[3] = {(cuuint64_t)stride, 1, 5};
  const cuuint64_t strides[2] = {(cuuint64_t)stride * sizeof(double),
                                 (cuuint64_t)stride * sizeof(double)};
  const cuuint32_t boxd[3] = {32, 2, 5};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((cs_encode_fn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides,
                                  boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// pd: 0 = push, 3 = push + full-step deposit.  Returns cudaErrorNotSupported when the
// configuration is not the one this kernel is built for (the caller then uses the
// generic kernel of gapped.cu): tiles other than 16 x 16, or particle arrays that are not
// the rows of one [5][stride] tensor (the TMA descriptor needs a constant row pitch).
int cell_stream_launch(int pd, int order, int modified, skb_particles_t p, const double *E,
                       const double *B, const DevGrid &g, const GapPush &q,
                       const GapDeposit &dq, int ntiles, cudaStream_t st) {
  if (q.key.tlx != 4 || q.key.tly != 4) return (int)cudaErrorNotSupported;
  if (pd != 0 && pd != 3) return (int)cudaErrorNotSupported;
  const long long stride = p.y - p.x;
  if (stride <= 0 || (stride & 1) || p.vx - p.y != stride || p.vy - p.vx != stride ||
      p.vz - p.vy != stride || ((uintptr_t)p.x & 15))
    return (int)cudaErrorNotSupported;
  CUtensorMap tm64, tm32;
  int rc = cs_particle_map(&tm64, (void *)p.x, stride, 64);
  if (rc) return rc;
  rc = CS_STAGE == 64 ? cs_particle_map32(&tm32, (void *)p.x, stride)
                      : cs_particle_map(&tm32, (void *)p.x, stride, 32);
  if (rc) return rc;
  void (*k)(const CUtensorMap, const CUtensorMap, skb_particles_t, long long, const double *,
            const double *, DevGrid, GapPush, GapDeposit);
  if (pd == 0) {
    if (order == 1) k = modified ? cell_stream_kernel<1, true, 0> : cell_stream_kernel<1, false, 0>;
    else k = modified ? cell_stream_kernel<2, true, 0> : cell_stream_kernel<2, false, 0>;
  } else {
    if (order == 1) k = modified ? cell_stream_kernel<1, true, 3> : cell_stream_kernel<1, false, 3>;
    else k = modified ? cell_stream_kernel<2, true, 3> : cell_stream_kernel<2, false, 3>;
  }
  const size_t smem = cell_stream_smem(pd);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // the scratch-pool claim spins until a block is free: the pool must hold at least as
  // many blocks as CTAs can be resident
  if (q.npool > 0) {
    static int resident[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int slot = (order - 1) * 4 + (modified ? 2 : 0) + (pd ? 1 : 0);
    if (resident[slot] == 0) {
      int dev = 0, sms = 0, per = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k, CS_THREADS, smem);
      if (e != cudaSuccess) return (int)e;
      resident[slot] = max(per, 1) * max(sms, 1);
    }
    if (q.npool < resident[slot]) return (int)cudaErrorInvalidValue;
  }
  k<<<ntiles * CS_PARTS, CS_THREADS, smem, st>>>(tm64, tm32, p, stride, E, B, g, q, dq);
  SKB_CHECK_LAUNCH();
  return 0;
}

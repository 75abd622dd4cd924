// Symmetric fused Hamming kNN sweep for sm_100a: every unordered pair {i, j} of one table is
// evaluated once and feeds the neighbour lists of BOTH rows.
//
// Replaces the same reference code as pg_sweep.cuh (hamming.py:34 + sort/slice of
// prograph.py:757-762) for the case build_graph actually runs -- a table against itself --
// where d(i,j) == d(j,i) makes half of the reference's N x N evaluations redundant.
//
// Mapping.  Row block rb (256 rows, one per thread, plane words in registers) sweeps the
// stream tiles from its own diagonal to the end of the table:
//   * stream rows of blocks  < rb : skipped (block pair handled by the lower block);
//   * stream rows of block  == rb : "row side" only -- both (i,j) and (j,i) are seen here;
//   * stream rows of blocks  > rb : row side for i (key (d,j) into the thread's list in shared
//     memory) AND column side for j (key (d,i) into row j's list in global memory).
// Every row has ONE global list of k1 sorted keys d<<32|index behind a per-row spin lock, plus a
// 64-bit filter word glast[j] = (~distance)<<32 | index of its k1-th key that travels to shared
// memory with the stream tile (a second bulk copy on the same mbarrier).  The hot loop tests four stream
// rows at a time against both thresholds with eight IMADs (FMA pipe) and four LOP3s, one vote;
// everything behind the vote is rare and out of line.  A stale filter word is only ever larger
// than the current one, so no candidate is lost; the locked insertion compares full keys, so the
// result is the k1 smallest (distance, index) keys whatever the arrival order.
// Work items are (row block, tile range) chunks from a host-built table; a chunk starts its
// shared-memory lists empty but seeds its filter from the row's global list, and merges its lists
// into the global ones when it ends, so chunks can be short (good load balance) without the
// cold-start insertion storm of the split lists in pg_sweep.cuh.
// Bootstrap: the host first sweeps all rows one-sided against the first boot_rows columns
// (pg_hamming_knn_boot) so that every filter is tight from the first symmetric tile on; the row
// blocks of those rows ("boot" items) then run row side only, behind the bootstrap columns.
// Epsilon mode (SYM_EPS) is the same sweep with another consumer (prograph.py:731-753): a pair whose
// distance lies in [lo, hi] appends both directed edges as packed 64-bit keys to a global buffer,
// which the host sorts into the CSR (pg_edge_keys_to_csr) -- no count pass, no second sweep.
#pragma once
#include "pg_sweep.cuh"

namespace pg {

constexpr int kTauInf = 0x3fffffff;   // filter value of a list that is not full yet

struct SymItem { int rb, t0, t1, boot; };   // row block, stream tiles [t0, t1); boot: block of the bootstrap rows

struct SymParams {
  const uint32_t* tab;        // packed table, own == stream
  long long rows;             // valid rows
  const SymItem* items;
  int n_items;
  unsigned one;               // opaque 1 (see SweepParams::one)
  int k1;                     // list length, <= 32
  unsigned long long* glist;  // [rows][k1] ascending keys, ~0 = empty
  unsigned long long* glast;  // [n_tiles * tile_cols] filter word per row: (~tau)<<32 | index of the last key
  unsigned* glock;            // [rows]
  unsigned long long* stats;  // null, or [8] slow-path counters: -, locks, lock spins, list writes, row inserts
  long long boot_rows;        // rows [0, boot_rows) were swept one-sided against every row beforehand
  // epsilon mode (SYM_EPS): edge <=> lo <= d <= hi; edges are appended to `keys` as
  // row << sh_row | column << sh_col | d, both directions of every unordered pair
  int lo, hi;
  int sh_row, sh_col;
  unsigned long long* keys;
  long long capacity;            // slots in `keys`
  unsigned long long* counters;  // [0] slots reserved (multiples of kEdgeChunk), [1] edges written or dropped
};

__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_cg_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// filter word of a list whose last key is `last` (~0 = list not full)
__host__ __device__ __forceinline__ unsigned long long sym_filter_word(unsigned long long last) {
  const unsigned tau = (last == ~0ull) ? static_cast<unsigned>(kTauInf) : static_cast<unsigned>(last >> 32);
  return (static_cast<unsigned long long>(~tau) << 32) | (last & 0xffffffffull);
}
__device__ __forceinline__ int mad_s32(int a, unsigned b, int c) {
  int r;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// Warp-cooperative merge of up to 32 candidate keys (one per lane, `is_cand`) into the global
// list of row j.  All lanes call it with the same j.  The list is edited in registers under the
// row's lock; keys that no longer beat the current last key simply do not get in.
static __device__ __noinline__ void sym_serve_col(const SymParams& prm, long long j, unsigned long long key, bool is_cand,
                                           int lane) {
  const int k1 = prm.k1;
  unsigned long long* lst = prm.glist + static_cast<size_t>(j) * k1;
  // the caller's filter (a snapshot at most a few tiles old) already passed: take the lock
  // right away, the merge below re-checks every key against the current list
  unsigned cand = __ballot_sync(0xffffffffu, is_cand);
  if (cand == 0u) return;
  unsigned* lock = prm.glock + j;
  if (lane == 0) {
    unsigned spins = 0;
    while (atomicCAS(lock, 0u, 1u) != 0u) {
      __nanosleep(100);
      if (++spins > (1u << 24)) __trap();     // a protocol bug traps instead of hanging the GPU
    }
    // no fence on the acquiring side: the list is only ever read with ld.global.cg (L2, never L1),
    // the loads below are issued after the CAS has returned, and the previous holder fenced its
    // stores before it released the lock
    if (prm.stats != nullptr) {
      atomicAdd(prm.stats + 1, 1ull);
      if (spins) atomicAdd(prm.stats + 2, static_cast<unsigned long long>(spins));
    }
  }
  __syncwarp();
  unsigned long long e = lane < k1 ? ld_cg_u64(lst + lane) : ~0ull;
  bool changed = false;
  while (cand) {
    const int src = __ffs(cand) - 1;
    cand &= cand - 1;
    const unsigned long long kk = __shfl_sync(0xffffffffu, key, src);
    const int pos = __popc(__ballot_sync(0xffffffffu, lane < k1 && e < kk));
    const unsigned long long up = __shfl_up_sync(0xffffffffu, e, 1);
    if (pos < k1) {
      if (lane == pos) e = kk;
      else if (lane > pos) e = up;
      changed = true;
    }
  }
  if (changed) {
    if (lane < k1) st_cg_u64(lst + lane, e);
    const unsigned long long last = __shfl_sync(0xffffffffu, e, k1 - 1);
    if (lane == 0) {
      st_cg_u64(prm.glast + j, sym_filter_word(last));
      if (prm.stats != nullptr) atomicAdd(prm.stats + 3, 1ull);
    }
    __threadfence();
  }
  __syncwarp();
  if (lane == 0) atomicExch(lock, 0u);
}

constexpr int kMaxSymList = 32;   // longest list the symmetric sweep keeps (k + drop)
constexpr int kPendFlush = 24;    // a warp merges its pending column-side candidates at this many

// Lane-parallel merge (deferred column side, the DEFER instantiation): every active lane merges
// the candidate keys slots[b], b in slot_mask, into the global list of ITS OWN row -- up to 32 rows
// per call, so the L2 round trips of lock, list load, write-back and unlock overlap across the
// lanes instead of costing one warp stall per event.  The critical section sits inside the try-lock
// loop: a lane that got its lock finishes and releases it in the same iteration, whatever its
// sibling lanes are still waiting for, so locks held by diverged lanes of different warps cannot
// wait on each other.
static __device__ __noinline__ void sym_merge_lanes(const SymParams& prm, bool active, long long row,
                                                    const unsigned long long* slots, unsigned slot_mask) {
  const int k1 = prm.k1;
  bool done = !active;
  unsigned tries = 0;
  while (!__all_sync(0xffffffffu, done)) {
    if (!done && atomicCAS(prm.glock + row, 0u, 1u) == 0u) {
      unsigned long long* lst = prm.glist + static_cast<size_t>(row) * k1;
      unsigned long long e[kMaxSymList];
      for (int i = 0; i < k1; ++i) e[i] = ld_cg_u64(lst + i);
      bool changed = false;
      for (unsigned mk = slot_mask; mk != 0u; mk &= mk - 1u) {
        const unsigned long long key = slots[__ffs(mk) - 1];
        if (key >= e[k1 - 1]) continue;                // also skips empty slots (~0)
        int i = k1 - 1;
        for (; i > 0 && e[i - 1] > key; --i) e[i] = e[i - 1];
        e[i] = key;
        changed = true;
      }
      if (changed) {
        for (int i = 0; i < k1; ++i) st_cg_u64(lst + i, e[i]);
        st_cg_u64(prm.glast + row, sym_filter_word(e[k1 - 1]));
        __threadfence();
      }
      atomicExch(prm.glock + row, 0u);
      done = true;
      if (prm.stats != nullptr) {
        atomicAdd(prm.stats + 1, 1ull);
        if (changed) atomicAdd(prm.stats + 3, 1ull);
      }
    }
    if (++tries > (1u << 22)) __trap();                // a protocol bug traps instead of hanging the GPU
  }
}

// Merge the warp's pending (row, key) candidates: one lane per distinct row.  The queue length
// lives in shared memory (*pcnt, warp-private) so that it costs the hot loop no register.
static __device__ __noinline__ void sym_flush(const SymParams& prm, const unsigned* prow, const unsigned long long* pkey,
                                              int* pcnt, int lane) {
  __syncwarp();
  const int qn = *pcnt;
  if (qn == 0) return;
  const bool have = lane < qn;
  const unsigned row = have ? prow[lane] : 0xffffffffu;
  const unsigned grp = __match_any_sync(0xffffffffu, row);
  const bool leader = have && lane == __ffs(grp) - 1;
  sym_merge_lanes(prm, leader, row, pkey, grp);
  __syncwarp();
  if (lane == 0) *pcnt = 0;
  __syncwarp();
}

// Queue the column-side candidates of stream row j (lanes with `cc`).
static __device__ __noinline__ void sym_enqueue(const SymParams& prm, unsigned* prow, unsigned long long* pkey, int* pcnt,
                                                unsigned j, unsigned long long key, bool cc, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, cc);
  const int n = __popc(m);
  if (*pcnt + n > 32) sym_flush(prm, prow, pkey, pcnt, lane);
  const int qn = *pcnt;
  __syncwarp();
  if (cc) {
    const int slot = qn + __popc(m & ((1u << lane) - 1u));
    prow[slot] = j;
    pkey[slot] = key;
  }
  if (lane == 0) *pcnt = qn + n;
  __syncwarp();
  if (qn + n >= kPendFlush) sym_flush(prm, prow, pkey, pcnt, lane);
}

// Row side: the warp inserts the candidates of one stream row (column `col`) into the
// shared-memory lists of its own rows; returns the lane's updated filter.
static __device__ __noinline__ int sym_serve_row(unsigned long long* warp_lists, int k1, unsigned cand, unsigned dv,
                                          unsigned col, int lane, int tau_seed, int tau,
                                          unsigned long long* stats) {
  if (stats != nullptr && lane == 0) atomicAdd(stats + 4, static_cast<unsigned long long>(__popc(cand)));
  while (cand) {
    const int src = __ffs(cand) - 1;
    cand &= cand - 1;
    const unsigned dd = __shfl_sync(0xffffffffu, dv, src);
    const unsigned long long key = (static_cast<unsigned long long>(dd) << 32) | col;
    const unsigned t_new = knn_insert_coop(warp_lists + static_cast<size_t>(src) * k1, k1, key, lane);
    if (lane == src) tau = min(tau_seed, t_new == 0xffffffffu ? kTauInf : static_cast<int>(t_new));
  }
  return tau;
}

struct SymCursor {
  int idx, t, t1;
  __device__ __forceinline__ void start(int i, const SymParams& prm) {
    idx = i;
    if (idx < prm.n_items) {
      const int4 it = __ldg(reinterpret_cast<const int4*>(prm.items) + idx);
      t = it.y;
      t1 = it.z;
    } else {
      t = t1 = 0;
    }
  }
  __device__ __forceinline__ bool valid(const SymParams& prm) const { return idx < prm.n_items; }
  __device__ __forceinline__ void advance(const SymParams& prm, int stride) {
    if (++t == t1) start(idx + stride, prm);
  }
};

enum SymMode { SYM_KNN = 0, SYM_EPS = 1 };

constexpr int kEdgeChunk = 512;   // edge slots a warp reserves at a time (one global atomic per chunk)

// Epsilon mode: the warp appends the edges of one stream row.  `hit` lanes emit key_fwd (own row ->
// stream row) and, outside the diagonal block, key_rev (stream row -> own row).  Slots come from
// warp-private chunks of the global key buffer; unused slots of a closed chunk hold the sentinel ~0.
// Past the buffer's capacity nothing is written, but the reservations keep counting so that the
// caller learns the size it needs.
static __device__ __noinline__ void sym_emit(const SymParams& prm, bool hit, unsigned long long key_fwd,
                                             unsigned long long key_rev, bool both, int lane, long long& ch_base,
                                             int& ch_used, unsigned long long& n_real) {
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if (m == 0u) return;
  const int mult = both ? 2 : 1;
  const int n = __popc(m) * mult;
  if (ch_base == -1 || ch_used + n > kEdgeChunk) {
    if (ch_base >= 0)
      for (int sl = ch_used + lane; sl < kEdgeChunk; sl += 32) prm.keys[ch_base + sl] = ~0ull;
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(prm.counters, static_cast<unsigned long long>(kEdgeChunk));
    b = __shfl_sync(0xffffffffu, b, 0);
    ch_base = (static_cast<long long>(b) + kEdgeChunk <= prm.capacity) ? static_cast<long long>(b) : -2;
    ch_used = 0;
  }
  if (ch_base >= 0 && hit) {
    const int rank = __popc(m & ((1u << lane) - 1u));
    unsigned long long* at = prm.keys + ch_base + ch_used + rank * mult;
    at[0] = key_fwd;
    if (both) at[1] = key_rev;
  }
  ch_used += n;
  n_real += static_cast<unsigned long long>(n);
}

// DEFER (kNN mode only, experimental, PG_SYM_DEFER=1): column-side candidates are queued per warp and
// merged lane-parallel (sym_merge_lanes) instead of being served one event at a time.
template <int P, int W, int MODE, bool DEFER = false>
__global__ void __launch_bounds__(kSweepThreads, 2) sweep_sym_kernel(const __grid_constant__ SymParams prm) {
  constexpr int BN = TileCols<W>::value;
  constexpr int COLW = P * W;
  constexpr uint32_t STAGE_BYTES = BN * COLW * 4;
  constexpr uint32_t TAU_BYTES = MODE == SYM_KNN ? BN * 8 : 0;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* stage_mem = reinterpret_cast<uint32_t*>(smem_raw);
  unsigned long long* taus = reinterpret_cast<unsigned long long*>(smem_raw + kStages * STAGE_BYTES);   // [kStages][BN]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * (STAGE_BYTES + TAU_BYTES));
  unsigned* done = reinterpret_cast<unsigned*>(full + kStages);
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(full + 2 * kStages);   // [256][k1]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int k1 = prm.k1;
  unsigned long long* warp_lists = lists + static_cast<size_t>(warp << 5) * k1;
  // DEFER: per-warp queue of pending column-side candidates, behind the lists; the pointers are
  // rebuilt where they are needed (rare path) to keep them out of the hot loop's registers
  auto pend_key = [&]() { return lists + static_cast<size_t>(kConsumers) * prm.k1 + (threadIdx.x & ~31u); };
  auto pend_row = [&]() {
    return reinterpret_cast<unsigned*>(lists + static_cast<size_t>(kConsumers) * prm.k1 + kConsumers) + (threadIdx.x & ~31u);
  };
  auto pend_cnt = [&]() {
    return reinterpret_cast<int*>(lists + static_cast<size_t>(kConsumers) * prm.k1 + kConsumers) + kConsumers + (threadIdx.x >> 5);
  };
  if constexpr (DEFER) {
    if (lane == 0) *pend_cnt() = 0;
    __syncwarp();
  }

  auto fill_stage = [&](int s, int t) {   // one elected thread: tile t (and its filter words) into ring stage s
    mbar_arrive_expect_tx(&full[s], STAGE_BYTES + TAU_BYTES);
    bulk_g2s(stage_mem + s * (BN * COLW), prm.tab + static_cast<size_t>(t) * BN * COLW, STAGE_BYTES, &full[s]);
    if constexpr (MODE == SYM_KNN) bulk_g2s(taus + s * BN, prm.glast + static_cast<size_t>(t) * BN, TAU_BYTES, &full[s]);
  };

  SymCursor la;
  la.start(blockIdx.x, prm);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      done[s] = 0;
    }
    fence_mbar_init();
  }
  __syncthreads();
#pragma unroll 1
  for (int s = 0; s < kStages; ++s) {
    if (tid == 0 && la.valid(prm)) fill_stage(s, la.t);
    if (la.valid(prm)) la.advance(prm, gridDim.x);
  }

  int stage = 0;
  uint32_t phase = 0;
  const unsigned one = prm.one;
  // epsilon mode: the warp's current chunk of the edge buffer
  long long ch_base = -1;
  int ch_used = 0;
  unsigned long long n_real = 0;

  for (int ii = blockIdx.x; ii < prm.n_items; ii += gridDim.x) {
    const int4 it = __ldg(reinterpret_cast<const int4*>(prm.items) + ii);
    const int rb = it.x, t0 = it.y, t1 = it.z;
    const bool boot = it.w != 0;
    const long long r = static_cast<long long>(rb) * kConsumers + tid;
    const bool valid = r < prm.rows;
    uint32_t q[COLW];
    {
      const uint32_t* src = prm.tab + static_cast<size_t>(valid ? r : 0) * COLW;
#pragma unroll
      for (int j = 0; j < COLW; ++j) q[j] = valid ? __ldg(src + j) : 0u;
    }
    const uint32_t(&qq)[1][COLW] = reinterpret_cast<const uint32_t(&)[1][COLW]>(q);
    // kNN: the row's global list already bounds what can still matter: ties at its last distance
    // stay admissible (the index decides), hence the +1
    int tau_seed = 0, tau = 0;
    if constexpr (MODE == SYM_KNN) {
      if (valid) {
        const int g = ~static_cast<int>(ld_cg_u64(prm.glast + r) >> 32);
        tau_seed = g >= kTauInf ? kTauInf : g + 1;
      }
      tau = tau_seed;
      unsigned long long* mine = lists + static_cast<size_t>(tid) * k1;
      for (int j = 0; j < k1; ++j) mine[j] = ~0ull;
    }
    // invalid own rows never feed a column list; blocks of the bootstrap rows have no column side
    // (every list already holds its candidates among the bootstrap rows) and start behind them
    const unsigned vmask = (valid && !boot) ? 0xffffffffu : 0u;
    const long long diag_begin = boot ? prm.boot_rows : static_cast<long long>(rb) * kConsumers;
    const long long diag_end = boot ? (1ll << 40) : diag_begin + kConsumers;
    // epsilon mode: edge  <=>  lo <= d <= hi; an invalid own row gets an unreachable lower bound
    const int nlo = valid ? -prm.lo : -(1 << 28);
    const int hi = prm.hi;
    const unsigned mone = 0u - one;
    __syncwarp();

    for (int t = t0; t < t1; ++t) {
      mbar_wait(&full[stage], phase);
      const uint32_t* tile = stage_mem + stage * (BN * COLW);
      const unsigned long long* tnt = taus + stage * BN;
      const long long col0 = static_cast<long long>(t) * BN;
      const int ncols = static_cast<int>(min(static_cast<long long>(BN), prm.rows - col0));
      int c = static_cast<int>(min(static_cast<long long>(ncols), max(0ll, diag_begin - col0)));
      const int c_mid = static_cast<int>(min(static_cast<long long>(ncols), max(static_cast<long long>(c), diag_end - col0)));

      // one stream row behind the vote (rare): kNN serves both lists, epsilon appends the edges
      auto rare = [&](int dv, int cc_, bool col_on) {
        const unsigned col = static_cast<unsigned>(col0 + cc_);
        if constexpr (MODE == SYM_KNN) {
          const unsigned rc = __ballot_sync(0xffffffffu, dv < tau);
          if (rc) tau = sym_serve_row(warp_lists, k1, rc, static_cast<unsigned>(dv), col, lane, tau_seed, tau, prm.stats);
          if (col_on) {
            // exact test against the snapshot of row col's last key: ties are decided by the index
            const unsigned long long fw = tnt[cc_];
            const unsigned long long lastk = (static_cast<unsigned long long>(~static_cast<unsigned>(fw >> 32)) << 32) |
                                             (fw & 0xffffffffull);
            const unsigned long long mine = (static_cast<unsigned long long>(static_cast<unsigned>(dv)) << 32) |
                                            static_cast<unsigned>(r);
            const bool cc = vmask != 0u && mine < lastk;
            if (__any_sync(0xffffffffu, cc)) {
              if constexpr (DEFER) sym_enqueue(prm, pend_row(), pend_key(), pend_cnt(), col, mine, cc, lane);
              else sym_serve_col(prm, col, mine, cc, lane);
            }
          }
        } else {
          const bool hit = (dv + nlo) >= 0 && dv <= hi;
          const unsigned long long own = static_cast<unsigned long long>(r), oth = col;
          const unsigned long long dd = static_cast<unsigned long long>(static_cast<unsigned>(dv));
          sym_emit(prm, hit, (own << prm.sh_row) | (oth << prm.sh_col) | dd, (oth << prm.sh_row) | (own << prm.sh_col) | dd,
                   col_on, lane, ch_base, ch_used, n_real);
        }
      };

#pragma unroll 1
      for (; c + 4 <= ncols; c += 4) {
        int d0[1], d1[1], d2[1], d3[1];
        ham_rows<P, W, 1>(qq, tile + (c + 0) * COLW, d0, one);
        ham_rows<P, W, 1>(qq, tile + (c + 1) * COLW, d1, one);
        ham_rows<P, W, 1>(qq, tile + (c + 2) * COLW, d2, one);
        ham_rows<P, W, 1>(qq, tile + (c + 3) * COLW, d3, one);
        int any;
        if constexpr (MODE == SYM_KNN) {
          // filter words of the four stream rows: .y / .w = ~tau_j, .x / .z = index of the last key
          const int4 na = *reinterpret_cast<const int4*>(tnt + c);
          const int4 nb = *reinterpret_cast<const int4*>(tnt + c + 2);
          const unsigned cm = c >= c_mid ? vmask : 0u;
          // sign bit set <=> candidate: d - tau < 0 (row side), d + ~tau_j < 0 i.e. d <= tau_j (column side)
          const int s0 = mad_s32(d0[0], one, -tau), s1 = mad_s32(d1[0], one, -tau);
          const int s2 = mad_s32(d2[0], one, -tau), s3 = mad_s32(d3[0], one, -tau);
          const int u0 = mad_s32(d0[0], one, na.y), u1 = mad_s32(d1[0], one, na.w);
          const int u2 = mad_s32(d2[0], one, nb.y), u3 = mad_s32(d3[0], one, nb.w);
          any = (s0 | s1 | s2) | s3 | static_cast<int>(static_cast<unsigned>((u0 | u1 | u2) | u3) & cm);
        } else {
          // sign bit set <=> no edge: d - lo < 0 or hi - d < 0; all four miss <=> the AND keeps the sign
          const int a0 = mad_s32(d0[0], one, nlo), a1 = mad_s32(d1[0], one, nlo);
          const int a2 = mad_s32(d2[0], one, nlo), a3 = mad_s32(d3[0], one, nlo);
          const int b0 = mad_s32(d0[0], mone, hi), b1 = mad_s32(d1[0], mone, hi);
          const int b2 = mad_s32(d2[0], mone, hi), b3 = mad_s32(d3[0], mone, hi);
          any = ~((a0 | b0) & (a1 | b1) & (a2 | b2) & (a3 | b3));
        }
        if (__any_sync(0xffffffffu, any < 0)) {
          const bool col_on = c >= c_mid;
          rare(d0[0], c + 0, col_on);
          rare(d1[0], c + 1, col_on);
          rare(d2[0], c + 2, col_on);
          rare(d3[0], c + 3, col_on);
        }
      }
#pragma unroll 1
      for (; c < ncols; ++c) {   // ragged end of the table (last tile only)
        int d[1];
        ham_rows<P, W, 1>(qq, tile + c * COLW, d, one);
        rare(d[0], c, c >= c_mid);
      }

      __syncwarp();
      if (lane == 0) {
        if (atom_add_acq_rel_cta(&done[stage], 1u) == kConsumerWarps - 1) {
          done[stage] = 0;
          if (la.valid(prm)) {
            fence_proxy_async();
            fill_stage(stage, la.t);
          }
        }
      }
      if (la.valid(prm)) la.advance(prm, gridDim.x);
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }

    if constexpr (MODE == SYM_KNN) {
      // merge the chunk's lists into the rows' global lists
      __syncwarp();
      const long long wrow0 = static_cast<long long>(rb) * kConsumers + (warp << 5);
      if constexpr (DEFER) {
        sym_flush(prm, pend_row(), pend_key(), pend_cnt(), lane);
        const unsigned long long* ml = warp_lists + static_cast<size_t>(lane) * k1;     // every lane: its own row
        sym_merge_lanes(prm, wrow0 + lane < prm.rows && ml[0] != ~0ull, wrow0 + lane, ml,
                        k1 >= 32 ? 0xffffffffu : (1u << k1) - 1u);
      } else {
#pragma unroll 1
        for (int src = 0; src < 32; ++src) {   // ascending keys, one row at a time
          if (wrow0 + src >= prm.rows) break;
          const unsigned long long lk = lane < k1 ? warp_lists[static_cast<size_t>(src) * k1 + lane] : ~0ull;
          sym_serve_col(prm, wrow0 + src, lk, lk != ~0ull, lane);
        }
      }
      __syncwarp();
    }
  }
  if constexpr (MODE == SYM_EPS) {
    if (ch_base >= 0)
      for (int sl = ch_used + lane; sl < kEdgeChunk; sl += 32) prm.keys[ch_base + sl] = ~0ull;
    if (lane == 0 && n_real) atomicAdd(prm.counters + 1, n_real);
  }
}

struct SymLaunch {
  int grid;
  size_t list_bytes;
  cudaStream_t stream;
  int mode = SYM_KNN;
  int defer = 0;
};


// grid == 0: only report the resident grid (CTAs) through *resident
template <int P, int W, int MODE, bool DEFER = false>
int launch_sweep_sym_mode(const SymParams& prm, const SymLaunch& l, int* resident) {
  auto kern = sweep_sym_kernel<P, W, MODE, DEFER>;
  const size_t smem = static_cast<size_t>(kStages) * (TileCols<W>::value * P * W * 4 +
                                                      (MODE == SYM_KNN ? TileCols<W>::value * 8 : 0)) +
                      2 * kStages * sizeof(uint64_t) + l.list_bytes + (DEFER ? kConsumers * 12 + 64 : 0);
  PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int occ = 0;
  PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSweepThreads, smem));
  if (occ < 1) { set_error("symmetric sweep does not fit on an SM (smem %zu bytes)", smem); return PG_ERR_UNSUPPORTED; }
  if (resident) *resident = num_sms() * occ;
  if (l.grid <= 0) return PG_OK;
  kern<<<static_cast<unsigned>(l.grid), kSweepThreads, smem, l.stream>>>(prm);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

template <int P, int W>
int launch_sweep_sym(const SymParams& prm, const SymLaunch& l, int* resident) {
  if (l.mode == SYM_EPS) return launch_sweep_sym_mode<P, W, SYM_EPS>(prm, l, resident);
  if constexpr (P == 5 && W == 8) {      // the experimental deferred merge is only built for the bench shape
    if (l.defer) return launch_sweep_sym_mode<P, W, SYM_KNN, true>(prm, l, resident);
  }
  return launch_sweep_sym_mode<P, W, SYM_KNN>(prm, l, resident);
}

#define PG_DECL_SWEEP_SYM(P, W) int sweep_sym_p##P##_w##W(const SymParams& prm, const SymLaunch& l, int* resident);
PG_DECL_SWEEP_SYM(5, 1) PG_DECL_SWEEP_SYM(5, 2) PG_DECL_SWEEP_SYM(5, 4) PG_DECL_SWEEP_SYM(5, 8) PG_DECL_SWEEP_SYM(5, 16)
PG_DECL_SWEEP_SYM(8, 1) PG_DECL_SWEEP_SYM(8, 2) PG_DECL_SWEEP_SYM(8, 4) PG_DECL_SWEEP_SYM(8, 8)
#undef PG_DECL_SWEEP_SYM

}  // namespace pg

// Exact kNN in PCA space, stage 1 on the 5th-generation tensor cores: 3xTF32 candidate
// generation with tcgen05.mma, TMEM accumulators and a TMA-fed shared-memory ring.
//
// Same contract as knn_candidates_kernel (knn_f32.cu): per query row the K_keep smallest
// scores s = |y|^2 - 2 x.y, the K_keep-th as a threshold for the FP64 re-rank's certificate,
// reference tiles visited outwards along the Morton order and skipped when their projection
// box is out of reach.  What changes is the engine of the contraction:
//   * every centred FP32 coordinate is split into two TF32 numbers, x = hi + lo with
//     hi = tf32(x), lo = tf32(x - hi); the dot product is accumulated as
//     hi.hi + hi.lo + lo.hi (the lo.lo term is below 2^-22 |x||y| and is covered by the
//     certificate's error bound) -- three tcgen05.mma.kind::tf32 per 8 coordinates;
//   * a CTA owns 128 query rows whose hi/lo operand stays RESIDENT IN TENSOR MEMORY for the whole
//     kernel (lane = row, one column per coordinate: 2 x d8 of the 512 columns, written once with
//     tcgen05.st) and is read from there by the MMAs -- with both operands in shared memory a
//     128x128x8 TF32 MMA has to read 8 KB per 64 cycles, all of the SM's shared-memory bandwidth,
//     and measured 151 cycles; reference tiles (128 points) stream through a ring of 16-coordinate stages,
//     one bulk copy each (the global copy is stored directly in the UMMA "K-major, no swizzle"
//     core-matrix layout: 8 rows x 16 bytes per core matrix, LBO = 2048 B between K-adjacent
//     core matrices, SBO = 128 B between 8-row groups);
//   * the 128 x 128 FP32 accumulators take the TMEM columns the operand leaves (two at d = 100,
//     up to four for small d), so that the MMAs of the next tile run while the four epilogue warps
//     drain tile t with software-pipelined tcgen05.ld (32 lanes x 32 columns per instruction; lane =
//     query row, so a thread owns one row and its candidate buffer, which lives in shared memory);
//   * warp roles: warp 0 = producer (tile choice two steps ahead with 32 box tests per step, bulk
//     copies), warp 1 = MMA issuer (one lane chosen with elect.sync, so that the compiler emits the
//     UTCHMMA / UTCBAR instructions without a loop over the active lanes) and TMEM allocator,
//     warps 2-5 = epilogue.
// Supports d <= 128 (the resident operand and two accumulators must fit the 512 TMEM columns);
// larger d uses the SIMT kernel.
#include <cstdlib>
#include "common.cuh"

namespace pb200 {

constexpr int TC_BM = 128;        // query rows per CTA (TMEM lanes)
constexpr int TC_BN = 128;        // reference points per tile (TMEM columns per accumulator)
constexpr int TC_KC = 16;         // coordinates per ring stage (two MMA k-steps)
constexpr int TC_HALF = TC_BN * TC_KC * 4;      // bytes of one hi (or lo) block of a stage: 8 KB
constexpr int TC_STAGE = 2 * TC_HALF;           // hi + lo: 16 KB
constexpr int TC_KSTEP = TC_BN * 8 * 4;         // bytes of one k-step (8 coordinates) inside a block: 4 KB
constexpr int TC_CAP = 128;       // most candidate buffer entries per row (shared memory); the launch picks cap <= TC_CAP,
                                  // row stride cap + 1 64-bit words (odd: rows start on different banks), and compacts a
                                  // row above cap - 32 entries before its thread may push (<= 32 pushes per step)
constexpr int TC_IDLE_MARGIN = 20;       // rows with more than k_keep + this many entries are compacted in idle time
constexpr int TC_NST = 5;         // default ring stages (16 KB each)
constexpr int TC_NST_MAX = 8;     // mbarrier slots
constexpr int TC_MAX_KEEP = 48;
constexpr int TC_NACC = 4;        // most TMEM accumulators in flight (what 512 columns minus the operand allow)
constexpr int TC_THREADS = 192;   // 6 warps

struct TcParams {
  const float* xtc;       // [n_tiles][kch][hi|lo][4 kcores][16 rowgroups][8 rows][4] TF32 values
  const float* yn;        // |y_j|^2 per point (FP32, +inf on padding)
  const double* xcn;      // |x_i|^2 per point
  const double* box_lo;   // [n_tiles][3] or NULL
  const double* box_hi;
  int q0, nq, n_tiles;
  int kch;                // ring stages per tile = ceil(d_pad8 / 16)
  int tail_steps;         // k-steps in the last stage of a tile (1 or 2)
  int nst;                // ring depth
  int d8;                 // coordinates rounded up to 8 (TMEM columns of one half of the query operand)
  int nacc;               // TMEM accumulators in flight: (512 - 2 d8) / 128, 2..4
  int k_keep;
  int cap;                // candidate buffer entries per row (<= TC_CAP); the row stride is cap + 1 words
  int slack;              // in-stream compactions keep k_keep .. k_keep + slack entries (0: exactly k_keep)
  int idle_margin;        // rows with more than k_keep + this many entries are compacted in idle time (> slack)
  int debug;              // dev only (PB200_TC_DEBUG): 1 = epilogue releases accumulators unread, 2 = no candidate pushes,
                          // 3 = accumulators read (tcgen05.ld) but not processed
  unsigned long long* lists;   // unused (kept for the ABI: the candidate buffers live in shared memory)
  int* out_idx;
  float* out_score;
  float* out_thr;
  unsigned long long* tiles_visited;
  const int* qtile_list;  // optional: query tile (relative to q0) of every CTA of the launch
  int* cta_visited;       // optional: reference tiles visited, per CTA
  int visit_limit;        // > 0 (sampled launches): after this many tiles the producer only COUNTS the tiles that pass
                          // the box test against the bound reached so far, without fetching them
};

struct TcCtl {
  uint64_t a_full;
  uint64_t b_full[8], b_empty[8];
  uint64_t acc_full[TC_NACC], acc_empty[TC_NACC];
  uint64_t yn_full[TC_NACC + 1];
  int tile_ring[16];      // tile ids, producer -> MMA issuer (one entry per tile)
  int acc_tile[TC_NACC];
  int cnt[TC_BM];
  float thr[TC_BM];
  float xcn[TC_BM];
  float warp_bound[4];
  alignas(16) float yn[TC_NACC + 1][TC_BN];   // bulk-copy destination and float4 reads; one slot more than
                                              // accumulators: a tile's norms are still read after its accumulator is released
  double qlo[3], qhi[3];
  uint32_t tmem_base;
};

__device__ __forceinline__ unsigned long long tc_pack(float s, int idx) {
  return ((unsigned long long)f32_ordered(s) << 32) | (unsigned int)idx;
}

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  // K-major, SWIZZLE_NONE: start address, LBO = 2048 B (K-adjacent core matrices), SBO = 128 B (8-row groups),
  // descriptor version 1 (sm_100)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
         (1ull << 46);
}

// A operand read from TMEM (lane = row, one 32-bit column per coordinate), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                               uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tc_st4(uint32_t taddr, const float4& v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
               "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
               : "memory");
}

// exactly one lane of a converged warp (elect.sync): lets the compiler issue the uniform-datapath
// tcgen05 instructions without its per-active-thread ELECT loop
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of an accumulator -> registers, issue and wait separately (software pipelining):
// the registers are tied to the wait so that no use moves above it
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld32_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

__device__ __forceinline__ int tc_tile_of(int s, int q, int T) {
  const int R = T - q, L = q;
  const int m = R < L ? R : L;
  if (s < 2 * m) return (s & 1) ? q - 1 - (s >> 1) : q + (s >> 1);
  const int rest = s - 2 * m;
  return R > L ? q + m + rest : q - 1 - m - rest;
}

// Warp-collective compaction of one row's candidate buffer: keeps the k_keep smallest of its c <= 128
// entries (four per lane) and sets the row's threshold to the k_keep-th smallest score.  The k_keep-th
// smallest 32-bit key is found by a bitwise search below the keys' common prefix (one warp reduction per
// bit), then the survivors are written back by ballot-prefix positions -- no ranking: the order inside
// the buffer is irrelevant until the final emit.
__device__ __noinline__ void tc_compact_row(TcCtl* ctl, unsigned long long* buf, int row, int k_keep, int lane, int cap) {
  int c = ctl->cnt[row];
  c = c < cap ? c : cap;
  if (c <= k_keep) return;
  unsigned long long e[4];
  uint32_t key[4];
  uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = lane + 32 * q;
    e[q] = i < c ? buf[i] : ~0ull;
    key[q] = (uint32_t)(e[q] >> 32);
    mn = key[q] < mn ? key[q] : mn;
    if (i < c) mx = key[q] > mx ? key[q] : mx;
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  uint32_t piv = mn;
  if (mn != mx) {
    const int hb = 31 - __clz(mn ^ mx);
    piv = hb == 31 ? 0u : (mn & ~((2u << hb) - 1u));
    for (int b = hb; b >= 0; --b) {
      const uint32_t cand = piv | (1u << b);
      const int cl = (int)(key[0] < cand) + (int)(key[1] < cand) + (int)(key[2] < cand) + (int)(key[3] < cand);
      if ((int)__reduce_add_sync(0xffffffffu, (unsigned)cl) < k_keep) piv = cand;
    }
  }
  // piv = k_keep-th smallest key: everything below it stays, ties at piv fill the remaining slots
  const int nl = (int)(key[0] < piv) + (int)(key[1] < piv) + (int)(key[2] < piv) + (int)(key[3] < piv);
  const int n_less = (int)__reduce_add_sync(0xffffffffu, (unsigned)nl);
  const int allowed = k_keep - n_less;
  const unsigned lt = (1u << lane) - 1u;
  int base = 0, tbase = 0;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const bool less = key[q] < piv, tie = key[q] == piv;
    const unsigned bl = __ballot_sync(0xffffffffu, less), bt = __ballot_sync(0xffffffffu, tie);
    if (less) buf[base + __popc(bl & lt)] = e[q];
    if (tie) {
      const int tp = tbase + __popc(bt & lt);
      if (tp < allowed) buf[n_less + tp] = e[q];
    }
    base += __popc(bl);
    tbase += __popc(bt);
  }
  if (lane == 0) {
    ctl->thr[row] = f32_from_ordered(piv);
    ctl->cnt[row] = k_keep;
  }
  __syncwarp();
}

// The compaction used while the tiles stream by: it stops the bitwise search as soon as the bucket of keys still
// undecided holds at most `slack` entries beyond the k_keep-th, keeps everything below the bucket's top (between
// k_keep and k_keep + slack entries) and makes that top the row's threshold.  The threshold is then an upper bound of
// the k_keep-th smallest score instead of the score itself -- every entry dropped or never pushed still has k_keep kept
// entries that are no larger, which is all the exact final selection (tc_compact_row at the end) and the re-rank's
// certificate need -- and the search takes ~3-6 warp reductions instead of ~22 (the keys of a row share their leading
// bits; ncu: the exact routine was 13 % of the epilogue warps' time, and a compaction in progress delays the next
// accumulator by up to its whole length).  Many equal keys (the search runs out of bits): the exact routine.
__device__ __noinline__ void tc_compact_row_fast(TcCtl* ctl, unsigned long long* buf, int row, int k_keep, int lane, int cap,
                                                 int slack) {
  int c = ctl->cnt[row];
  c = c < cap ? c : cap;
  if (c <= k_keep + slack) return;
  unsigned long long e[4];
  uint32_t key[4];
  uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = lane + 32 * q;
    e[q] = i < c ? buf[i] : ~0ull;
    key[q] = (uint32_t)(e[q] >> 32);
    mn = key[q] < mn ? key[q] : mn;
    if (i < c) mx = key[q] > mx ? key[q] : mx;
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  bool done = false;
  uint32_t top = 0u;
  int hi_cnt = c;
  if (mn != mx) {
    const int hb = 31 - __clz(mn ^ mx);
    uint32_t piv = hb == 31 ? 0u : (mn & ~((2u << hb) - 1u));
    for (int b = hb; b >= 0; --b) {
      const uint32_t cand = piv | (1u << b);
      const int cl = (int)(key[0] < cand) + (int)(key[1] < cand) + (int)(key[2] < cand) + (int)(key[3] < cand);
      const int tot = (int)__reduce_add_sync(0xffffffffu, (unsigned)cl);
      if (tot < k_keep) {
        piv = cand;
      } else {
        top = cand;
        hi_cnt = tot;
        if (hi_cnt <= k_keep + slack) { done = true; break; }
      }
    }
  }
  if (!done) {
    __syncwarp();
    tc_compact_row(ctl, buf, row, k_keep, lane, cap);
    return;
  }
  // keep the hi_cnt entries below `top`
  const unsigned lt = (1u << lane) - 1u;
  int base = 0;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const bool keep = key[q] < top;
    const unsigned bl = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(bl & lt)] = e[q];
    base += __popc(bl);
  }
  if (lane == 0) {
    ctl->thr[row] = f32_from_ordered(top);
    ctl->cnt[row] = hi_cnt;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(TC_THREADS, 1) knn_candidates_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [B ring: nst stages][candidate buffers 128 x TC_LSTRIDE words][control]; the query operand lives in TMEM
  unsigned char* b_smem = smem_raw;
  unsigned long long* lists = reinterpret_cast<unsigned long long*>(b_smem + (size_t)p.nst * TC_STAGE);
  const int CAP = p.cap, LSTRIDE = p.cap + 1, HIGH = p.cap - 32;
  TcCtl& ctl = *reinterpret_cast<TcCtl*>(lists + TC_BM * LSTRIDE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // query tile of this CTA: the blockIdx-th of the launch, or the one listed (sampled launches: pb200_knn_tc_visit_counts)
  const int qrow0 = (p.qtile_list ? p.qtile_list[blockIdx.x] : (int)blockIdx.x) * TC_BM;   // first query row, relative to q0
  const int qbase = p.q0 + qrow0;
  const int row_local0 = blockIdx.x * TC_BM;       // first output row
  const int qtile = qbase / TC_BN;
  const int KCH = p.kch, NST = p.nst, T = p.n_tiles;

  for (int r = tid; r < TC_BM; r += TC_THREADS) {
    ctl.cnt[r] = 0;
    ctl.thr[r] = INFINITY;
    ctl.xcn[r] = (qrow0 + r < p.nq) ? __double2float_ru(p.xcn[qbase + r]) : -INFINITY;
  }
  if (tid < 4) ctl.warp_bound[tid] = (qrow0 + tid * 32 < p.nq) ? INFINITY : -INFINITY;
  if (tid == 0) {
    mbar_init(&ctl.a_full, 4);
    for (int s = 0; s < NST; ++s) { mbar_init(&ctl.b_full[s], 1); mbar_init(&ctl.b_empty[s], 1); }
    for (int b = 0; b < TC_NACC; ++b) { mbar_init(&ctl.acc_full[b], 1); mbar_init(&ctl.acc_empty[b], 4); }
    for (int b = 0; b <= TC_NACC; ++b) mbar_init(&ctl.yn_full[b], 1);
    fence_mbar_init();
    if (p.box_lo) {
      for (int a = 0; a < 3; ++a) { ctl.qlo[a] = p.box_lo[(long long)qtile * 3 + a]; ctl.qhi[a] = p.box_hi[(long long)qtile * 3 + a]; }
    }
  }
  if (warp == 1) {
    // TMEM, all 512 columns: nacc accumulators x 128 columns, then the query operand (hi | lo, d8 columns each)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;
  const uint32_t last_bytes = (uint32_t)p.tail_steps * TC_KSTEP;   // bytes of a hi (or lo) block in the last stage
  const int NACC = p.nacc;
  const uint32_t acol_hi = (uint32_t)(NACC * TC_BN), acol_lo = acol_hi + (uint32_t)p.d8;

  if (warp == 0) {
    // ================= producer =================
    int cursor = 0, st = 0, tslot = 0;     // st / ph (ring slot and pass parity) live in the issuing lane
    uint32_t ph = 0;
    unsigned long long visited = 0;
    // Tile choice, 32 candidates of the outward sequence per step (one per lane): the squared gap between
    // the candidate's projection box and the query tile's is computed two steps AHEAD (its six global loads
    // are in flight while the previous step's tiles are being copied) and compared with the pruning bound
    // only when the step is consumed; all accepted tiles of a step are kept in a mask and popped one by one.
    auto gap2_of = [&](int sidx) -> double {
      if (sidx >= T) return INFINITY;
      if (!p.box_lo) return 0.0;
      const int tile = tc_tile_of(sidx, qtile, T);
      double g2 = 0.0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double d1 = __ldg(p.box_lo + (long long)tile * 3 + a) - ctl.qhi[a];
        const double d2 = ctl.qlo[a] - __ldg(p.box_hi + (long long)tile * 3 + a);
        double gp = d1 > d2 ? d1 : d2;
        gp = gp > 0.0 ? gp : 0.0;
        g2 = fma(gp, gp, g2);
      }
      return g2 * (1.0 - 1e-9);
    };
    double g2_next = gap2_of(lane), g2_next2 = gap2_of(32 + lane);
    unsigned mask = 0;
    int mask_base = 0;
    for (;;) {
      // ---- choose the next tile ----
      int chosen = -1;
      if (mask == 0) {
        float bmax = *(volatile float*)&ctl.warp_bound[0];
        bmax = fmaxf(bmax, *(volatile float*)&ctl.warp_bound[1]);
        bmax = fmaxf(bmax, *(volatile float*)&ctl.warp_bound[2]);
        bmax = fmaxf(bmax, *(volatile float*)&ctl.warp_bound[3]);
        const double bound = (double)bmax;
        while (cursor < T) {
          const double g2 = g2_next;
          mask_base = cursor;
          cursor += 32;
          g2_next = g2_next2;
          g2_next2 = gap2_of(cursor + 32 + lane);
          mask = __ballot_sync(0xffffffffu, !(g2 > bound));     // candidates past the end are masked off below
          if (mask_base + 32 > T) mask &= (1u << (T - mask_base)) - 1u;
          if (mask) break;
        }
      }
      if (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        chosen = tc_tile_of(mask_base + bit, qtile, T);
      }
      // The copies are issued by ONE lane chosen with elect.sync (always the same lane for a full warp): under
      // `if (lane == 0)` nvcc wraps every UBLKCP in an ELECT ... R2UR.BROADCAST ... BRA.U.ANY loop over the active
      // lanes, and the issuing lane's own instruction stream (~70 instructions per 16 KB stage) was what bounded
      // the whole pipeline (ncu: the producer never waits for a free slot, the MMA thread waits for data).
      const bool issuer = tc_elect_one();
      if (chosen < 0) {
        // end of stream
        if (issuer) {
          mbar_wait(&ctl.b_empty[st], ph ^ 1u);
          ctl.tile_ring[tslot] = -1;
          mbar_arrive(&ctl.b_full[st]);
        }
        break;
      }
      ++visited;
      if (p.visit_limit > 0 && visited > (unsigned long long)p.visit_limit) continue;   // counted, not fetched
      if (issuer) {
        ctl.tile_ring[tslot] = chosen;           // published by the release of the first stage's arrive
        const float* src = p.xtc + (long long)chosen * KCH * (TC_STAGE / 4);
        const int full = p.tail_steps == 2 ? KCH : KCH - 1;      // stages copied whole
        for (int kc = 0; kc < full; ++kc) {
          mbar_wait(&ctl.b_empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&ctl.b_full[st], TC_STAGE);
          bulk_g2s(b_smem + (size_t)st * TC_STAGE, src, TC_STAGE, &ctl.b_full[st]);
          src += TC_STAGE / 4;
          if (++st == NST) { st = 0; ph ^= 1u; }
        }
        if (full < KCH) {
          // last stage of a tile with d8 % 16 == 8: one k-step of hi, one of lo
          mbar_wait(&ctl.b_empty[st], ph ^ 1u);
          unsigned char* dst = b_smem + (size_t)st * TC_STAGE;
          mbar_arrive_expect_tx(&ctl.b_full[st], 2 * last_bytes);
          bulk_g2s(dst, src, last_bytes, &ctl.b_full[st]);
          bulk_g2s(dst + TC_HALF, src + TC_HALF / 4, last_bytes, &ctl.b_full[st]);
          if (++st == NST) { st = 0; ph ^= 1u; }
        }
      }
      __syncwarp();
      tslot = (tslot + 1) & 15;
    }
    if (lane == 0 && p.tiles_visited) atomicAdd(p.tiles_visited, visited);
    if (lane == 0 && p.cta_visited) p.cta_visited[blockIdx.x] = (int)visited;
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (tc_elect_one()) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      mbar_wait(&ctl.a_full, 0);                 // query operand stored to TMEM by the epilogue warps
      tc_fence_after();
      const uint64_t b_desc0 = tc_smem_desc(smem_u32(b_smem));
      int st = 0, tslot = 0, buf = 0, ys = 0;
      uint32_t ph = 0, accph = 0;
      for (;;) {
        // ---- one reference tile: KCH ring stages into accumulator `buf` ----
        mbar_wait(&ctl.b_full[st], ph);
        const int tile = ctl.tile_ring[tslot];
        tslot = (tslot + 1) & 15;
        mbar_wait(&ctl.acc_empty[buf], accph ^ 1u);
        tc_fence_after();
        if (tile < 0) {
          // end of stream: tell the epilogue (no MMA pending on this buffer: plain arrive)
          ctl.acc_tile[buf] = -1;
          mbar_arrive(&ctl.acc_full[buf]);
          break;
        }
        // reference norms of this tile for the epilogue
        mbar_arrive_expect_tx(&ctl.yn_full[ys], TC_BN * 4);
        bulk_g2s(&ctl.yn[ys][0], p.yn + (long long)tile * TC_BN, TC_BN * 4, &ctl.yn_full[ys]);
        if (++ys > NACC) ys = 0;
        const uint32_t d_tmem = tmem + (uint32_t)buf * TC_BN;
        uint32_t a_hi = tmem + acol_hi, a_lo = tmem + acol_lo;
        for (int kc = 0; kc < KCH; ++kc) {
          if (kc > 0) {
            mbar_wait(&ctl.b_full[st], ph);
            tc_fence_after();
          }
          // three products per k-step: hi.hi + hi.lo + lo.hi; B descriptors differ only in the start-address field
          const uint64_t b_hi = b_desc0 + (uint64_t)((uint32_t)st * (TC_STAGE >> 4));
          tc_mma_tf32_ts(d_tmem, a_hi, b_hi, idesc, kc != 0 ? 1u : 0u);
          tc_mma_tf32_ts(d_tmem, a_hi, b_hi + (TC_HALF >> 4), idesc, 1u);
          tc_mma_tf32_ts(d_tmem, a_lo, b_hi, idesc, 1u);
          if (kc < KCH - 1 || p.tail_steps == 2) {
            tc_mma_tf32_ts(d_tmem, a_hi + 8, b_hi + (TC_KSTEP >> 4), idesc, 1u);
            tc_mma_tf32_ts(d_tmem, a_hi + 8, b_hi + ((TC_KSTEP + TC_HALF) >> 4), idesc, 1u);
            tc_mma_tf32_ts(d_tmem, a_lo + 8, b_hi + (TC_KSTEP >> 4), idesc, 1u);
          }
          tc_commit(&ctl.b_empty[st]);               // stage reusable once these MMAs have read it
          a_hi += TC_KC; a_lo += TC_KC;
          if (++st == NST) { st = 0; ph ^= 1u; }
        }
        ctl.acc_tile[buf] = tile;
        tc_commit(&ctl.acc_full[buf]);               // accumulator complete
        if (++buf == NACC) { buf = 0; accph ^= 1u; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5): TMEM lane quarter = warp % 4 =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // CTA-local query row owned by this thread
    unsigned long long* mybuf = lists + row * LSTRIDE;
    bool compacted_any = false;
    int c = 0;                                     // entries in this row's buffer (shared copy written before compactions)
    {
      // this thread's query row -> TMEM lane `row`: hi coordinates at columns acol_hi.., lo at acol_lo..
      const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
      const float* rowp = p.xtc + (long long)qtile * KCH * (TC_STAGE / 4) + (row >> 3) * 32 + (row & 7) * 4;
      for (int kc = 0; kc < KCH; ++kc) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = kc * TC_KC + kk * 4;
          if (k < p.d8) {
            const float* src = rowp + (long long)kc * (TC_STAGE / 4) + kk * 512;
            const float4 vh = __ldg(reinterpret_cast<const float4*>(src));
            const float4 vl = __ldg(reinterpret_cast<const float4*>(src + TC_HALF / 4));
            tc_st4(lane_addr + acol_hi + (uint32_t)k, vh);
            tc_st4(lane_addr + acol_lo + (uint32_t)k, vl);
          }
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.a_full);
    }
    int buf = -1, ys = -1;
    uint32_t accph = 1, yph = 1;
    for (;;) {
      if (++buf == NACC || buf == 0) { buf = 0; accph ^= 1u; }
      if (++ys > NACC || ys == 0) { ys = 0; yph ^= 1u; }
      // the next accumulator is not ready: use the time to compact this warp's fuller rows (tighter
      // thresholds earlier, and no bursts of compactions on the critical path later)
      while (!__all_sync(0xffffffffu, mbar_test(&ctl.acc_full[buf], accph))) {
        const unsigned fuller = __ballot_sync(0xffffffffu, c > p.k_keep + p.idle_margin);
        if (!fuller) break;
        const int l = __ffs(fuller) - 1;
        ctl.cnt[row] = c;
        __syncwarp();
        tc_compact_row_fast(&ctl, lists + (quarter * 32 + l) * LSTRIDE, quarter * 32 + l, p.k_keep, lane, CAP, p.slack);
        c = ctl.cnt[row];
        compacted_any = true;
      }
      mbar_wait(&ctl.acc_full[buf], accph);
      tc_fence_after();
      const int tile = ctl.acc_tile[buf];
      if (tile < 0) break;
      mbar_wait(&ctl.yn_full[ys], yph);
      const int colbase = tile * TC_BN;
      if (p.debug == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.acc_empty[buf]);
        continue;
      }
      // The four 32-column blocks of the tile, software pipelined: the tcgen05.ld of block cb+1 is in
      // flight while block cb is processed, and the accumulator is handed back to the MMA issuer as
      // soon as the last block sits in registers, before it is processed.
      float thr = ctl.thr[row];
      const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TC_BN);
      auto process = [&](uint32_t (&v)[32], const int cb) {
        if (p.debug == 3) return;                    // dev: accumulators read from tensor memory, nothing else
        if (p.debug == 4 && cb >= 2) return;         // dev: only half of the columns processed (wrong results; timing only)
        float q[8];                                  // minimum of each group of four columns
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 y = *reinterpret_cast<const float4*>(&ctl.yn[ys][cb * 32 + j4 * 4]);
          const float s0 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 0]), y.x);
          const float s1 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 1]), y.y);
          const float s2 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 2]), y.z);
          const float s3 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 3]), y.w);
          v[j4 * 4 + 0] = __float_as_uint(s0); v[j4 * 4 + 1] = __float_as_uint(s1);
          v[j4 * 4 + 2] = __float_as_uint(s2); v[j4 * 4 + 3] = __float_as_uint(s3);
          q[j4] = fminf(fminf(s0, s1), fminf(s2, s3));
        }
        const float m = fminf(fminf(fminf(q[0], q[1]), fminf(q[2], q[3])), fminf(fminf(q[4], q[5]), fminf(q[6], q[7])));
        if (__any_sync(0xffffffffu, m < thr) && p.debug != 2) {
          // room for this step's (at most 32) pushes: rows above TC_HIGH are compacted first
          const unsigned need = __ballot_sync(0xffffffffu, c > HIGH);
          if (need) {
            ctl.cnt[row] = c;
            __syncwarp();
            for (unsigned mm = need; mm; mm &= mm - 1) {
              const int l = __ffs(mm) - 1;
              tc_compact_row_fast(&ctl, lists + (quarter * 32 + l) * LSTRIDE, quarter * 32 + l, p.k_keep, lane, CAP, p.slack);
            }
            compacted_any = true;
            thr = ctl.thr[row];
            c = ctl.cnt[row];
          }
          // groups of four columns holding a candidate for any row of the warp; only those are scanned
          unsigned hit = 0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) hit |= (q[j4] < thr ? 1u : 0u) << j4;
          const unsigned any_hit = __reduce_or_sync(0xffffffffu, hit);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            if (any_hit >> j4 & 1u) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float sc = __uint_as_float(v[j4 * 4 + i]);
                if (sc < thr) {
                  mybuf[c] = tc_pack(sc, colbase + cb * 32 + j4 * 4 + i);
                  ++c;
                }
              }
            }
          }
        }
      };
      // (Measured and dropped: all four blocks into registers first and the accumulator released before anything is
      // processed -- 163.4 ms against 149.5 ms for the kNN call: the tensor-memory loads no longer overlap the
      // processing, and the time the accumulator is held was not what the MMA issuer waits for.)
      {
        uint32_t va[32], vb[32];
        tc_ld32_issue(tbase, va);
#pragma unroll 1
        for (int pair = 0; pair < 2; ++pair) {
          tc_ld32_wait(va);
          tc_ld32_issue(tbase + (uint32_t)(pair * 64 + 32), vb);
          process(va, pair * 2);
          tc_ld32_wait(vb);
          if (pair == 0) {
            tc_ld32_issue(tbase + 64u, va);
          } else {
            // every column of this accumulator has been read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl.acc_empty[buf]);
          }
          process(vb, pair * 2 + 1);
        }
      }
      if (__any_sync(0xffffffffu, compacted_any)) {
        float b = ctl.xcn[row] != -INFINITY ? __fadd_ru(ctl.thr[row], ctl.xcn[row]) : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
        if (lane == 0) ctl.warp_bound[quarter] = b;
        compacted_any = false;
      }
    }
    // ---- emit: final sort of each row of this warp ----
    ctl.cnt[row] = c;
    __syncwarp();
    for (int l = 0; l < 32; ++l) {
      const int r = quarter * 32 + l;
      const int grow = row_local0 + r;
      if (qrow0 + r >= p.nq) continue;
      tc_compact_row(&ctl, lists + r * LSTRIDE, r, p.k_keep, lane, CAP);   // leaves <= k_keep <= 48 entries
      int c = ctl.cnt[r];
      c = c < 64 ? c : 64;
      const unsigned long long* buf = lists + r * LSTRIDE;
      const unsigned long long e0 = lane < c ? buf[lane] : ~0ull;
      const unsigned long long e1 = lane + 32 < c ? buf[lane + 32] : ~0ull;
      int q0r = 0, q1r = 0;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const unsigned long long o0 = __shfl_sync(0xffffffffu, e0, j);
        const unsigned long long o1 = __shfl_sync(0xffffffffu, e1, j);
        q0r += (o0 < e0) + (o1 < e0);
        q1r += (o0 < e1) + (o1 < e1);
      }
      const int kk = p.k_keep;
      int* oi = p.out_idx + (long long)grow * kk;
      float* os = p.out_score + (long long)grow * kk;
      if (lane < c && q0r < kk) { oi[q0r] = (int)(uint32_t)e0; os[q0r] = f32_from_ordered((uint32_t)(e0 >> 32)); }
      if (lane + 32 < c && q1r < kk) { oi[q1r] = (int)(uint32_t)e1; os[q1r] = f32_from_ordered((uint32_t)(e1 >> 32)); }
      for (int s = c + lane; s < kk; s += 32) { oi[s] = -1; os[s] = INFINITY; }
      float thr_out = INFINITY;
      if (c >= kk) {
        float cand = -INFINITY;
        if (lane < c && q0r == kk - 1) cand = f32_from_ordered((uint32_t)(e0 >> 32));
        if (lane + 32 < c && q1r == kk - 1) cand = f32_from_ordered((uint32_t)(e1 >> 32));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = fmaxf(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        thr_out = cand;
      }
      if (lane == 0) p.out_thr[grow] = thr_out;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// Operand preparation: centred coordinates split into TF32 hi / lo, written in the UMMA K-major
// core-matrix layout, one 16 KB block per (tile, 16-coordinate stage): [hi 8 KB][lo 8 KB], each
// [4 kcores][16 rowgroups][8 rows][4 coordinates].
template <typename T>
__global__ void knn_prep_tc_kernel(const T* __restrict__ x, long long n, int d, int kch, long long n_pad,
                                   const double* __restrict__ mean, const int* __restrict__ perm,
                                   float* __restrict__ xtc) {
  // one thread per (point, 4-coordinate group)
  const long long groups = (long long)kch * 4;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_pad * groups) return;
  const long long i = e / groups;           // work-order point
  const int gq = (int)(e - i * groups);     // 4-coordinate group (global kcore index)
  const long long src = i < n ? (perm ? (long long)perm[i] : i) : -1;
  float hi[4], lo[4];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const int c = gq * 4 + c4;
    float v = 0.f;
    if (src >= 0 && c < d) v = (float)((double)x[src * d + c] - mean[c]);
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float hf = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hf));
    hi[c4] = hf;
    lo[c4] = __uint_as_float(l);
  }
  const long long tile = i >> 7;
  const int r = (int)(i & 127);
  const int kc = gq >> 2, kcore = gq & 3;
  float* blk = xtc + (tile * kch + kc) * (long long)(TC_STAGE / 4);
  const int off = (kcore * 16 + (r >> 3)) * 32 + (r & 7) * 4;   // floats
  *reinterpret_cast<float4*>(blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(blk + TC_HALF / 4 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------------------------------------
// TF32 tensor peak probe: the same instruction the candidate kernel issues (tcgen05.mma.cta_group::1.kind::tf32,
// M = N = 128, K = 8, A from tensor memory, B from shared memory), back to back, nothing else: one CTA per SM,
// `iters` MMAs per CTA rotating over four accumulators, operands resident (no loads inside the loop).  The
// roofline denominator of the candidate kernel is this kernel's measured FLOP/s (MEASURED_PEAKS.json has no
// TF32 figure).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters, int rot) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* b = reinterpret_cast<float*>(smem_raw);
  for (int i = tid; i < TC_HALF / 4; i += 128) b[i] = 1.0f + 0.001f * (float)(i & 63);
  if (tid == 0) { mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  {
    // A operand: 8 columns of finite values in every lane (columns 504..511)
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16) + 504u;
    const float4 v = make_float4(1.0f + 0.01f * lane, 0.5f, 0.25f, 0.125f);
    tc_st4(lane_addr, v);
    tc_st4(lane_addr + 4, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1 && tc_elect_one()) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint64_t bdesc = tc_smem_desc(smem_u32(b));
    for (int i = 0; i < iters; ++i) tc_mma_tf32_ts(tmem + (uint32_t)(i & rot) * TC_BN, tmem + 504u, bdesc, idesc, i >= 4 ? 1u : 0u);
    tc_commit(&done_bar);
  }
  if (warp == 1) mbar_wait(&done_bar, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// Floats needed for the tensor-core operand copy of n_pad points with d coordinates (0 if unsupported).
long long pb200_knn_tc_operand_floats(long long n_pad, int d) {
  const int d8 = (d + 7) / 8 * 8;
  if (d8 > 128) return 0;
  const int kch = (d8 + 15) / 16;
  return n_pad * (long long)kch * 32;
}

int pb200_knn_prepare_tc(const void* x, int is_f64, long long n, int d, long long n_pad, const double* mean,
                         const int* perm, float* xtc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int d8 = (d + 7) / 8 * 8;
  if (d8 > 128 || n_pad % 256 != 0 || n_pad < n) { set_error_msg("pb200_knn_prepare_tc: need d <= 128, n_pad % 256 == 0"); return -1; }
  const int kch = (d8 + 15) / 16;
  const long long total = n_pad * (long long)kch * 4;
  if (is_f64) knn_prep_tc_kernel<double><<<div_up(total, 256), 256, 0, st>>>((const double*)x, n, d, kch, n_pad, mean, perm, xtc);
  else        knn_prep_tc_kernel<float><<<div_up(total, 256), 256, 0, st>>>((const float*)x, n, d, kch, n_pad, mean, perm, xtc);
  PB_CHECK_LAUNCH("knn_prep_tc_kernel");
  return 0;
}

// Entries per row of the candidate scratch passed to pb200_knn_candidates_tc.
int pb200_knn_tc_list_cap(void) { return 1; }   // candidate buffers live in shared memory; `lists` is not used

// Tensor-core version of pb200_knn_candidates_f32 (same outputs; q0 % 128 == 0; lists:
// ceil(nq/128)*128*pb200_knn_tc_list_cap() words).
static int launch_knn_tc(const float* xtc, long long n_pad, const float* yn, const double* xcn,
                         const double* box_lo, const double* box_hi, int q0, int nq, int d, int k_keep,
                         unsigned long long* lists, int* out_idx, float* out_score, float* out_thr,
                         unsigned long long* tiles_visited, const int* qtile_list, int n_list, int* cta_visited,
                         int visit_limit, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int d8 = (d + 7) / 8 * 8;
  if (d8 > 128 || k_keep < 1 || k_keep > TC_MAX_KEEP || q0 % TC_BM != 0 || n_pad % 256 != 0) {
    set_error_msg("pb200_knn_candidates_tc: need d <= 128, k_keep <= 48, q0 % 128 == 0");
    return -1;
  }
  if (nq <= 0) return 0;
  TcParams p;
  p.xtc = xtc; p.yn = yn; p.xcn = xcn; p.box_lo = box_lo; p.box_hi = box_hi;
  p.q0 = q0; p.nq = nq; p.n_tiles = (int)(n_pad / TC_BN);
  p.kch = (d8 + 15) / 16;
  p.tail_steps = (d8 % 16 == 8) ? 1 : 2;
  p.k_keep = k_keep; p.lists = lists; p.out_idx = out_idx; p.out_score = out_score; p.out_thr = out_thr;
  p.tiles_visited = tiles_visited;
  p.qtile_list = qtile_list;
  p.cta_visited = cta_visited;
  p.visit_limit = visit_limit;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("PB200_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.d8 = d8;
  p.nacc = (512 - 2 * d8) / TC_BN;
  if (p.nacc > TC_NACC) p.nacc = TC_NACC;
  { const char* e = getenv("PB200_TC_NACC"); if (e && atoi(e) >= 1 && atoi(e) < p.nacc) p.nacc = atoi(e); }   // dev: fewer accumulators
  // ring depth and candidate buffer size share the shared memory: nst * 16 KB + 128 * (cap + 1) * 8 B + control <= 227 KB
  int nst = TC_NST, cap = TC_CAP;
  { const char* e = getenv("PB200_TC_NST"); if (e && atoi(e) >= 2 && atoi(e) <= TC_NST_MAX) nst = atoi(e); }
  { const char* e = getenv("PB200_TC_CAP"); if (e && atoi(e) >= k_keep + 33 && atoi(e) <= TC_CAP) cap = atoi(e); }
  p.nst = nst;
  p.cap = cap;
  p.slack = 4;                                     // < TC_IDLE_MARGIN, so that an idle-time compaction always shrinks its row (0 / 4 / 8 / 12: 152.0 / 148.4 / 149.9 / 150.4 ms at config 3)
  p.idle_margin = TC_IDLE_MARGIN;
  { const char* e = getenv("PB200_TC_IDLE"); if (e && atoi(e) >= 1 && atoi(e) <= 64) p.idle_margin = atoi(e); }
  { const char* e = getenv("PB200_TC_SLACK"); if (e && atoi(e) >= 0 && atoi(e) < p.idle_margin) p.slack = atoi(e); }
  if (p.slack >= p.idle_margin) p.slack = p.idle_margin - 1;
  const int smem = nst * TC_STAGE + TC_BM * (cap + 1) * 8 + (int)sizeof(TcCtl) + 64;
  if (smem > 227 * 1024) { set_error_msg("pb200_knn_candidates_tc: ring depth x candidate buffers exceed shared memory"); return -1; }
  PB_CHECK(cudaFuncSetAttribute(knn_candidates_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  knn_candidates_tc_kernel<<<qtile_list ? n_list : div_up(nq, TC_BM), TC_THREADS, smem, st>>>(p);
  PB_CHECK_LAUNCH("knn_candidates_tc_kernel");
  return 0;
}

int pb200_knn_candidates_tc(const float* xtc, long long n_pad, const float* yn, const double* xcn,
                            const double* box_lo, const double* box_hi, int q0, int nq, int d, int k_keep,
                            unsigned long long* lists, int* out_idx, float* out_score, float* out_thr,
                            unsigned long long* tiles_visited, void* stream) {
  return launch_knn_tc(xtc, n_pad, yn, xcn, box_lo, box_hi, q0, nq, d, k_keep, lists, out_idx, out_score, out_thr,
                       tiles_visited, nullptr, 0, nullptr, 0, stream);
}

// The candidate kernel on a SAMPLE of query tiles (qtiles[n_list], each < ceil(n / 128)): counts[i] = reference tiles
// the CTA of qtiles[i] visited -- or, with visit_limit > 0, the tiles it visited up to that limit plus the tiles that
// still pass the box test against the pruning bound reached by then (an estimate that costs visit_limit tiles per CTA
// instead of ~1000; the bound after a dozen tiles is looser than the final one, alike for every query tile).  The ranks of a sharded run cut the query rows where the cumulative counts are equal
// (the work of a query tile is the tiles it visits: 5 % to 25 % of them along a trajectory).  Scratch outputs:
// out_idx / out_score [n_list * 128 * k_keep], out_thr [n_list * 128].
int pb200_knn_tc_visit_counts(const float* xtc, long long n_pad, const float* yn, const double* xcn,
                              const double* box_lo, const double* box_hi, long long n, int d, int k_keep,
                              const int* qtiles, int n_list, int visit_limit, int* out_idx, float* out_score,
                              float* out_thr, int* counts, void* stream) {
  if (n_list <= 0) return 0;
  return launch_knn_tc(xtc, n_pad, yn, xcn, box_lo, box_hi, 0, (int)n, d, k_keep, nullptr, out_idx, out_score, out_thr,
                       nullptr, qtiles, n_list, counts, visit_limit, stream);
}

int pb200_tf32_peak_rot(int blocks, int iters, int rot, void* stream);
// TF32 tensor peak probe (tf32_peak_kernel): `blocks` CTAs, `iters` 128x128x8 MMAs each = blocks * iters * 262144 FLOP.
int pb200_tf32_peak(int blocks, int iters, void* stream) { return pb200_tf32_peak_rot(blocks, iters, 3, stream); }

// rot = 0, 1 or 3: the MMAs rotate over 1, 2 or 4 accumulators (dependent-accumulation probe)
int pb200_tf32_peak_rot(int blocks, int iters, int rot, void* stream) {
  if (blocks < 1 || iters < 4) { set_error_msg("pb200_tf32_peak: need blocks >= 1, iters >= 4"); return -1; }
  PB_CHECK(cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_HALF + 1024));
  tf32_peak_kernel<<<blocks, 128, TC_HALF + 1024, (cudaStream_t)stream>>>(iters, rot);
  PB_CHECK_LAUNCH("tf32_peak_kernel");
  return 0;
}

}  // extern "C"

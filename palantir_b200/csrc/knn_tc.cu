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
//   * a CTA owns 128 query rows whose hi/lo operand stays RESIDENT in shared memory for the
//     whole kernel; reference tiles (128 points) stream through a ring of 16-coordinate stages,
//     one bulk copy each (the global copy is stored directly in the UMMA "K-major, no swizzle"
//     core-matrix layout: 8 rows x 16 bytes per core matrix, LBO = 2048 B between K-adjacent
//     core matrices, SBO = 128 B between 8-row groups);
//   * the 128 x 128 FP32 accumulator lives in TMEM, double buffered (2 x 128 columns) so that
//     the MMAs of tile t+1 run while the four epilogue warps drain tile t with tcgen05.ld
//     (32 lanes x 32 columns per instruction; lane = query row, so a thread owns one row and
//     its candidate list);
//   * warp roles: warp 0 = producer (tile choice with 32 box tests per step + bulk copies),
//     warp 1 = MMA issuer (one elected lane) and TMEM allocator, warps 2-5 = epilogue.
// Supports d <= 128 (the resident operand and the ring must fit 227 KB); larger d uses the
// SIMT kernel.
#include "common.cuh"

namespace pb200 {

constexpr int TC_BM = 128;        // query rows per CTA (TMEM lanes)
constexpr int TC_BN = 128;        // reference points per tile (TMEM columns per accumulator)
constexpr int TC_KC = 16;         // coordinates per ring stage (two MMA k-steps)
constexpr int TC_HALF = TC_BN * TC_KC * 4;      // bytes of one hi (or lo) block of a stage: 8 KB
constexpr int TC_STAGE = 2 * TC_HALF;           // hi + lo: 16 KB
constexpr int TC_KSTEP = TC_BN * 8 * 4;         // bytes of one k-step (8 coordinates) inside a block: 4 KB
constexpr int TC_CAP = 64;
constexpr int TC_HIGH = 56;
constexpr int TC_MAX_KEEP = 48;
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
  int k_keep;
  unsigned long long* lists;
  int* out_idx;
  float* out_score;
  float* out_thr;
  unsigned long long* tiles_visited;
};

struct TcCtl {
  uint64_t a_full;
  uint64_t b_full[8], b_empty[8];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t yn_full[2];
  int meta_tile[8], meta_kc[8];
  int acc_tile[2];
  int cnt[TC_BM];
  float thr[TC_BM];
  float xcn[TC_BM];
  float warp_bound[4];
  float yn[2][TC_BN];
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

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ int tc_tile_of(int s, int q, int T) {
  const int R = T - q, L = q;
  const int m = R < L ? R : L;
  if (s < 2 * m) return (s & 1) ? q - 1 - (s >> 1) : q + (s >> 1);
  const int rest = s - 2 * m;
  return R > L ? q + m + rest : q - 1 - m - rest;
}

// warp-collective compaction of one row's candidate buffer (see knn_f32.cu)
__device__ __noinline__ void tc_compact_row(TcCtl* ctl, unsigned long long* buf, int row, int k_keep, int lane) {
  int c = ctl->cnt[row];
  c = c < TC_CAP ? c : TC_CAP;
  const unsigned long long e0 = lane < c ? __ldcg(buf + lane) : ~0ull;
  const unsigned long long e1 = lane + 32 < c ? __ldcg(buf + lane + 32) : ~0ull;
  int r0 = 0, r1 = 0;
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const unsigned long long o0 = __shfl_sync(0xffffffffu, e0, j);
    const unsigned long long o1 = __shfl_sync(0xffffffffu, e1, j);
    r0 += (o0 < e0) + (o1 < e0);
    r1 += (o0 < e1) + (o1 < e1);
  }
  __syncwarp();
  if (lane < c && r0 < k_keep) __stcg(buf + r0, e0);
  if (lane + 32 < c && r1 < k_keep) __stcg(buf + r1, e1);
  if (c >= k_keep) {
    if (lane < c && r0 == k_keep - 1) ctl->thr[row] = f32_from_ordered((uint32_t)(e0 >> 32));
    if (lane + 32 < c && r1 == k_keep - 1) ctl->thr[row] = f32_from_ordered((uint32_t)(e1 >> 32));
  }
  if (lane == 0) ctl->cnt[row] = c < k_keep ? c : k_keep;
  __threadfence_block();
  __syncwarp();
}

__global__ void __launch_bounds__(TC_THREADS, 1) knn_candidates_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [A resident: kch stages][B ring: nst stages][control]
  unsigned char* a_smem = smem_raw;
  unsigned char* b_smem = smem_raw + (size_t)p.kch * TC_STAGE;
  TcCtl& ctl = *reinterpret_cast<TcCtl*>(b_smem + (size_t)p.nst * TC_STAGE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qbase = p.q0 + blockIdx.x * TC_BM;
  const int row_local0 = blockIdx.x * TC_BM;
  const int qtile = qbase / TC_BN;
  const int KCH = p.kch, NST = p.nst, T = p.n_tiles;
  unsigned long long* lists = p.lists + (long long)row_local0 * TC_CAP;

  for (int r = tid; r < TC_BM; r += TC_THREADS) {
    ctl.cnt[r] = 0;
    ctl.thr[r] = INFINITY;
    ctl.xcn[r] = (row_local0 + r < p.nq) ? __double2float_ru(p.xcn[qbase + r]) : -INFINITY;
  }
  if (tid < 4) ctl.warp_bound[tid] = (row_local0 + tid * 32 < p.nq) ? INFINITY : -INFINITY;
  if (tid == 0) {
    mbar_init(&ctl.a_full, 1);
    for (int s = 0; s < NST; ++s) { mbar_init(&ctl.b_full[s], 1); mbar_init(&ctl.b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&ctl.acc_full[b], 1); mbar_init(&ctl.acc_empty[b], 4); mbar_init(&ctl.yn_full[b], 1); }
    fence_mbar_init();
    if (p.box_lo) {
      for (int a = 0; a < 3; ++a) { ctl.qlo[a] = p.box_lo[(long long)qtile * 3 + a]; ctl.qhi[a] = p.box_hi[(long long)qtile * 3 + a]; }
    }
  }
  if (warp == 1) {
    // TMEM: 2 accumulators x 128 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;
  const uint32_t last_bytes = (uint32_t)p.tail_steps * TC_KSTEP;   // bytes of a hi (or lo) block in the last stage

  if (warp == 0) {
    // ================= producer =================
    if (lane == 0) {
      // resident query operand
      uint32_t tx = 0;
      for (int kc = 0; kc < KCH; ++kc) tx += (kc == KCH - 1) ? 2 * last_bytes : (uint32_t)TC_STAGE;
      mbar_arrive_expect_tx(&ctl.a_full, tx);
      for (int kc = 0; kc < KCH; ++kc) {
        const float* src = p.xtc + ((long long)qtile * KCH + kc) * (TC_STAGE / 4);
        if (kc < KCH - 1 || p.tail_steps == 2) {
          bulk_g2s(a_smem + (size_t)kc * TC_STAGE, src, TC_STAGE, &ctl.a_full);
        } else {
          bulk_g2s(a_smem + (size_t)kc * TC_STAGE, src, last_bytes, &ctl.a_full);
          bulk_g2s(a_smem + (size_t)kc * TC_STAGE + TC_HALF, src + TC_HALF / 4, last_bytes, &ctl.a_full);
        }
      }
    }
    int cursor = 0, g = 0;
    unsigned long long visited = 0;
    for (;;) {
      // ---- choose the next tile (warp-collective box tests) ----
      float bmax = *(volatile float*)&ctl.warp_bound[0];
      bmax = fmaxf(bmax, *(volatile float*)&ctl.warp_bound[1]);
      bmax = fmaxf(bmax, *(volatile float*)&ctl.warp_bound[2]);
      bmax = fmaxf(bmax, *(volatile float*)&ctl.warp_bound[3]);
      const double bound = (double)bmax;
      int chosen = -1;
      while (cursor < T) {
        const int sidx = cursor + lane;
        const int tile = sidx < T ? tc_tile_of(sidx, qtile, T) : -1;
        bool ok = false;
        if (tile >= 0) {
          if (!p.box_lo) ok = true;
          else {
            double g2 = 0.0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const double d1 = p.box_lo[(long long)tile * 3 + a] - ctl.qhi[a];
              const double d2 = ctl.qlo[a] - p.box_hi[(long long)tile * 3 + a];
              double gp = d1 > d2 ? d1 : d2;
              gp = gp > 0.0 ? gp : 0.0;
              g2 = fma(gp, gp, g2);
            }
            ok = !(g2 * (1.0 - 1e-9) > bound);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) {
          const int first = __ffs(m) - 1;
          chosen = __shfl_sync(0xffffffffu, tile, first);
          cursor += first + 1;
          break;
        }
        cursor += 32;
      }
      if (chosen < 0) {
        // end of stream
        const int st = g % NST;
        if (lane == 0) {
          mbar_wait(&ctl.b_empty[st], (((uint32_t)(g / NST)) & 1u) ^ 1u);
          ctl.meta_tile[st] = -1;
          mbar_arrive(&ctl.b_full[st]);
        }
        break;
      }
      ++visited;
      if (lane == 0) {
        for (int kc = 0; kc < KCH; ++kc, ++g) {
          const int st = g % NST;
          mbar_wait(&ctl.b_empty[st], (((uint32_t)(g / NST)) & 1u) ^ 1u);
          ctl.meta_tile[st] = chosen;
          ctl.meta_kc[st] = kc;
          const float* src = p.xtc + ((long long)chosen * KCH + kc) * (TC_STAGE / 4);
          unsigned char* dst = b_smem + (size_t)st * TC_STAGE;
          if (kc < KCH - 1 || p.tail_steps == 2) {
            mbar_arrive_expect_tx(&ctl.b_full[st], TC_STAGE);
            bulk_g2s(dst, src, TC_STAGE, &ctl.b_full[st]);
          } else {
            mbar_arrive_expect_tx(&ctl.b_full[st], 2 * last_bytes);
            bulk_g2s(dst, src, last_bytes, &ctl.b_full[st]);
            bulk_g2s(dst + TC_HALF, src + TC_HALF / 4, last_bytes, &ctl.b_full[st]);
          }
        }
      }
      g = __shfl_sync(0xffffffffu, g, 0);
    }
    if (lane == 0 && p.tiles_visited) atomicAdd(p.tiles_visited, visited);
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      mbar_wait(&ctl.a_full, 0);
      tc_fence_after();
      const uint64_t a_desc0 = tc_smem_desc(smem_u32(a_smem)), b_desc0 = tc_smem_desc(smem_u32(b_smem));
      int tcount = 0;
      for (int g = 0;; ++g) {
        const int st = g % NST;
        mbar_wait(&ctl.b_full[st], ((uint32_t)(g / NST)) & 1u);
        tc_fence_after();
        const int tile = ctl.meta_tile[st];
        const int buf = tcount & 1;
        if (tile < 0) {
          // end of stream: tell the epilogue (no MMA pending on this buffer: plain arrive)
          mbar_wait(&ctl.acc_empty[buf], (((uint32_t)(tcount >> 1)) & 1u) ^ 1u);
          ctl.acc_tile[buf] = -1;
          mbar_arrive(&ctl.acc_full[buf]);
          break;
        }
        const int kc = ctl.meta_kc[st];
        if (kc == 0) {
          mbar_wait(&ctl.acc_empty[buf], (((uint32_t)(tcount >> 1)) & 1u) ^ 1u);
          tc_fence_after();
          // reference norms of this tile for the epilogue
          mbar_arrive_expect_tx(&ctl.yn_full[buf], TC_BN * 4);
          bulk_g2s(&ctl.yn[buf][0], p.yn + (long long)tile * TC_BN, TC_BN * 4, &ctl.yn_full[buf]);
        }
        const uint32_t d_tmem = tmem + (uint32_t)buf * TC_BN;
        // descriptors differ only in the 14-bit start-address field: add (byte offset >> 4) to a base
        const uint64_t a_hi = a_desc0 + (uint64_t)((uint32_t)kc * (TC_STAGE >> 4));
        const uint64_t b_hi = b_desc0 + (uint64_t)((uint32_t)st * (TC_STAGE >> 4));
        tc_mma_tf32(d_tmem, a_hi, b_hi, idesc, kc != 0 ? 1u : 0u);
        tc_mma_tf32(d_tmem, a_hi, b_hi + (TC_HALF >> 4), idesc, 1u);
        tc_mma_tf32(d_tmem, a_hi + (TC_HALF >> 4), b_hi, idesc, 1u);
        if (kc < KCH - 1 || p.tail_steps == 2) {
          tc_mma_tf32(d_tmem, a_hi + (TC_KSTEP >> 4), b_hi + (TC_KSTEP >> 4), idesc, 1u);
          tc_mma_tf32(d_tmem, a_hi + (TC_KSTEP >> 4), b_hi + ((TC_KSTEP + TC_HALF) >> 4), idesc, 1u);
          tc_mma_tf32(d_tmem, a_hi + ((TC_KSTEP + TC_HALF) >> 4), b_hi + (TC_KSTEP >> 4), idesc, 1u);
        }
        tc_commit(&ctl.b_empty[st]);               // stage reusable once these MMAs have read it
        if (kc == KCH - 1) {
          ctl.acc_tile[buf] = tile;
          tc_commit(&ctl.acc_full[buf]);           // accumulator complete
          ++tcount;
        }
      }
    }
  } else {
    // ================= epilogue (warps 2..5): TMEM lane quarter = warp % 4 =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // CTA-local query row owned by this thread
    unsigned long long* mybuf = lists + (long long)row * TC_CAP;
    bool compacted_any = false;
    for (int t = 0;; ++t) {
      const int buf = t & 1;
      mbar_wait(&ctl.acc_full[buf], ((uint32_t)(t >> 1)) & 1u);
      tc_fence_after();
      const int tile = ctl.acc_tile[buf];
      if (tile < 0) break;
      mbar_wait(&ctl.yn_full[buf], ((uint32_t)(t >> 1)) & 1u);
      const int colbase = tile * TC_BN;
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t v[32];
        tc_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TC_BN + cb * 32), v);
        float thr = ctl.thr[row];
        float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 y = *reinterpret_cast<const float4*>(&ctl.yn[buf][cb * 32 + j4 * 4]);
          const float s0 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 0]), y.x);
          const float s1 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 1]), y.y);
          const float s2 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 2]), y.z);
          const float s3 = fmaf(-2.f, __uint_as_float(v[j4 * 4 + 3]), y.w);
          v[j4 * 4 + 0] = __float_as_uint(s0); v[j4 * 4 + 1] = __float_as_uint(s1);
          v[j4 * 4 + 2] = __float_as_uint(s2); v[j4 * 4 + 3] = __float_as_uint(s3);
          m0 = fminf(m0, s0); m1 = fminf(m1, s1); m2 = fminf(m2, s2); m3 = fminf(m3, s3);
        }
        const float m = fminf(fminf(m0, m1), fminf(m2, m3));
        if (__any_sync(0xffffffffu, m < thr)) {
          unsigned int pending = 0xffffffffu;
          for (;;) {
            thr = ctl.thr[row];
            int c = ctl.cnt[row];
            bool fail = false;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = __uint_as_float(v[j]);
              if ((pending >> j & 1u) && s < thr) {
                if (c < TC_CAP) {
                  __stcg(mybuf + c, tc_pack(s, colbase + cb * 32 + j));
                  ++c;
                  pending &= ~(1u << j);
                } else {
                  fail = true;
                }
              }
            }
            ctl.cnt[row] = c;
            __threadfence_block();
            __syncwarp();
            const unsigned need = __ballot_sync(0xffffffffu, c >= TC_HIGH);
            for (unsigned mm = need; mm; mm &= mm - 1) {
              const int l = __ffs(mm) - 1;
              tc_compact_row(&ctl, lists + (long long)(quarter * 32 + l) * TC_CAP, quarter * 32 + l, p.k_keep, lane);
              compacted_any = true;
            }
            if (!__any_sync(0xffffffffu, fail)) break;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.acc_empty[buf]);
      if (__any_sync(0xffffffffu, compacted_any)) {
        float b = ctl.xcn[row] != -INFINITY ? __fadd_ru(ctl.thr[row], ctl.xcn[row]) : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
        if (lane == 0) ctl.warp_bound[quarter] = b;
        compacted_any = false;
      }
    }
    // ---- emit: final sort of each row of this warp ----
    for (int l = 0; l < 32; ++l) {
      const int r = quarter * 32 + l;
      const int grow = row_local0 + r;
      if (grow >= p.nq) continue;
      int c = ctl.cnt[r];
      c = c < TC_CAP ? c : TC_CAP;
      const unsigned long long* buf = lists + (long long)r * TC_CAP;
      const unsigned long long e0 = lane < c ? __ldcg(buf + lane) : ~0ull;
      const unsigned long long e1 = lane + 32 < c ? __ldcg(buf + lane + 32) : ~0ull;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
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

// Tensor-core version of pb200_knn_candidates_f32 (same outputs; q0 % 128 == 0; lists: ceil(nq/128)*128*64 words).
int pb200_knn_candidates_tc(const float* xtc, long long n_pad, const float* yn, const double* xcn,
                            const double* box_lo, const double* box_hi, int q0, int nq, int d, int k_keep,
                            unsigned long long* lists, int* out_idx, float* out_score, float* out_thr,
                            unsigned long long* tiles_visited, void* stream) {
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
  const int avail = 227 * 1024 - (int)sizeof(TcCtl) - 1024 - p.kch * TC_STAGE;
  int nst = avail / TC_STAGE;
  if (nst > 8) nst = 8;
  if (nst < 2) { set_error_msg("pb200_knn_candidates_tc: not enough shared memory for the ring"); return -1; }
  p.nst = nst;
  const int smem = (p.kch + nst) * TC_STAGE + (int)sizeof(TcCtl) + 64;
  PB_CHECK(cudaFuncSetAttribute(knn_candidates_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  knn_candidates_tc_kernel<<<div_up(nq, TC_BM), TC_THREADS, smem, st>>>(p);
  PB_CHECK_LAUNCH("knn_candidates_tc_kernel");
  return 0;
}

}  // extern "C"

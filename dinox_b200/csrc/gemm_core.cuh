// Persistent, warp-specialised tcgen05 GEMM core for sm_100a:
//   warp 0      : TMA producer (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulators in TMEM)
//   warps 2..   : epilogue (tcgen05.ld 32x32b -> registers -> fused row math -> global)
// D[M,N] = A[M,K] * B[N,K]^T with bf16 operands and fp32 accumulation; either operand may be
// K-major (reduction dim contiguous) or MN-major (M/N contiguous) in global memory.
// NSUB = 2 runs two independent GEMMs (own A and B) into adjacent TMEM accumulators of the same
// tile so that an epilogue can combine them element-wise (student/teacher logit tiles).
// Accumulators are double-buffered in TMEM: the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "sm100.cuh"

// Pair mode, "accumulator stage drained" hand-off of the non-leader CTA: 0 = its epilogue warps arrive on
// the leader's barrier directly with a relaxed remote arrive (TMEM reads have completed: nothing to
// publish); 1 = they arrive on a local barrier and one forwarder thread does a release.cluster arrive
// (formally conservative, ~1.6k cycles of extra latency per tile).
#ifndef DINOX_PAIR_FORWARDER
#define DINOX_PAIR_FORWARDER 0
#endif
#ifndef DINOX_EXP_NO_TMA
#define DINOX_EXP_NO_TMA 0
#endif
#ifndef DINOX_EXP_NO_MMA
#define DINOX_EXP_NO_MMA 0
#endif
#ifndef DINOX_EXP_NO_EPI
#define DINOX_EXP_NO_EPI 0
#endif

namespace dinox {
namespace gemm {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
// Warp roles: epilogue warps FIRST (0 .. kEpiWarps-1), then the TMA producer and the MMA issuer as the
// two highest warp ids.  The SM sub-partition arbiter prefers the highest warp id among eligible warps,
// so the two single-instruction-stream control warps are never starved by the busy epilogue warps that
// share their sub-partitions (with the control warps as warps 0/1 the MMA issue stalled behind the math).
#ifndef DINOX_CTRL_WARPS_LAST
#define DINOX_CTRL_WARPS_LAST 1
#endif

// A-operand collector reuse across the N sub-tiles of a 384-wide tile (see sm100::umma_bf16_collect)
#ifndef DINOX_COLLECTOR
#define DINOX_COLLECTOR 1
#endif

struct TileCoord {
  int m_tile, n_tile, batch, split;
  int row_shift = 0;   // sharded output only: rows of the problem that precede the owner's shard (m_tile is shard-local)
  int kpart = 0, kparts = 1;   // ordered split: this item reduces part kpart of kparts of the K range into the SAME output
};

struct CoreParams {
  int M, N, K;                 // K = reduction length
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int a_mn_major, b_mn_major;  // operand layouts in global memory
  int m_fastest;               // tile order: consecutive CTAs walk M (1) or N (0)
  int batches;                 // > 1: independent problems along a third tensor-map dimension
  int splits;                  // > 1: split-K, split s reduces k-blocks [s*kb_per_split, ...) into output slab s
  int kb_per_split;
  // Ordered split (os_parts >= 2; splits == batches == 1): super tiles [0, os_first) are whole items; each of the
  // remaining ones is cut into os_parts K ranges of os_kb k-blocks that accumulate into the same output IN PART ORDER
  // (the epilogue of part j waits for part j - 1 of its tile, see EpiStore), so the result does not depend on timing.
  // Items are numbered whole tiles first, then part-major: every item only depends on items with a smaller number,
  // which the round-robin walk has already started - no deadlock on a resident (persistent) grid.  Used to fill
  // the last, partial wave of a GEMM whose tile count is not a multiple of the cluster count.
  int os_first = 0, os_parts = 0, os_kb = 0;
  int rows_per_owner;          // > 0: the output rows are sharded over several buffers (one tensor map each, e.g. the
                               // peer-mapped gradient shards of the data-parallel ranks): M tile t goes to map
                               // t*BM / rows_per_owner at local row t*BM % rows_per_owner.  Multiple of BM.  0 = one map.
};

constexpr int kResKBlocks = 6;   // resident-A mode holds up to 6 k-blocks (K <= 384) of this CTA's 128 A rows

// Resident-operand modes (RES): an operand that consecutive tiles of a CTA share stays in shared memory
// (all k-blocks, K <= 384) and only the other operand streams through the stage ring.
//   kResA : A rows of sub-GEMM 0, tiles walked N-fastest in contiguous runs (pass 1: the H rows of an M tile)
//   kResB : B rows of EVERY sub-GEMM, tiles walked in M-columns (pass 2: the entries of an N tile; a CTA of
//           a pair holds 1/CL of them, so student + teacher fit: 2 x 6 x 8 KB)
enum : int { kResNone = 0, kResA = 1, kResB = 2 };
template <int BN, int CL, int EPI_SMEM, int RES = kResNone, int NSUB = 1>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;  // 16 KB
  static constexpr int kBBytes = (BN / CL) * BK * 2;   // a CTA of a pair holds 1/CL of the B tile
  static constexpr bool kStageHasA = !(RES == kResA && NSUB == 1);
  static constexpr bool kStageHasB = RES != kResB;
  static constexpr int kStageBytes = (kStageHasA ? kABytes : 0) + (kStageHasB ? kBBytes : 0);
  static constexpr int kResBytes = RES == kResA ? kResKBlocks * kABytes : RES == kResB ? NSUB * kResKBlocks * kBBytes : 0;
  static constexpr int kEpiBytes = (EPI_SMEM + 255) / 256 * 256;
  // 227 KB per CTA minus alignment slack (1 KB), control block (256 B), the epilogue's staging area and the resident operand
  static constexpr int kBudget = 227 * 1024 - 1024 - 256 - kEpiBytes - kResBytes;
#ifndef DINOX_MAX_STAGES
#define DINOX_MAX_STAGES 8
#endif
  static constexpr int kFit = kBudget / kStageBytes;
  static constexpr int kStages = kFit > DINOX_MAX_STAGES ? DINOX_MAX_STAGES : kFit;
  static_assert(kStages >= 2, "tile too large for the smem pipeline");
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static constexpr int kTotal = kResBytes + kPipeBytes + kEpiBytes + 256 + 1024;
};

// Walks the tile indices cid, cid + ncl, cid + 2 ncl, ... of a persistent CTA without a division per
// tile: the stride is decomposed once into (fast, slow, outer) digits and added with carries.
// Digit order: fast = M tiles (m_fastest) or N tiles, slow = the other, outer = batch / split.
struct TileWalker {
  int fast, slow, outer;
  int dfast, dslow, douter;
  int nfast, nslow;
  int remaining;
  bool m_fastest;
  __host__ __device__ __forceinline__ void init(const CoreParams& p, int num_m_super, int first, int stride, int total) {
    m_fastest = p.m_fastest != 0;
    nfast = m_fastest ? num_m_super : p.num_n_tiles;
    nslow = m_fastest ? p.num_n_tiles : num_m_super;
    fast = first % nfast; int t = first / nfast; slow = t % nslow; outer = t / nslow;
    dfast = stride % nfast; t = stride / nfast; dslow = t % nslow; douter = t / nslow;
    remaining = first < total ? (total - first + stride - 1) / stride : 0;
  }
  // contiguous range [first, first + count) with unit stride (consecutive tiles share the slow digit)
  __host__ __device__ __forceinline__ void init_range(const CoreParams& p, int num_m_super, int first, int count) {
    m_fastest = p.m_fastest != 0;
    nfast = m_fastest ? num_m_super : p.num_n_tiles;
    nslow = m_fastest ? p.num_n_tiles : num_m_super;
    fast = first % nfast; int t = first / nfast; slow = t % nslow; outer = t / nslow;
    dfast = 1; dslow = 0; douter = 0;
    remaining = count > 0 ? count : 0;
  }
  // Column schedule (resident B): every cluster sweeps the M tiles of ONE N tile at a time, all clusters in
  // step (they read the same A tile at the same time -> it is fetched from HBM once).  Whole rounds of `ncl`
  // columns first; of the `rem` left-over columns each goes to one cluster, which gives up its last `tail`
  // M tiles to the ncl - rem otherwise idle clusters so that everybody finishes together.
  bool columns = false;
  int cm, cn, m_lo, m_hi, n_step, rem_a;          // current position / phase A (whole rounds) bookkeeping
  int b_cm, b_cn, b_lo, b_hi, b_step;             // phase B (left-over columns) start state
  __host__ __device__ __forceinline__ void init_columns(const CoreParams& p, int num_m, int cid, int ncl) {
    columns = true;
    const int num_n = p.num_n_tiles;
    const int full = num_n / ncl, rem = num_n - full * ncl, n_base = full * ncl;
    rem_a = full * num_m;
    int count_b = 0;
    b_cm = b_cn = b_lo = 0; b_hi = num_m; b_step = 0;
    if (rem > 0) {
      const int spares = ncl - rem;
      const int tail = (int)(((long long)num_m * spares) / ncl);
      if (cid < rem) {
        b_cn = n_base + cid; b_cm = 0; b_lo = 0; b_hi = num_m; b_step = 0; count_b = num_m - tail;
      } else if (tail > 0) {
        const int total = rem * tail, per = (total + spares - 1) / spares, j0 = (cid - rem) * per;
        count_b = per < total - j0 ? per : total - j0;
        if (count_b < 0) count_b = 0;
        b_cn = n_base + j0 / tail; b_lo = num_m - tail; b_cm = b_lo + j0 % tail; b_hi = num_m; b_step = 1;
      }
    }
    remaining = rem_a + count_b;
    if (rem_a > 0) { cn = cid; cm = 0; m_lo = 0; m_hi = num_m; n_step = ncl; }
    else { cn = b_cn; cm = b_cm; m_lo = b_lo; m_hi = b_hi; n_step = b_step; }
  }
  // Ordered split (CoreParams::os_*): item index o_i walks cid, cid + ncl, ...; a division per item is nothing
  // next to the >= 10 us an item of these GEMMs takes.
  bool ordered = false;
  int o_i = 0, o_stride = 0, o_first = 0, o_left = 1, o_parts = 1;
  __host__ __device__ __forceinline__ void init_ordered(const CoreParams& p, int num_m_super, int cid, int ncl) {
    ordered = true;
    m_fastest = p.m_fastest != 0;
    nfast = m_fastest ? num_m_super : p.num_n_tiles;
    nslow = m_fastest ? p.num_n_tiles : num_m_super;
    const int tiles = num_m_super * p.num_n_tiles;
    o_first = p.os_first; o_left = tiles - p.os_first > 0 ? tiles - p.os_first : 1; o_parts = p.os_parts;
    const int total = p.os_first + (tiles - p.os_first) * p.os_parts;
    o_i = cid; o_stride = ncl;
    remaining = cid < total ? (total - cid + ncl - 1) / ncl : 0;
  }
  __host__ __device__ __forceinline__ bool valid() const { return remaining > 0; }
  __host__ __device__ __forceinline__ void next() {
    --remaining;
    if (ordered) { o_i += o_stride; return; }
    if (columns) {
      if (rem_a > 0 && --rem_a == 0) { cn = b_cn; cm = b_cm; m_lo = b_lo; m_hi = b_hi; n_step = b_step; return; }
      if (++cm == m_hi) { cm = m_lo; cn += n_step; }
      return;
    }
    fast += dfast;
    int carry = 0;
    if (fast >= nfast) { fast -= nfast; carry = 1; }
    slow += dslow + carry;
    carry = 0;
    if (slow >= nslow) { slow -= nslow; carry = 1; }
    outer += douter + carry;
  }
  // CL-wide super tile -> this CTA's tile
  __host__ __device__ __forceinline__ TileCoord coord(const CoreParams& p, int cl, int crank) const {
    TileCoord c;
    if (ordered) {
      int t = o_i;
      if (o_i >= o_first) {
        const int j = o_i - o_first;
        c.kpart = j / o_left; c.kparts = o_parts; t = o_first + j - c.kpart * o_left;
      }
      const int s_ = t / nfast, f_ = t - s_ * nfast;
      c.m_tile = (m_fastest ? f_ : s_) * cl + crank;
      c.n_tile = m_fastest ? s_ : f_;
      c.batch = 0; c.split = 0;
      return c;
    }
    if (columns) { c.m_tile = cm * cl + crank; c.n_tile = cn; c.batch = 0; c.split = 0; return c; }
    c.m_tile = (m_fastest ? fast : slow) * cl + crank;
    c.n_tile = m_fastest ? slow : fast;
    c.batch = p.splits > 1 ? 0 : outer;
    c.split = p.splits > 1 ? outer : 0;
    return c;
  }
};

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  template <int kStages> __device__ __forceinline__ void advance() {
    if (++stage == kStages) { stage = 0; phase ^= 1; }
  }
};

struct SharedCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t tmem_empty_local[2];   // pair mode, non-leader CTA: its own epilogue warps report here
  uint64_t res_full, res_empty;       // resident operand: "tile landed" / "last MMA reading it retired"
  uint32_t tmem_base;
  uint32_t pad;
};

// fast exp2 (MUFU.EX2), inputs far below -126 flush to 0
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The kernel body shared by all GEMM flavours.
//   BN     : tile width (N);  NSPLIT : MMA instructions along N per k-step sharing the A tile
//            (BN/NSPLIT <= 256 is the UMMA N);  NSUB : independent sub-GEMMs per tile
//   CL     : 1 = one CTA per 128 x BN tile (tcgen05 cta_group::1)
//            2 = CTA pair (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile over two SMs.  Each
//                CTA loads its own 128 A rows and HALF of the B tile; the leader CTA issues the MMAs
//                for both; each CTA's epilogue drains the 128 rows that live in its own TMEM.  This
//                halves the B-operand bytes every SM pulls from L2 per FLOP - the K<=1024 GEMMs of
//                the loss head are bound by L2->SMEM operand traffic, not by the tensor pipe.
// `Epi` provides kEpiWarps, kEpiSmemBytes, Params, State, fetch(), prologue(), tile(), finish().
// `tmC` is the output tensor map of epilogues that store through TMA (others ignore it).
template <int BN, int NSPLIT, int NSUB, int CL, class Epi, int RES = kResNone>
__device__ __forceinline__ void gemm_body(const CoreParams& p, const typename Epi::Params& ep,
                                          const CUtensorMap* tmA0, const CUtensorMap* tmB0,
                                          const CUtensorMap* tmA1, const CUtensorMap* tmB1,
                                          const CUtensorMap* tmC, uint8_t* smem_raw) {
  using L = SmemLayout<BN, CL, Epi::kEpiSmemBytes, RES, NSUB>;
  constexpr bool RESA = (RES == kResA), RESB = (RES == kResB);
  static_assert(!RESB || NSPLIT == 1, "resident B: one MMA per k-step");
  constexpr int kStages = L::kStages;
  constexpr int BNI = BN / NSPLIT;                           // UMMA N
  constexpr int BNL = BNI / CL;                              // rows of one N sub-tile held by this CTA
  constexpr int kAccCols = NSUB * BN;                        // TMEM columns per accumulator stage
  constexpr int kAccStages = (2 * kAccCols <= 512) ? 2 : 1;  // double-buffer when it fits
  constexpr uint32_t kNeed = kAccStages * kAccCols;
  constexpr uint32_t kTmemCols = kNeed <= 32 ? 32 : kNeed <= 64 ? 64 : kNeed <= 128 ? 128 : kNeed <= 256 ? 256 : 512;
  static_assert(kAccCols <= 512, "accumulators do not fit TMEM");
  static_assert(BNI % 16 == 0 && BNI <= 256 && BN % 64 == 0, "invalid tile shape");
  static_assert(CL == 1 || CL == 2, "cluster size 1 or 2");
  static_assert(BNL % 8 == 0, "per-CTA B slice must be whole 8-row swizzle groups");
  constexpr bool kPair = (CL == 2);

  // 1024-B aligned carve-up (swizzle-128B atoms need it): [pipeline stages][epilogue staging][control]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res_buf = smem;                       // resident operand: A k-blocks (kResA) or [sub][k-block] B slabs (kResB)
  uint8_t* pipe = smem + L::kResBytes;
  uint8_t* epi_smem = pipe + L::kPipeBytes;
  SharedCtl* ctl = reinterpret_cast<SharedCtl*>(pipe + L::kPipeBytes + L::kEpiBytes);

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // logical role index: 0 = producer, 1 = MMA issuer, 2.. = epilogue
#if DINOX_CTRL_WARPS_LAST
  const int warp = warp_id >= Epi::kEpiWarps ? warp_id - Epi::kEpiWarps : warp_id + 2;
#else
  const int warp = warp_id;
#endif
  const int crank = kPair ? (int)sm100::cluster_ctarank() : 0;
  const bool leader = (crank == 0);
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  // pairs walk "super tiles" of CL adjacent M tiles; a ragged last M super tile computes on
  // zero-filled (out-of-bounds) rows and the epilogues mask it
  const int num_m_super = (p.num_m_tiles + CL - 1) / CL;
  const int num_super = num_m_super * p.num_n_tiles * (p.splits > 1 ? p.splits : p.batches);

  if (warp == 0 && lane == 0) {
    sm100::prefetch_tmap(tmA0);
    sm100::prefetch_tmap(tmB0);
    if (NSUB == 2) { sm100::prefetch_tmap(tmA1); sm100::prefetch_tmap(tmB1); }
    if (Epi::kUsesTmaStore) sm100::prefetch_tmap(tmC);
    for (int i = 0; i < kStages; ++i) { sm100::mbar_init(&ctl->full[i], 1); sm100::mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      sm100::mbar_init(&ctl->tmem_full[i], 1);
      // pair: the leader's barrier takes its own epilogue warps plus ONE forwarded arrival for the peer
      // CTA, whose epilogue warps report to their local barrier (see the forwarder in warp 1)
#if DINOX_PAIR_FORWARDER
      sm100::mbar_init(&ctl->tmem_empty[i], Epi::kEpiWarps + (kPair ? 1 : 0));
#else
      sm100::mbar_init(&ctl->tmem_empty[i], CL * Epi::kEpiWarps);   // both CTAs' epilogue warps arrive directly
#endif
      sm100::mbar_init(&ctl->tmem_empty_local[i], Epi::kEpiWarps);
    }
    sm100::mbar_init(&ctl->res_full, 1);
    sm100::mbar_init(&ctl->res_empty, 1);
    sm100::fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) { sm100::tmem_alloc_pair(&ctl->tmem_base, kTmemCols); sm100::tmem_relinquish_pair(); }
    else { sm100::tmem_alloc(&ctl->tmem_base, kTmemCols); sm100::tmem_relinquish(); }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (kPair) sm100::cluster_sync();   // the peer's barriers / TMEM are set up before anyone signals them
  sm100::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  // barriers, TMEM and descriptor prefetch are set up: only now wait for the producer kernels of our inputs
  // (the launch may have been programmatic, DINOX_PDL)
  sm100::grid_dependency_wait();

  TileWalker walk;
  if (RESA) {   // contiguous chunk per cluster: consecutive tiles share the M tile whose A rows stay resident
    const int per = (num_super + ncl - 1) / ncl;
    const int first = cid * per;
    walk.init_range(p, num_m_super, first, min(per, num_super - first));
  } else if (RESB) {
    walk.init_columns(p, num_m_super, cid, ncl);
  } else if (p.os_parts > 1) {
    walk.init_ordered(p, num_m_super, cid, ncl);
  } else {
    walk.init(p, num_m_super, cid, ncl, num_super);
  }

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and its share of B) =====
    // The WHOLE warp walks the loop with warp-uniform state and one elected lane issues: inside a
    // divergent `if (lane == 0)` region nvcc wraps every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST
    // + BRA.U.ANY "waterfall" (~20 dependent instructions per issue), which made the single issuing
    // thread - not the tensor pipe - the limiter of the 64-cycle N=128 MMAs.
    {
      PipeState st;
      int res_tile = -1, res_batch = -1;     // M tile (kResA) / N tile (kResB) whose rows are resident
      uint32_t res_phase = 0;
      for (; walk.valid(); walk.next()) {
        const TileCoord tc = walk.coord(p, CL, crank);
        const int m0 = tc.m_tile * BM, n0 = tc.n_tile * BN;
        int kb0 = tc.split * p.kb_per_split;
        int kb1 = p.splits > 1 ? min(kb0 + p.kb_per_split, p.num_k_blocks) : p.num_k_blocks;
        if (tc.kparts > 1) { kb0 = tc.kpart * p.os_kb; kb1 = min(kb0 + p.os_kb, p.num_k_blocks); }
        if (RESA && (tc.m_tile != res_tile || tc.batch != res_batch)) {
          // (re)load the resident A rows of sub-GEMM 0: every k-block, once per M tile.  The previous
          // resident tile must have been read by its last MMA (res_empty, committed by the MMA issuer).
          sm100::mbar_wait(&ctl->res_empty, res_phase ^ 1, 8);
          if (sm100::elect_one()) {
            const uint32_t af = kPair ? sm100::mapa_u32(sm100::smem_u32(&ctl->res_full), 0) : 0;
            if (leader) sm100::mbar_expect_tx(&ctl->res_full, CL * p.num_k_blocks * L::kABytes);
            for (int kb = 0; kb < p.num_k_blocks; ++kb) {
              uint8_t* dst = res_buf + kb * L::kABytes;
              if (kPair) {
                if (p.batches > 1) sm100::tma_load_3d_pair(dst, tmA0, af, kb * BK, m0, tc.batch);
                else sm100::tma_load_2d_pair(dst, tmA0, af, kb * BK, m0);
              } else {
                if (p.batches > 1) sm100::tma_load_3d(dst, tmA0, &ctl->res_full, kb * BK, m0, tc.batch);
                else sm100::tma_load_2d(dst, tmA0, &ctl->res_full, kb * BK, m0);
              }
            }
          }
          __syncwarp();
          res_tile = tc.m_tile; res_batch = tc.batch; res_phase ^= 1;
        }
        if (RESB && tc.n_tile != res_tile) {
          // (re)load the resident B rows (this CTA's share of the N tile) of every sub-GEMM, all k-blocks
          sm100::mbar_wait(&ctl->res_empty, res_phase ^ 1, 8);
          if (sm100::elect_one()) {
            const uint32_t af = kPair ? sm100::mapa_u32(sm100::smem_u32(&ctl->res_full), 0) : 0;
            if (leader) sm100::mbar_expect_tx(&ctl->res_full, CL * NSUB * p.num_k_blocks * L::kBBytes);
            for (int sub = 0; sub < NSUB; ++sub) {
              const CUtensorMap* mb = sub ? tmB1 : tmB0;
              for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                uint8_t* dst = res_buf + (sub * kResKBlocks + kb) * L::kBBytes;
                if (kPair) sm100::tma_load_2d_pair(dst, mb, af, kb * BK, n0 + crank * BNL);
                else sm100::tma_load_2d(dst, mb, &ctl->res_full, kb * BK, n0 + crank * BNL);
              }
            }
          }
          __syncwarp();
          res_tile = tc.n_tile; res_phase ^= 1;
        }
        for (int sub = 0; sub < NSUB; ++sub) {
          const CUtensorMap* ma = sub ? tmA1 : tmA0;
          const CUtensorMap* mb = sub ? tmB1 : tmB0;
          const bool load_a = !(RESA && sub == 0);
          for (int kb = kb0; kb < kb1; ++kb) {
            sm100::mbar_wait(&ctl->empty[st.stage], st.phase ^ 1, 1);
            if (sm100::elect_one()) {
            uint8_t* sa = pipe + st.stage * L::kStageBytes;
            uint8_t* sb = sa + (L::kStageHasA ? L::kABytes : 0);
            uint64_t* full = &ctl->full[st.stage];
            // pair: all bytes (both CTAs) are counted on the leader's barrier.  (Counting per CTA and
            // forwarding "my half landed" with a release.cluster arrive was measured 2x slower: the
            // cluster-scope release costs ~1.6k cycles per stage on the forwarding thread.)
            constexpr bool kRemoteTx = kPair;
#if DINOX_EXP_NO_TMA   // experiment: no operand loads at all, the MMAs re-read whatever is in smem
            if (leader) sm100::mbar_arrive(full);
#else
            if (leader) sm100::mbar_expect_tx(full, CL * ((load_a ? L::kABytes : 0) + (RESB ? 0 : L::kBBytes)));
            const uint32_t full_addr = kRemoteTx ? sm100::mapa_u32(sm100::smem_u32(full), 0) : 0;
            const int k0 = kb * BK;
            const bool b3 = p.batches > 1;
            auto load = [&](uint8_t* dst, const CUtensorMap* tm, int c0, int c1) {
              if (kRemoteTx) {
                if (b3) sm100::tma_load_3d_pair(dst, tm, full_addr, c0, c1, tc.batch);
                else sm100::tma_load_2d_pair(dst, tm, full_addr, c0, c1);
              } else {
                if (b3) sm100::tma_load_3d(dst, tm, full, c0, c1, tc.batch);
                else sm100::tma_load_2d(dst, tm, full, c0, c1);
              }
            };
            // ---- A: this CTA's own 128 rows (not for the sub-GEMM whose A is resident)
            if (load_a) {
              if (!p.a_mn_major) {
                load(sa, ma, k0, m0);
              } else {
#pragma unroll
                for (int c = 0; c < BM / 64; ++c) load(sa + c * (BK * 128), ma, m0 + c * 64, k0);
              }
            }
            // ---- B: for every N sub-tile h, this CTA's BNL of its BNI rows
#pragma unroll
            for (int h = 0; h < (RESB ? 0 : NSPLIT); ++h) {
              const int r0 = n0 + h * BNI + crank * BNL;
              if (!p.b_mn_major) {
                load(sb + h * (BNL * 128), mb, k0, r0);
              } else {
                constexpr int kAtoms = BNL / 64;   // host guarantees BNL % 64 == 0 for MN-major B
#pragma unroll
                for (int c = 0; c < (kAtoms > 0 ? kAtoms : 1); ++c)
                  load(sb + (h * (kAtoms > 0 ? kAtoms : 1) + c) * (BK * 128), mb, r0 + c * 64, k0);
              }
            }
#endif
            }   // elected lane
            __syncwarp();
            st.advance<kStages>();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== pair, non-leader CTA: forward "accumulator stage drained" ==============
    // One release.cluster arrive per tile from a thread with no memory traffic of its own; the
    // epilogue warps only pay a CTA-local arrive.
    if (DINOX_PAIR_FORWARDER && kPair && !leader) {
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (; walk.valid(); walk.next()) {
        sm100::mbar_wait(&ctl->tmem_empty_local[acc_stage], acc_phase, 7);
        if (sm100::elect_one())
          sm100::mbar_arrive_cluster(sm100::mapa_u32(sm100::smem_u32(&ctl->tmem_empty[acc_stage]), 0));
        __syncwarp();
        if (kAccStages == 2) { acc_stage ^= 1; if (acc_stage == 0) acc_phase ^= 1; }
        else acc_phase ^= 1;
      }
    }
    // ===================== MMA issuer (one thread; pair: the leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = sm100::umma_idesc_bf16(BM * CL, BNI, p.a_mn_major, p.b_mn_major);
      // per-UMMA_K advance of the descriptor start address, and LBO per layout
      const uint32_t a_adv = p.a_mn_major ? (UMMA_K * 128) : (UMMA_K * 2);
      const uint32_t b_adv = p.b_mn_major ? (UMMA_K * 128) : (UMMA_K * 2);
      const uint32_t a_lbo = p.a_mn_major ? (BK * 128) : 16, b_lbo = p.b_mn_major ? (BK * 128) : 16;
      // byte offset of N sub-tile h inside this CTA's B stage buffer
      const uint32_t b_split = p.b_mn_major ? (BNL / 64) * (BK * 128) : BNL * 128;
      PipeState st;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      int res_tile = -1, res_batch = -1;
      uint32_t res_phase = 0;
      for (; walk.valid(); walk.next()) {
        const TileCoord tc = walk.coord(p, CL, crank);
        int kb0 = tc.split * p.kb_per_split;
        int kb1 = p.splits > 1 ? min(kb0 + p.kb_per_split, p.num_k_blocks) : p.num_k_blocks;
        if (tc.kparts > 1) { kb0 = tc.kpart * p.os_kb; kb1 = min(kb0 + p.os_kb, p.num_k_blocks); }
        bool res_last = false;    // is this the last tile that reads the current resident A?
        if (RESA) {
          if (tc.m_tile != res_tile || tc.batch != res_batch) {
            sm100::mbar_wait(&ctl->res_full, res_phase, 9);
            res_tile = tc.m_tile; res_batch = tc.batch; res_phase ^= 1;
          }
          TileWalker nxt = walk;
          nxt.next();
          const TileCoord tn = nxt.coord(p, CL, crank);
          res_last = !nxt.valid() || tn.m_tile != tc.m_tile || tn.batch != tc.batch;
        }
        if (RESB) {
          if (tc.n_tile != res_tile) {
            sm100::mbar_wait(&ctl->res_full, res_phase, 9);
            res_tile = tc.n_tile; res_phase ^= 1;
          }
          TileWalker nxt = walk;
          nxt.next();
          res_last = !nxt.valid() || nxt.coord(p, CL, crank).n_tile != tc.n_tile;
        }
        sm100::mbar_wait(&ctl->tmem_empty[acc_stage], acc_phase ^ 1, 2);
        sm100::tc_fence_after();
        for (int sub = 0; sub < NSUB; ++sub) {
          const uint32_t d_tmem = tmem_base + acc_stage * kAccCols + sub * BN;
          for (int kb = kb0; kb < kb1; ++kb) {
            sm100::mbar_wait(&ctl->full[st.stage], st.phase, 3);
            sm100::tc_fence_after();
            const uint32_t stage_base = sm100::smem_u32(pipe + st.stage * L::kStageBytes);
            const uint32_t sa = (RESA && sub == 0) ? sm100::smem_u32(res_buf + kb * L::kABytes) : stage_base;
            const uint32_t sb = RESB ? sm100::smem_u32(res_buf + (sub * kResKBlocks + kb) * L::kBBytes)
                                     : stage_base + (L::kStageHasA ? L::kABytes : 0);
            if (sm100::elect_one()) {
#if DINOX_EXP_NO_MMA   // experiment: operands stream through smem but nothing reads them
            sm100::mbar_arrive(&ctl->empty[st.stage]);
            if (kPair) sm100::mbar_arrive_cluster_relaxed(sm100::mapa_u32(sm100::smem_u32(&ctl->empty[st.stage]), 1));
#else
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = sm100::umma_smem_desc(sa + k * a_adv, a_lbo, 1024);
              const uint32_t acc = (uint32_t)((kb != kb0) | (k != 0));
              if (NSPLIT == 3 && DINOX_COLLECTOR) {
                // the three N sub-tiles multiply the same A slab: read it from shared memory once
                const uint64_t db0 = sm100::umma_smem_desc(sb + k * b_adv, b_lbo, 1024);
                const uint64_t db1 = sm100::umma_smem_desc(sb + b_split + k * b_adv, b_lbo, 1024);
                const uint64_t db2 = sm100::umma_smem_desc(sb + 2 * b_split + k * b_adv, b_lbo, 1024);
                sm100::umma_bf16_collect<1, kPair>(d_tmem, da, db0, idesc, acc);
                sm100::umma_bf16_collect<2, kPair>(d_tmem + BNI, da, db1, idesc, acc);
                sm100::umma_bf16_collect<3, kPair>(d_tmem + 2 * BNI, da, db2, idesc, acc);
              } else {
#pragma unroll
                for (int h = 0; h < NSPLIT; ++h) {
                  const uint64_t db = sm100::umma_smem_desc(sb + h * b_split + k * b_adv, b_lbo, 1024);
                  if (kPair) sm100::umma_bf16_pair(d_tmem + h * BNI, da, db, idesc, acc);
                  else sm100::umma_bf16(d_tmem + h * BNI, da, db, idesc, acc);
                }
              }
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if (kPair) sm100::umma_commit_pair(&ctl->empty[st.stage]);
            else sm100::umma_commit(&ctl->empty[st.stage]);
            // ... and the resident A tile after the last MMA of the last tile that reads it
            if (((RESA && sub == 0) || (RESB && sub == NSUB - 1)) && res_last && kb == kb1 - 1) {
              if (kPair) sm100::umma_commit_pair(&ctl->res_empty);
              else sm100::umma_commit(&ctl->res_empty);
            }
#endif
            }   // elected lane
            __syncwarp();
            st.advance<kStages>();
          }
        }
        // accumulators complete -> epilogue (of both CTAs)
        if (sm100::elect_one()) {
          if (kPair) sm100::umma_commit_pair(&ctl->tmem_full[acc_stage]);
          else sm100::umma_commit(&ctl->tmem_full[acc_stage]);
        }
        __syncwarp();
        if (kAccStages == 2) { acc_stage ^= 1; if (acc_stage == 0) acc_phase ^= 1; }
        else acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int epi_warp = warp - 2;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    typename Epi::State state;
    TileCoord tc = walk.coord(p, CL, crank);
    if (walk.valid()) Epi::fetch(ep, p, tc, epi_warp, lane, state);   // per-tile constants: global -> registers
    while (walk.valid()) {
      Epi::prologue(ep, p, tc, acc_stage, epi_warp, lane, epi_smem, state);   // registers -> smem (+ barrier)
      walk.next();
      const TileCoord tn = walk.coord(p, CL, crank);
      if (walk.valid()) Epi::fetch(ep, p, tn, epi_warp, lane, state);   // next tile's loads fly during this tile
      sm100::mbar_wait(&ctl->tmem_full[acc_stage], acc_phase, 4);
      sm100::tc_fence_after();
#if !DINOX_EXP_NO_EPI   // experiment knob: skip the epilogue math, keep the barrier protocol
      {
        // sharded output: pick the owner's tensor map and address the tile by its row inside that shard
        const CUtensorMap* tmCt = tmC;
        TileCoord tce = tc;
        if (p.rows_per_owner > 0) {
          const int owner = (tc.m_tile * BM) / p.rows_per_owner;
          tmCt = tmC + owner;
          tce.m_tile = tc.m_tile - owner * (p.rows_per_owner / BM);
          tce.row_shift = owner * p.rows_per_owner;
        }
        Epi::tile(ep, p, tce, tmCt, tmem_base + acc_stage * kAccCols, acc_stage, epi_warp, lane, epi_smem, state);
      }
#endif
      sm100::tc_fence_before();
      __syncwarp();
#if DINOX_PAIR_FORWARDER
      if (lane == 0) sm100::mbar_arrive(leader ? &ctl->tmem_empty[acc_stage] : &ctl->tmem_empty_local[acc_stage]);
#else
      if (lane == 0) {
        if (leader) sm100::mbar_arrive(&ctl->tmem_empty[acc_stage]);
        else sm100::mbar_arrive_cluster_relaxed(sm100::mapa_u32(sm100::smem_u32(&ctl->tmem_empty[acc_stage]), 0));
      }
#endif
      if (kAccStages == 2) { acc_stage ^= 1; if (acc_stage == 0) acc_phase ^= 1; }
      else acc_phase ^= 1;
      tc = tn;
    }
    Epi::finish(ep, p, epi_warp, lane, state);
  }

  sm100::tc_fence_before();
  __syncthreads();
  if (kPair) sm100::cluster_sync();   // nobody exits while the peer may still signal / read its smem
  if (warp == 1) {
    sm100::tc_fence_after();
    if (kPair) sm100::tmem_dealloc_pair(tmem_base, kTmemCols);
    else sm100::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int CL, class Epi, int RES = kResNone, int NSUB = 1>
constexpr int smem_bytes() {
  return SmemLayout<BN, CL, Epi::kEpiSmemBytes, RES, NSUB>::kTotal;
}

// lane quarter of TMEM this warp may read (hardware: warp_id % 4), and which column half it owns
__device__ __forceinline__ int epi_quarter() { return (threadIdx.x >> 5) & 3; }

// ---------------------------------------------------------------------------------------------------
// Per-warp output staging for TMA stores: 32 rows x 128 B, SWIZZLE_128B (1024-B aligned, 4 KB).
// Row r = lane; 16-byte piece j of the row lives at r*128 + ((j ^ (r & 7)) << 4): a quarter warp's
// STS.128 of one logical piece covers all 32 banks -> conflict free, and it is exactly the layout
// a SWIZZLE_128B tensor map reads.
// ---------------------------------------------------------------------------------------------------
struct WarpStage {
  uint32_t base;   // shared-space address of this warp's 4 KB buffer
  uint32_t row;    // base + lane*128
  uint32_t sw;     // (lane & 7) << 4
  __device__ __forceinline__ void init(uint8_t* buf, int lane) {
    base = sm100::smem_u32(buf);
    row = base + lane * 128;
    sw = (lane & 7) << 4;
  }
  __device__ __forceinline__ void put(int piece, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + ((piece << 4) ^ sw)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  }
};

}  // namespace gemm
}  // namespace dinox

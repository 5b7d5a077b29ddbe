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

namespace dinox {
namespace gemm {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int kFirstEpiWarp = 2;

struct TileCoord {
  int m_tile, n_tile, batch;
};

struct CoreParams {
  int M, N, K;                 // K = reduction length
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int a_mn_major, b_mn_major;  // operand layouts in global memory
  int m_fastest;               // tile order: consecutive CTAs walk M (1) or N (0)
  int batches;                 // > 1: independent problems along a third tensor-map dimension
};

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;  // 16 KB
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBudget = 196 * 1024;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kPipeBytes = kStages * kStageBytes;
};

__device__ __forceinline__ TileCoord tile_coord(const CoreParams& p, int t) {
  TileCoord c;
  const int per = p.num_m_tiles * p.num_n_tiles;
  c.batch = t / per;
  t -= c.batch * per;
  if (p.m_fastest) { c.m_tile = t % p.num_m_tiles; c.n_tile = t / p.num_m_tiles; }
  else { c.n_tile = t % p.num_n_tiles; c.m_tile = t / p.num_n_tiles; }
  return c;
}

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
  uint32_t tmem_base;
  uint32_t pad;
};

// fast exp2 (MUFU.EX2), inputs far below -126 flush to 0
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The kernel body shared by all GEMM flavours.  `Epi` provides:
//   static constexpr int kEpiWarps;                      // 4 or 8
//   struct Params;                                       // POD, passed by value
//   __device__ static void tile(const Params&, const CoreParams&, TileCoord, uint32_t tmem_acc,
//                               int acc_stage, int epi_warp, int lane, uint8_t* epi_smem);
//   static constexpr int kEpiSmemBytes;
template <int BN, int NSUB, class Epi>
__device__ __forceinline__ void gemm_body(const CoreParams& p, const typename Epi::Params& ep,
                                          const CUtensorMap* tmA0, const CUtensorMap* tmB0,
                                          const CUtensorMap* tmA1, const CUtensorMap* tmB1,
                                          uint8_t* smem_raw) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  constexpr int kAccCols = NSUB * BN;                       // TMEM columns per accumulator stage
  constexpr uint32_t kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64
                                 : (2 * kAccCols <= 128) ? 128 : (2 * kAccCols <= 256) ? 256 : 512;
  static_assert(2 * kAccCols <= 512, "accumulators do not fit TMEM");
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be a multiple of 64 (MN-major TMA atoms), <= 256");

  // 1024-B aligned carve-up (swizzle-128B atoms need it)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* pipe = smem;
  SharedCtl* ctl = reinterpret_cast<SharedCtl*>(smem + L::kPipeBytes);
  uint8_t* epi_smem = smem + L::kPipeBytes + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.batches;

  if (warp == 0 && lane == 0) {
    sm100::prefetch_tmap(tmA0);
    sm100::prefetch_tmap(tmB0);
    if (NSUB == 2) { sm100::prefetch_tmap(tmA1); sm100::prefetch_tmap(tmB1); }
    for (int i = 0; i < kStages; ++i) { sm100::mbar_init(&ctl->full[i], 1); sm100::mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { sm100::mbar_init(&ctl->tmem_full[i], 1); sm100::mbar_init(&ctl->tmem_empty[i], Epi::kEpiWarps); }
    sm100::fence_barrier_init();
  }
  if (warp == 1) {
    sm100::tmem_alloc(&ctl->tmem_base, kTmemCols);
    sm100::tmem_relinquish();
  }
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      PipeState st;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(p, t);
        const int m0 = tc.m_tile * BM, n0 = tc.n_tile * BN;
        for (int sub = 0; sub < NSUB; ++sub) {
          const CUtensorMap* ma = sub ? tmA1 : tmA0;
          const CUtensorMap* mb = sub ? tmB1 : tmB0;
          for (int kb = 0; kb < p.num_k_blocks; ++kb) {
            sm100::mbar_wait(&ctl->empty[st.stage], st.phase ^ 1, 1);
            uint8_t* sa = pipe + st.stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            sm100::mbar_expect_tx(&ctl->full[st.stage], L::kStageBytes);
            const int k0 = kb * BK;
            if (p.batches > 1) {
              if (!p.a_mn_major) {
                sm100::tma_load_3d(sa, ma, &ctl->full[st.stage], k0, m0, tc.batch);
              } else {
#pragma unroll
                for (int c = 0; c < BM / 64; ++c)
                  sm100::tma_load_3d(sa + c * (BK * 128), ma, &ctl->full[st.stage], m0 + c * 64, k0, tc.batch);
              }
              if (!p.b_mn_major) {
                sm100::tma_load_3d(sb, mb, &ctl->full[st.stage], k0, n0, tc.batch);
              } else {
#pragma unroll
                for (int c = 0; c < BN / 64; ++c)
                  sm100::tma_load_3d(sb + c * (BK * 128), mb, &ctl->full[st.stage], n0 + c * 64, k0, tc.batch);
              }
            } else {
              if (!p.a_mn_major) {
                sm100::tma_load_2d(sa, ma, &ctl->full[st.stage], k0, m0);
              } else {
#pragma unroll
                for (int c = 0; c < BM / 64; ++c)
                  sm100::tma_load_2d(sa + c * (BK * 128), ma, &ctl->full[st.stage], m0 + c * 64, k0);
              }
              if (!p.b_mn_major) {
                sm100::tma_load_2d(sb, mb, &ctl->full[st.stage], k0, n0);
              } else {
#pragma unroll
                for (int c = 0; c < BN / 64; ++c)
                  sm100::tma_load_2d(sb + c * (BK * 128), mb, &ctl->full[st.stage], n0 + c * 64, k0);
              }
            }
            st.advance<kStages>();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = sm100::umma_idesc_bf16(BM, BN, p.a_mn_major, p.b_mn_major);
      // per-UMMA_K advance of the descriptor start address, and LBO/SBO per layout
      const uint32_t a_adv = p.a_mn_major ? (UMMA_K * 128) : (UMMA_K * 2);
      const uint32_t b_adv = p.b_mn_major ? (UMMA_K * 128) : (UMMA_K * 2);
      const uint32_t a_lbo = p.a_mn_major ? (BK * 128) : 16, b_lbo = p.b_mn_major ? (BK * 128) : 16;
      PipeState st;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        sm100::mbar_wait(&ctl->tmem_empty[acc_stage], acc_phase ^ 1, 2);
        sm100::tc_fence_after();
        for (int sub = 0; sub < NSUB; ++sub) {
          const uint32_t d_tmem = tmem_base + acc_stage * kAccCols + sub * BN;
          for (int kb = 0; kb < p.num_k_blocks; ++kb) {
            sm100::mbar_wait(&ctl->full[st.stage], st.phase, 3);
            sm100::tc_fence_after();
            const uint32_t sa = sm100::smem_u32(pipe + st.stage * L::kStageBytes);
            const uint32_t sb = sa + L::kABytes;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = sm100::umma_smem_desc(sa + k * a_adv, a_lbo, 1024);
              const uint64_t db = sm100::umma_smem_desc(sb + k * b_adv, b_lbo, 1024);
              sm100::umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
            }
            sm100::umma_commit(&ctl->empty[st.stage]);  // frees the smem slot when these MMAs retire
            st.advance<kStages>();
          }
        }
        sm100::umma_commit(&ctl->tmem_full[acc_stage]);  // accumulators complete -> epilogue
        acc_stage ^= 1;
        if (acc_stage == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int epi_warp = warp - kFirstEpiWarp;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    typename Epi::State state;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord(p, t);
      Epi::prologue(ep, p, tc, acc_stage, epi_warp, lane, epi_smem);
      sm100::mbar_wait(&ctl->tmem_full[acc_stage], acc_phase, 4);
      sm100::tc_fence_after();
      Epi::tile(ep, p, tc, t, tmem_base + acc_stage * kAccCols, acc_stage, epi_warp, lane, epi_smem, state);
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&ctl->tmem_empty[acc_stage]);
      acc_stage ^= 1;
      if (acc_stage == 0) acc_phase ^= 1;
    }
    Epi::finish(ep, p, epi_warp, lane, state);
  }

  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, class Epi>
constexpr int smem_bytes() {
  return SmemLayout<BN>::kPipeBytes + 256 + Epi::kEpiSmemBytes + 1024 /* alignment slack */;
}

// lane quarter of TMEM this warp may read (hardware: warp_id % 4), and which column half it owns
__device__ __forceinline__ int epi_quarter() { return (threadIdx.x >> 5) & 3; }

}  // namespace gemm
}  // namespace dinox

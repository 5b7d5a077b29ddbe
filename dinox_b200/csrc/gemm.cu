// tcgen05 GEMM flavours of the loss head (all bf16 x bf16 -> fp32 in TMEM):
//   dinox_gemm_bf16      plain GEMM with alpha / column bias / accumulate epilogue (layer-1 fwd+bwd,
//                        dW2 = G^T H, dH = G W2).   zoo/arch.py:252-256 + autograd of nn.Linear
//   dinox_head_stats     prototype-logit GEMM fused with online row (max, sum-exp): the logits
//                        never leave the SM (pass 1).     scripts/phase5_big_run.py:703,706
//   dinox_head_grad      student + teacher logit tiles recomputed side by side in TMEM; epilogue
//                        forms softmax / teacher prob, the CE term and dL/dlogits (bf16) (pass 2).
//                        scripts/phase5_big_run.py:703-717 + its autograd
#include "gemm_core.cuh"
#include "tmap.cuh"
#include <stdlib.h>

namespace dinox {
namespace gemm {

// =============================================================================================
// Epilogue 1: store   C = alpha * acc (+ bias[n])  ->  fp32 | bf16, optional C += ...
// Each epilogue warp owns 32 accumulator rows.  It drains them in 128-byte-per-row chunks
// (32 fp32 or 64 bf16 columns): TMEM -> registers (next chunk's tcgen05.ld already in flight) ->
// swizzled per-warp staging buffer in shared memory -> one TMA store per chunk, so global writes
// are full 128 B lines issued asynchronously instead of 32 scattered 16-byte stores per instruction.
// accumulate = TMA reduce-add: the += happens in L2, the old values never pass through the SM.
// Split-K / batched launches address output slab tc.batch + tc.split through the 3rd coordinate.
// =============================================================================================
struct EpiStore {
  static constexpr bool kUsesTmaStore = true;
  static constexpr int kEpiWarps = 4;
  static constexpr int kBufs = 2;
  static constexpr int kEpiSmemBytes = kEpiWarps * kBufs * 4096;
  struct Params {
    int out_bf16;
    int accumulate;
    float alpha;
    const float* alpha_dev;  // optional device scalar multiplied into alpha
    const float* bias_n;     // optional per-column bias
    // optional fp32 addend (M, ld_add) in PROBLEM row coordinates, scaled by add_scale: the locally accumulated
    // gradient of the earlier micro-steps that leaves with this GEMM's tile (reduce-scatter variant)
    const float* add_src = nullptr;
    int64_t ld_add = 0;
    float add_scale = 0.f;
    // ordered split (CoreParams::os_*): one counter per (tile, TMEM lane quarter), zero between launches.  Part j of a
    // tile waits for the counter to read j, adds its contribution (part 0: store / accumulate as configured, parts
    // > 0: always reduce-add) and publishes j + 1 (the last part: 0) once its bulk stores have completed.
    uint32_t* order_flags = nullptr;
  };
  struct State {
    int flip = 0;
  };
  static __device__ __forceinline__ void finish(const Params&, const CoreParams&, int, int lane, State&) {
    if (lane == 0) sm100::tma_store_wait_all<0>();
  }
  template <int BN>
  struct Impl {
    static __device__ __forceinline__ void fetch(const Params&, const CoreParams&, TileCoord, int, int, State&) {}
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int, int, uint8_t*, State&) {}

    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap* tmC,
                                                uint32_t tmem_acc, int, int epi_warp, int lane, uint8_t* smem, State& st) {
      const int q = epi_quarter();
      const int row0 = tc.m_tile * BM + q * 32;
      const float alpha = e.alpha * (e.alpha_dev ? *e.alpha_dev : 1.f);
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
      const int slab = tc.batch + tc.split;
      uint8_t* wbase = smem + epi_warp * (kBufs * 4096);
      constexpr int kChunks = BN / 32;           // 32-column TMEM chunks
      // (the phantom second tile of a ragged last CTA pair stores nothing and owns no counters)
      const bool ordered = tc.kparts > 1 && tc.m_tile < p.num_m_tiles;
      const bool accumulate = e.accumulate || (ordered && tc.kpart > 0);
      const bool first_part = tc.kpart == 0;     // bias / addend enter once
      uint32_t* flag = ordered ? e.order_flags + ((int64_t)tc.n_tile * p.num_m_tiles + tc.m_tile) * 4 + q : nullptr;
      uint32_t ra[32], rb[32];
      sm100::tmem_ld32_nowait(taddr, ra);
      if (ordered && tc.kpart > 0) {             // my predecessor's rows are in memory before my first reduce-add
        if (lane == 0) sm100::flag_wait(flag, (uint32_t)tc.kpart);
        __syncwarp();
      }
      sm100::tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t (&cur)[32] = (c & 1) ? rb : ra;
        uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
        if (c + 1 < kChunks) sm100::tmem_ld32_nowait(taddr + (c + 1) * 32, nxt);
        sm100::pin32(cur);
        const int col0 = tc.n_tile * BN + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]) * alpha;
        if (e.add_src && first_part) {
          const int64_t grow = (int64_t)tc.row_shift + row0 + lane;     // this thread's row of the whole problem
          if (grow < (int64_t)p.M + tc.row_shift && grow < p.M && col0 + 32 <= p.N) {
            const float4* src = reinterpret_cast<const float4*>(e.add_src + grow * e.ld_add + col0);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 a4 = __ldg(src + j / 4);
              v[j] = fmaf(a4.x, e.add_scale, v[j]); v[j + 1] = fmaf(a4.y, e.add_scale, v[j + 1]);
              v[j + 2] = fmaf(a4.z, e.add_scale, v[j + 2]); v[j + 3] = fmaf(a4.w, e.add_scale, v[j + 3]);
            }
          } else if (grow < p.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] = fmaf(__ldg(e.add_src + grow * e.ld_add + col0 + j), e.add_scale, v[j]);
          }
        }
        if (e.bias_n && first_part) {
          if (col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias_n + col0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (col0 + j < p.N) ? __ldg(e.bias_n + col0 + j) : 0.f;
          }
        }
        WarpStage stg;
        if (e.out_bf16) {
          // two TMEM chunks (64 columns) fill one 128-byte row; flush after the odd chunk
          const bool first = (c & 1) == 0;
          uint8_t* wbuf = wbase + st.flip * 4096;
          if (first) {
            if (lane == 0) sm100::tma_store_wait_read<kBufs - 1>();
            __syncwarp();
          }
          stg.init(wbuf, lane);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
            stg.put((first ? 0 : 4) + j, *reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1),
                    *reinterpret_cast<uint32_t*>(&t2), *reinterpret_cast<uint32_t*>(&t3));
          }
          if (!first || c + 1 == kChunks) {
            sm100::fence_proxy_async_smem();
            __syncwarp();
            // the group is committed even when the box is entirely out of range: wait_read<kBufs-1>
            // counts groups, so a skipped commit would let the NEXT chunk overwrite a buffer whose
            // store (from the previous tile) is still being read
            if (lane == 0) {
              if (row0 < p.M && col0 - (first ? 0 : 32) < p.N) sm100::tma_store_3d(tmC, wbuf, col0 - (first ? 0 : 32), row0, slab);
              sm100::tma_store_commit();
            }
            st.flip ^= 1;
          }
        } else {
          uint8_t* wbuf = wbase + st.flip * 4096;
          if (lane == 0) sm100::tma_store_wait_read<kBufs - 1>();
          __syncwarp();
          stg.init(wbuf, lane);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            stg.put(j, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                    __float_as_uint(v[4 * j + 3]));
          sm100::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (row0 < p.M && col0 < p.N) {
              if (accumulate) sm100::tma_reduce_add_3d(tmC, wbuf, col0, row0, slab);
              else sm100::tma_store_3d(tmC, wbuf, col0, row0, slab);
            }
            sm100::tma_store_commit();   // always: see the bf16 branch
          }
          st.flip ^= 1;
        }
        if (c + 1 < kChunks) sm100::tmem_wait_ld();
      }
      if (ordered) {   // hand the tile's rows of this lane quarter to the next part
        if (lane == 0) {
          sm100::tma_store_wait_all<0>();        // completed, not merely read out of the staging buffers
          __threadfence();
          sm100::flag_release(flag, tc.kpart + 1 == tc.kparts ? 0u : (uint32_t)(tc.kpart + 1));
        }
        __syncwarp();
      }
    }
  };
};

// =============================================================================================
// Epilogue 1b: direct stores (one output row per thread).  Only the rare bf16 read-modify-write
// (accumulate into a bf16 matrix) still goes this way; everything else uses EpiStore below.
// =============================================================================================
struct EpiStoreDirect {
  static constexpr bool kUsesTmaStore = false;
  static constexpr int kEpiWarps = 4;
  static constexpr int kEpiSmemBytes = 0;
  struct Params {
    void* out;
    int64_t ldo;
    int out_bf16;
    int accumulate;
    float alpha;
    const float* alpha_dev;  // optional device scalar multiplied into alpha
    const float* bias_n;     // optional per-column bias
    int64_t batch_stride;    // output elements between problems of a batched launch
  };
  struct State {};
  static __device__ __forceinline__ void finish(const Params&, const CoreParams&, int, int, State&) {}
  template <int BN>
  struct Impl {
    static __device__ __forceinline__ void fetch(const Params&, const CoreParams&, TileCoord, int, int, State&) {}
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int, int, uint8_t*, State&) {}
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap*,
                                                uint32_t tmem_acc, int, int, int lane, uint8_t*, State&) {
      const int q = epi_quarter();
      const int row = tc.m_tile * BM + q * 32 + lane;
      const float alpha = e.alpha * (e.alpha_dev ? *e.alpha_dev : 1.f);
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        sm100::tmem_ld32(taddr + c * 32, v);
        const int col0 = tc.n_tile * BN + c * 32;
        if (row < p.M && col0 < p.N) {
          const bool full = (col0 + 32 <= p.N);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float b = (e.bias_n && (full || col0 + j < p.N)) ? __ldg(e.bias_n + col0 + j) : 0.f;
            v[j] = fmaf(v[j], alpha, b);
          }
          if (e.out_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + tc.batch * e.batch_stride + (int64_t)row * e.ldo + col0;
            if (full && !e.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 w;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                w.x = *reinterpret_cast<uint32_t*>(&t0); w.y = *reinterpret_cast<uint32_t*>(&t1);
                w.z = *reinterpret_cast<uint32_t*>(&t2); w.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(o + j) = w;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16_rn(v[j] + (e.accumulate ? __bfloat162float(o[j]) : 0.f));
            }
          } else {
            float* o = reinterpret_cast<float*>(e.out) + tc.batch * e.batch_stride + (int64_t)row * e.ldo + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 w = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (e.accumulate) {
                  float4 old = *reinterpret_cast<float4*>(o + j);
                  w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
                }
                *reinterpret_cast<float4*>(o + j) = w;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = v[j] + (e.accumulate ? o[j] : 0.f);
            }
          }
        }
      }
    }
  };
};

// =============================================================================================
// Epilogue 2: row statistics of u2 = acc*scale2 + col2[n]  (log2 units).
// Each epilogue warp owns 32 rows x (BN/2) columns; it writes one (max, sumexp2) pair per row
// into partial[row][n_tile*2 + half]; dinox_head_stats then merges the pairs per row.
// The column offsets of the tile are staged in shared memory by the prologue, with -inf for
// columns >= N: out-of-range columns then contribute exp2(-inf) = 0 without any per-element
// bounds check (the accumulator itself is 0 there because TMA zero-fills the W2 rows).
// Issue budget per element: FFMA, 1/2 FMNMX3, FADD, MUFU.EX2, FADD (+ 1/4 LDS.128).
// =============================================================================================
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sm100::smem_u32(p)));
  return v;
}

#ifndef DINOX_EXP_EPI_MODE
#define DINOX_EXP_EPI_MODE 0
#endif
#ifndef DINOX_EPI_WARPS
#define DINOX_EPI_WARPS 8   // epilogue warps of the two row-math kernels (pass 1 / pass 2): 8 or 16
#endif

// ---- packed fp32 pairs (sm_100: add / mul / fma .f32x2 operate on two floats held in a 64-bit register) ----
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// kRun: the kernel walks contiguous runs of N tiles per M tile (resident-A schedule), so every thread keeps the
// running (max, sum) of its row across the tiles of a run and writes ONE partial per (row, cluster that touched
// the row's M tile) instead of one per (row, 128 prototypes): 33 MB -> 0.5 MB of partials at C2.
template <bool kRun>
struct EpiStatsT {
  static constexpr bool kUsesTmaStore = false;
  static constexpr int kEpiWarps = DINOX_EPI_WARPS;
  static constexpr int kGroups = kEpiWarps / 4;          // column groups (each TMEM lane quarter has kGroups warps)
  static constexpr int kColsW = 256 / kGroups;           // columns per warp (128 or 64)
  // every warp keeps its own copy of the column offsets of its kColsW columns: no CTA-wide barrier per tile
  static constexpr int kEpiSmemBytes = kEpiWarps * kColsW * 4;
  struct Params {
    float scale2;
    const float* col2;   // (N) log2-unit column offsets, may be NULL
    float2* partial;     // (M, kGroups*num_n_tiles); kRun: (M, kGroups*run_slots)
    int cl, per, run_slots;   // kRun: cluster size, super tiles per cluster, partial slots per row
  };
  struct State {
    float col[kColsW / 32];   // raw prefetched column offsets of the NEXT tile (this lane's columns of the warp's group)
    float run_m = -INFINITY, run_s = 0.f;   // kRun: statistics of this thread's row over the current run
    int run_mtile = -1;
  };
  // slot of this cluster among the clusters whose contiguous chunk [c*per, (c+1)*per) meets the row's M tile
  static __device__ __forceinline__ void flush(const Params& e, const CoreParams& p, int epi_warp, int lane, State& st) {
    if (st.run_mtile < 0) return;
    const int row = st.run_mtile * BM + epi_quarter() * 32 + lane;
    const int c_lo = (int)(((int64_t)(st.run_mtile / e.cl) * p.num_n_tiles) / e.per);
    const int slot = ((int)blockIdx.x / e.cl - c_lo) * kGroups + (epi_warp >> 2);
    if (row < p.M) e.partial[(int64_t)row * (kGroups * e.run_slots) + slot] = make_float2(st.run_m, st.run_s);
    st.run_m = -INFINITY; st.run_s = 0.f;
  }
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams& p, int epi_warp, int lane, State& st) {
    if (kRun) flush(e, p, epi_warp, lane, st);
  }
  template <int BN>
  struct Impl {
    static_assert(BN == 256, "EpiStats is written for 256-wide tiles");
    static constexpr int kCols = kColsW;
    static __device__ __forceinline__ void fetch(const Params& e, const CoreParams& p, TileCoord tc, int epi_warp, int lane,
                                                 State& st) {
      const int col0 = tc.n_tile * BN + (epi_warp >> 2) * kCols;
#pragma unroll
      for (int j = 0; j < kCols / 32; ++j) {
        const int col = col0 + j * 32 + lane;
        st.col[j] = (col < p.N) ? (e.col2 ? __ldg(e.col2 + col) : 0.f) : -INFINITY;
      }
    }
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int epi_warp, int lane,
                                                    uint8_t* smem, State& st) {
      float* buf = reinterpret_cast<float*>(smem) + epi_warp * kCols;
      __syncwarp();   // every lane is done with the previous tile's offsets
#pragma unroll
      for (int j = 0; j < kCols / 32; ++j) buf[j * 32 + lane] = st.col[j];
      __syncwarp();
    }
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap*,
                                                uint32_t tmem_acc, int acc_stage, int epi_warp, int lane, uint8_t* smem, State& st) {
      const float* buf = reinterpret_cast<const float*>(smem) + epi_warp * kCols;
      const int q = epi_quarter();
      const int grp = epi_warp >> 2;
      const int row = tc.m_tile * BM + q * 32 + lane;
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + grp * kCols;
      if (kRun && tc.m_tile != st.run_mtile) {   // a new run: hand the finished one over
        flush(e, p, epi_warp, lane, st);
        st.run_mtile = tc.m_tile;
      }
      float m = kRun ? st.run_m : -INFINITY, s = kRun ? st.run_s : 0.f;
#pragma unroll 1
      for (int c = 0; c < kCols / 64; ++c) {
        float v[2][32];
        sm100::tmem_ld32x2(taddr + c * 64, taddr + c * 64 + 32, v[0], v[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float* cb = buf + c * 64 + h * 32;
          const uint64_t sc2 = pack2(e.scale2, e.scale2);
          uint64_t u[16];          // 32 values as packed pairs
          float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = lds128(cb + j);
            u[j / 2] = fma2(pack2(v[h][j], v[h][j + 1]), sc2, pack2(b4.x, b4.y));
            u[j / 2 + 1] = fma2(pack2(v[h][j + 2], v[h][j + 3]), sc2, pack2(b4.z, b4.w));
            float x0, x1, x2, x3;
            unpack2(u[j / 2], x0, x1); unpack2(u[j / 2 + 1], x2, x3);
            cm0 = fmaxf(cm0, fmaxf(x0, x1));
            cm1 = fmaxf(cm1, fmaxf(x2, x3));
          }
          const float mn = fmaxf(m, fmaxf(cm0, cm1));
          // a chunk that is entirely out of range keeps (m, s) untouched (mn may still be -inf)
          const float msafe = (mn == -INFINITY) ? 0.f : mn;
          const uint64_t nm2 = pack2(-msafe, -msafe);
          uint64_t acc0 = pack2(s * fast_ex2(m - msafe), 0.f), acc1 = pack2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float x0, x1, x2, x3;
            unpack2(add2(u[j], nm2), x0, x1); unpack2(add2(u[j + 1], nm2), x2, x3);
            acc0 = add2(acc0, pack2(fast_ex2(x0), fast_ex2(x1)));
            acc1 = add2(acc1, pack2(fast_ex2(x2), fast_ex2(x3)));
          }
          float a0, a1, a2, a3;
          unpack2(acc0, a0, a1); unpack2(acc1, a2, a3);
          s = (a0 + a1) + (a2 + a3);
          m = mn;
        }
      }
      if (kRun) { st.run_m = m; st.run_s = s; }
      else if (row < p.M) e.partial[(int64_t)row * (kGroups * p.num_n_tiles) + tc.n_tile * kGroups + grp] = make_float2(m, s);
    }
  };
};
using EpiStats = EpiStatsT<false>;
using EpiStatsRun = EpiStatsT<true>;

__global__ void stats_merge_kernel(const float2* __restrict__ partial, int64_t rows, int n_part,
                                   float* __restrict__ lse_nat, float* __restrict__ lse2) {
  // one warp per row
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  MaxSum a{-INFINITY, 0.f};
  for (int i = lane; i < n_part; i += 32) {
    float2 v = partial[row * n_part + i];
    a = maxsum_merge(a, MaxSum{v.x, v.y});
  }
  a = warp_maxsum(a);
  if (lane == 0) {
    const float l2 = a.m + log2f(a.s);
    if (lse2) lse2[row] = l2;
    if (lse_nat) lse_nat[row] = l2 * DINOX_LN2;
  }
}

// run-mode partials: row r of M super tile mt has one slot per (cluster meeting mt, column group)
__global__ void stats_merge_run_kernel(const float2* __restrict__ partial, int64_t rows, int groups, int run_slots,
                                       int num_n, int per, int cl, float* __restrict__ lse_nat, float* __restrict__ lse2) {
  // one warp per row (a single M tile spread over every CTA has > 100 slots)
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int64_t mt = row / (BM * cl);
  const int c_lo = (int)((mt * num_n) / per), c_hi = (int)(((mt + 1) * num_n - 1) / per);
  MaxSum a{-INFINITY, 0.f};
  const float2* pr = partial + row * (int64_t)(groups * run_slots);
  for (int i = lane; i < (c_hi - c_lo + 1) * groups; i += 32) a = maxsum_merge(a, MaxSum{pr[i].x, pr[i].y});
  a = warp_maxsum(a);
  if (lane == 0) {
    const float l2 = a.m + log2f(a.s);
    if (lse2) lse2[row] = l2;
    if (lse_nat) lse_nat[row] = l2 * DINOX_LN2;
  }
}

// =============================================================================================
// Epilogue 2b (teacher, ONE pass over the prototypes): the row statistics of EpiStatsT PLUS the teacher
// probabilities themselves, un-normalised, so that pass 2 never recomputes a teacher logit.
// Each epilogue warp owns 32 rows x one 128-prototype GRANULE of the 256-wide tile:
//   x[row,k]     = acc*scale2 + col2[k]                     (log2 units; col2 = (b2 - centre)/tau * log2e)
//   gmax         = max over the granule of x                (sweep 1 over TMEM)
//   qt[row,k]    = fp16( 2^(x - gmax) )  in (0, 1]          (sweep 2; swizzled staging + TMA store)
//   ref[g][row]  = gmax                                      g = 2*n_tile + column group
//   (m, s)      += online (max, sum 2^x) of the row          (fp32, from the unrounded exponentials)
// Later  q[row,k] = qt[row,k] * 2^(ref[g][row] - rowbias2[row])  with rowbias2 = the row's log2 LSE (centre
// teacher) or its Sinkhorn row offset.  Every value that matters for the row (within 14 binades of the row
// maximum) is a NORMAL fp16 number relative to its own granule maximum (gmax <= row max), i.e. carries a
// 2^-12 relative rounding; smaller ones fall into fp16 denormals with an absolute error below 2^-25 of a
// probability.  (Storing the teacher LOGITS in 16 bits instead would cost 2^-9 * |x| ~ 10 % on q at
// 1/tau_t = 25 - the reason pass 2 used to recompute them.)
// scripts/phase5_big_run.py:703 (teacher softmax) with the statistics of :686-690 / the SK row sums.
// =============================================================================================
__device__ __forceinline__ uint32_t f16x2_bits(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

#ifndef DINOX_RB_EPI_WARPS
#define DINOX_RB_EPI_WARPS 8   // epilogue warps of the two read-back kernels (teacher pass / pass 2): 8, 12 or 16 (16: register cap 96, 3 operand stages - measured slower)
#endif
// Tile width (prototypes) of the two read-back kernels: 256 with 8 or 16 epilogue warps (granules of 128 / 64
// prototypes), 192 with 12 (three granules of 64: three epilogue warps per scheduler instead of two)
#ifndef DINOX_RB_TILE
#define DINOX_RB_TILE 256
#endif
constexpr int kRbTile = DINOX_RB_TILE;
static_assert((DINOX_RB_TILE == 256 && (DINOX_RB_EPI_WARPS == 8 || DINOX_RB_EPI_WARPS == 16)) ||
                  (DINOX_RB_TILE == 192 && DINOX_RB_EPI_WARPS == 12),
              "read-back kernels: 256-wide tiles with 8 / 16 epilogue warps or 192-wide tiles with 12");
template <bool kRun, int W>
struct EpiTeachQT {
  static constexpr bool kUsesTmaStore = true;
  static constexpr int kEpiWarps = W;                    // 8, 12 or 16
  static constexpr int kGroups = W / 4;                  // column groups per TMEM lane quarter
  static constexpr int kColsW = kRbTile / kGroups;       // one granule (128 or 64 prototypes) per warp and tile
  // one staging buffer per warp (32 rows x 128 B) leaves five 16 KB operand stages beside the resident A rows;
  // the wait for the previous store's read sits behind the math of the next 32 columns
#ifndef DINOX_TEACHQ_STAGING_BUFS
#define DINOX_TEACHQ_STAGING_BUFS 1
#endif
  static constexpr int kBufs = DINOX_TEACHQ_STAGING_BUFS;
  static constexpr int kStageBytes = kEpiWarps * kBufs * 4096;
  static constexpr int kEpiSmemBytes = kStageBytes + kEpiWarps * kColsW * 4;
  struct Params {
    float scale2;
    const float* col2;        // (N) log2-unit column offsets, may be NULL
    const float* col2_alt;    // (N) offsets of the rows of M tiles >= alt_from_mtile (iBOT patch centre), may be NULL
    int alt_from_mtile;
    float2* partial;          // (M, kGroups*num_n_tiles); kRun: (M, kGroups*run_slots)
    int cl, per, run_slots;
    float* refs;              // (kGroups*num_n_tiles, ld_refs) granule maxima, log2 units
    int64_t ld_refs;
  };
  struct State {
    float col[kColsW / 32];
    float run_m = -INFINITY, run_s = 0.f;
    int run_mtile = -1;
    int flip = 0;
  };
  static __device__ __forceinline__ void flush(const Params& e, const CoreParams& p, int epi_warp, int lane, State& st) {
    if (st.run_mtile < 0) return;
    const int row = st.run_mtile * BM + epi_quarter() * 32 + lane;
    const int c_lo = (int)(((int64_t)(st.run_mtile / e.cl) * p.num_n_tiles) / e.per);
    const int slot = ((int)blockIdx.x / e.cl - c_lo) * kGroups + (epi_warp >> 2);
    if (row < p.M) e.partial[(int64_t)row * (kGroups * e.run_slots) + slot] = make_float2(st.run_m, st.run_s);
    st.run_m = -INFINITY; st.run_s = 0.f;
  }
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams& p, int epi_warp, int lane, State& st) {
    if (kRun) flush(e, p, epi_warp, lane, st);
    if (lane == 0) sm100::tma_store_wait_all<0>();
  }
  template <int BN>
  struct Impl {
    static_assert(BN == kRbTile && kColsW % 64 == 0, "EpiTeachQ: tile width / epilogue warp count mismatch");
    static __device__ __forceinline__ void fetch(const Params& e, const CoreParams& p, TileCoord tc, int epi_warp, int lane,
                                                 State& st) {
      const int col0 = tc.n_tile * BN + (epi_warp >> 2) * kColsW;
      const float* c2 = (e.col2_alt && tc.m_tile >= e.alt_from_mtile) ? e.col2_alt : e.col2;
#pragma unroll
      for (int j = 0; j < kColsW / 32; ++j) {
        const int col = col0 + j * 32 + lane;
        st.col[j] = (col < p.N) ? (c2 ? __ldg(c2 + col) : 0.f) : -INFINITY;
      }
    }
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int epi_warp, int lane,
                                                    uint8_t* smem, State& st) {
      float* buf = reinterpret_cast<float*>(smem + kStageBytes) + epi_warp * kColsW;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kColsW / 32; ++j) buf[j * 32 + lane] = st.col[j];
      __syncwarp();
    }
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap* tmC,
                                                uint32_t tmem_acc, int, int epi_warp, int lane, uint8_t* smem, State& st) {
      const float* buf = reinterpret_cast<const float*>(smem + kStageBytes) + epi_warp * kColsW;
      const int q = epi_quarter();
      const int grp = epi_warp >> 2;
      const int row0 = tc.m_tile * BM + q * 32;
      const int row = row0 + lane;
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + grp * kColsW;
      if (kRun && tc.m_tile != st.run_mtile) {
        flush(e, p, epi_warp, lane, st);
        st.run_mtile = tc.m_tile;
      }
      const uint64_t sc2 = pack2(e.scale2, e.scale2);
      // ---- sweep 1: granule maximum of this thread's row
      float gm = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kColsW / 32; ++c) {
        float v[32];
        sm100::tmem_ld32(taddr + c * 32, v);
        const float* cb = buf + c * 32;
        float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = lds128(cb + j);
          float x0, x1, x2, x3;
          unpack2(fma2(pack2(v[j], v[j + 1]), sc2, pack2(b4.x, b4.y)), x0, x1);
          unpack2(fma2(pack2(v[j + 2], v[j + 3]), sc2, pack2(b4.z, b4.w)), x2, x3);
          cm0 = fmaxf(cm0, fmaxf(x0, x1));
          cm1 = fmaxf(cm1, fmaxf(x2, x3));
        }
        gm = fmaxf(gm, fmaxf(cm0, cm1));
      }
      const float gsafe = (gm == -INFINITY) ? 0.f : gm;   // a granule entirely beyond N: every exponential is 2^-inf = 0
      const uint64_t ng2 = pack2(-gsafe, -gsafe);
      // ---- sweep 2: exponentials -> row sum (fp32) and fp16 staging -> TMA store
      uint64_t acc0 = pack2(0.f, 0.f), acc1 = pack2(0.f, 0.f);
      uint8_t* wbase = smem + epi_warp * (kBufs * 4096);
      uint32_t ra[32], rb[32];
      sm100::tmem_ld32_nowait(taddr, ra);
      sm100::tmem_wait_ld();
      WarpStage stg;
      uint8_t* wbuf = wbase;
#pragma unroll
      for (int c = 0; c < kColsW / 32; ++c) {
        uint32_t (&cur)[32] = (c & 1) ? rb : ra;
        uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
        if (c + 1 < kColsW / 32) sm100::tmem_ld32_nowait(taddr + (c + 1) * 32, nxt);
        sm100::pin32(cur);
        const float* cb = buf + c * 32;
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = lds128(cb + j);
          float x0, x1, x2, x3;
          unpack2(add2(fma2(pack2u(cur[j], cur[j + 1]), sc2, pack2(b4.x, b4.y)), ng2), x0, x1);
          unpack2(add2(fma2(pack2u(cur[j + 2], cur[j + 3]), sc2, pack2(b4.z, b4.w)), ng2), x2, x3);
          const float e0 = fast_ex2(x0), e1 = fast_ex2(x1), e2 = fast_ex2(x2), e3 = fast_ex2(x3);
          acc0 = add2(acc0, pack2(e0, e1));
          acc1 = add2(acc1, pack2(e2, e3));
          packed[j / 2] = f16x2_bits(e0, e1);
          packed[j / 2 + 1] = f16x2_bits(e2, e3);
        }
        if ((c & 1) == 0) {   // the staging buffer about to be refilled must no longer be read by the previous store
          wbuf = wbase + st.flip * 4096;
          if (lane == 0) sm100::tma_store_wait_read<kBufs - 1>();
          __syncwarp();
          stg.init(wbuf, lane);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stg.put((c & 1) * 4 + j, packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        if (c & 1) {
          sm100::fence_proxy_async_smem();
          __syncwarp();
          const int col0 = tc.n_tile * BN + grp * kColsW + (c >> 1) * 64;
          if (lane == 0) {
            // the map's column extent is the PADDED row length, so columns in [N, ld) receive the zeros computed
            // for them (col2 = -inf) and pass 2 may read whole 256-byte row segments
#if DINOX_STREAM_EVICT_FIRST
            if (row0 < p.M) sm100::tma_store_3d_hint(tmC, wbuf, col0, row0, 0, sm100::l2_policy_evict_first());
#else
            if (row0 < p.M) sm100::tma_store_3d(tmC, wbuf, col0, row0, 0);
#endif
            sm100::tma_store_commit();   // always: wait_read<kBufs-1> counts groups
          }
          if (kBufs > 1) st.flip ^= 1;
        }
        if (c + 1 < kColsW / 32) sm100::tmem_wait_ld();
      }
      float a0, a1, a2, a3;
      unpack2(acc0, a0, a1); unpack2(acc1, a2, a3);
      const float s_tile = (a0 + a1) + (a2 + a3);
      float m = kRun ? st.run_m : -INFINITY, s = kRun ? st.run_s : 0.f;
      const float mn = fmaxf(m, gm);
      const float msafe = (mn == -INFINITY) ? 0.f : mn;
      s = s * fast_ex2(m - msafe) + s_tile * fast_ex2(gsafe - msafe);
      m = mn;
      if (kRun) { st.run_m = m; st.run_s = s; }
      else if (row < p.M) e.partial[(int64_t)row * (kGroups * p.num_n_tiles) + tc.n_tile * kGroups + grp] = make_float2(m, s);
      if (row < p.M) e.refs[(int64_t)(tc.n_tile * kGroups + grp) * e.ld_refs + row] = gm;
    }
  };
};

// =============================================================================================
// Epilogue 3b (pass 2, student logits only): TMEM lanes = entries e, columns = prototypes k.
//   u      = S*as2 + cs2[k] - lse2[e]                 log2 of the student softmax prob, p = 2^u
//   q'     = qt[trow[e],k] * cwt*2^(ref - rb2[e])     cwt = cw[e]/tau_s; the teacher prob from EpiTeachQ (fp16, HBM)
//   G[e,k] = cwt*p - q'                               -> bf16, (E, ldg) entry-major
//   loss  -= q' * u * ln2 * tau_s                     (-cw q ln p)
//   db2[k] partial = sum of the bf16 G over the 32 entries of the warp (column sums of the staged tile)
// One sub-GEMM, so the tile is 256 wide with double-buffered accumulators like pass 1 and the A rows of an M
// tile can stay resident; the teacher probabilities arrive as 256-byte row segments straight from global
// memory into registers (one 32-byte load per 16 prototypes, 4 chunks in flight per thread).
// scripts/phase5_big_run.py:706-717 and its autograd (grad of log_softmax) in one kernel.
// =============================================================================================
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
// ... with an L2 eviction-priority hint: streams that are read exactly once must not displace the W2 tiles that
// every M tile re-reads from L2
__device__ __forceinline__ void ldg256_hint(const void* p, uint32_t (&r)[8], uint64_t pol) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p), "l"(pol));
}
#ifndef DINOX_STREAM_EVICT_FIRST
#define DINOX_STREAM_EVICT_FIRST 1
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// "minus infinity" that stays finite under 0 * x
#define DINOX_NEG_HUGE (-1.0e30f)
#ifndef DINOX_EXP_G2
#define DINOX_EXP_G2 0
#endif
template <int W>
struct EpiGradRT {
  static constexpr bool kUsesTmaStore = true;
  static constexpr int kEpiWarps = W;                    // 8: 2 x 128 columns per TMEM lane quarter; 16: 4 x 64; 12: 3 x 64 (192-wide tile)
  static constexpr int kGroups = W / 4;
  static constexpr int kColsW = kRbTile / kGroups;
  static constexpr int kChunks = kColsW / 16;
#ifndef DINOX_GRADR_SLOTS
#define DINOX_GRADR_SLOTS 4
#endif
  static constexpr int kSlots = DINOX_GRADR_SLOTS;       // 16-prototype chunks of teacher probabilities in flight
#ifndef DINOX_GRADR_STAGING_BUFS
#define DINOX_GRADR_STAGING_BUFS 1
#endif
  static constexpr int kBufs = DINOX_GRADR_STAGING_BUFS;
  static constexpr int kStageBytes = kEpiWarps * kBufs * 4096;
  static constexpr int kEpiSmemBytes = kStageBytes + kEpiWarps * kColsW * 4;
  struct Params {
    float as2, inv_tau_s;
    const float* cs2;        // (N) student column offsets, log2 units
    const float* lse2;       // (M) student log2 LSE per entry
    const float* cw;         // (M) entry weight, 0 for padding
    const float* rb2;        // (M) teacher row offset (log2) per entry
    const int* trow;         // (M) row of qt / refs holding the entry's teacher
    const int* srow;         // NULL: lse2 / rb2 are per ENTRY.  Else (M) student row of the entry (-1 = padding) and
                             // lse2 is indexed by student row, rb2 by teacher row (trow): no per-entry gathers upstream
    const __half* qt;        // (teacher rows, ldq) un-normalised teacher probabilities
    int64_t ldq;
    const float* refs;       // (kGroups*num_n_tiles, ld_refs)
    int64_t ld_refs;
    int alt_from;            // entries >= alt_from go to loss[1]
    int run_rows;            // 1: contiguous prototype runs per M tile (consecutive tiles of a CTA share the row)
    int m_step;              // > 0: column sweep, consecutive tiles of a CTA are m_step M tiles apart
    float* db2_partial;      // (ceil(M/32), N) or NULL
    float* loss_partial;     // (gridDim.x * kEpiWarps * 2)
    // ticket != NULL: the epilogue warp that finishes LAST adds up all loss partials itself (fixed order) and writes
    // loss_out[0..2]; the counter returns to 0 for the next launch.  NULL: a pair_sum launch follows.
    unsigned* ticket = nullptr;
    float* loss_out = nullptr;
    int loss_accumulate = 0;
  };
  struct State {
    float loss_a = 0.f, loss_b = 0.f;
    float col[kColsW / 32];
    // raw per-row values of the NEXT tile (loaded by fetch(); no arithmetic on them until prologue)
    float n_lse = 0.f, n_cw = 0.f, n_rb = 0.f;
    int n_trow = 0, n_mtile = -1, n_ntile = 0;
    int n2_trow = 0, n2_mtile = -1;   // column sweep: teacher row of the tile after the next one
    // the tile being processed
    float nl = 0.f, cwt = 0.f, rb = 0.f;
    int cur_mtile = -1, cur_trow = 0;
    bool in_b = false;
    // teacher probabilities of the tile about to be processed: kSlots x 16 prototypes.  tile() refills the slots it
    // has drained with the first chunks of the NEXT tile (its row is known from fetch()), so that their global-memory
    // latency hides behind the rest of this tile instead of standing between two tiles
    uint32_t qb[kSlots][8];
    float ref_ahead = 0.f;   // granule maximum of the tile about to be processed
    bool ahead = false;      // qb / ref_ahead already hold the next tile's values
    int flip = 0;
  };
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams&, int epi_warp, int lane, State& st) {
    const float a = warp_sum(st.loss_a), b = warp_sum(st.loss_b);
    const float sc = -DINOX_LN2 / e.inv_tau_s;   // the tiles accumulate (cw/tau_s) * q * u
    if (lane == 0) {
      e.loss_partial[((int64_t)blockIdx.x * kEpiWarps + epi_warp) * 2 + 0] = a * sc;
      e.loss_partial[((int64_t)blockIdx.x * kEpiWarps + epi_warp) * 2 + 1] = b * sc;
      sm100::tma_store_wait_all<0>();
    }
    if (e.ticket) {
      const unsigned n = gridDim.x * kEpiWarps;
      unsigned t = 0;
      if (lane == 0) {
        __threadfence();                       // the partials above are visible before the ticket is
        t = atomicAdd(e.ticket, 1u);
      }
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t == n - 1) {                        // every other warp of the grid has published its partials
        __threadfence();
        float s0 = 0.f, s1 = 0.f;
        for (unsigned i = lane; i < n; i += 32) {
          s0 += __ldcg(e.loss_partial + 2 * i);
          s1 += __ldcg(e.loss_partial + 2 * i + 1);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) {
          const float o0 = s0 + (e.loss_accumulate ? e.loss_out[0] : 0.f), o1 = s1 + (e.loss_accumulate ? e.loss_out[1] : 0.f);
          e.loss_out[0] = o0; e.loss_out[1] = o1; e.loss_out[2] = o0 + o1;
          *e.ticket = 0u;
        }
      }
    }
  }
  template <int BN>
  struct Impl {
    static_assert(BN == kRbTile && kColsW % 64 == 0, "EpiGradR: tile width / epilogue warp count mismatch");
    static __device__ __forceinline__ void ldq(const __half* p, uint32_t (&r)[8], uint64_t pol) {
#if DINOX_EXP_G2 & 1   // experiment: no teacher-probability loads
      return;
#endif
#if DINOX_STREAM_EVICT_FIRST
      ldg256_hint(p, r, pol);
#else
      ldg256(p, r);
#endif
    }
    // first prototype of the granule this warp owns in tile n_tile, clamped into the padded row (ldq >= 256 * tiles)
    static __device__ __forceinline__ const __half* granule(const Params& e, int trow, int n_tile, int grp) {
      return e.qt + (int64_t)trow * e.ldq + (n_tile * BN + grp * kColsW);
    }
    static __device__ __forceinline__ void fetch(const Params& e, const CoreParams& p, TileCoord tc, int epi_warp, int lane,
                                                 State& st) {
      const int grp = epi_warp >> 2;
      const int col0 = tc.n_tile * BN + grp * kColsW;
#pragma unroll
      for (int j = 0; j < kColsW / 32; ++j) {
        const int col = col0 + j * 32 + lane;
        st.col[j] = (col < p.N) ? __ldg(e.cs2 + col) : DINOX_NEG_HUGE;
      }
      if (tc.m_tile != st.n_mtile) {   // a new row for this thread (once per run in the resident-A schedule)
        const int row = tc.m_tile * BM + epi_quarter() * 32 + lane;
        const bool ok = row < p.M;
        st.n_cw = ok ? __ldg(e.cw + row) : 0.f;
        st.n_trow = ok ? __ldg(e.trow + row) : 0;
        if (e.srow) {   // row statistics live per student / teacher ROW; dead entries (cw = 0) never use them
          const int sr = ok ? __ldg(e.srow + row) : -1;
          st.n_lse = sr >= 0 ? __ldg(e.lse2 + sr) : 0.f;
          st.n_rb = ok ? __ldg(e.rb2 + st.n_trow) : 0.f;
        } else {
          st.n_lse = ok ? __ldg(e.lse2 + row) : 0.f;
          st.n_rb = ok ? __ldg(e.rb2 + row) : 0.f;
        }
        st.n_mtile = tc.m_tile;
        if (e.m_step > 0) {
          // column sweep: the row changes with every tile.  The teacher row of THIS tile was looked up one fetch ago
          // (n2_*): pull its granule of probabilities towards L2 now, a whole tile before the register loads ask
          // for it, and look up the row of the tile after this one.
          if (st.n2_mtile == tc.m_tile) {
            const __half* nq = granule(e, st.n2_trow, tc.n_tile, grp);
            prefetch_l2(nq);
            if (kColsW > 64) prefetch_l2(nq + 64);
          }
          const int row2 = row + e.m_step * BM;
          st.n2_mtile = tc.m_tile + e.m_step;
          st.n2_trow = row2 < p.M ? __ldg(e.trow + row2) : 0;
        }
      }
      st.n_ntile = tc.n_tile;
    }
    static __device__ __forceinline__ void prologue(const Params& e, const CoreParams& p, TileCoord tc, int, int epi_warp, int lane,
                                                    uint8_t* smem, State& st) {
      float* buf = reinterpret_cast<float*>(smem + kStageBytes) + epi_warp * kColsW;
      const int grp = epi_warp >> 2;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kColsW / 32; ++j) buf[j * 32 + lane] = st.col[j];
      if (tc.m_tile != st.cur_mtile) {
        const int row = tc.m_tile * BM + epi_quarter() * 32 + lane;
        const bool live = st.n_cw != 0.f;
        st.cwt = st.n_cw * e.inv_tau_s;
        st.nl = live ? -st.n_lse : DINOX_NEG_HUGE;     // padding entries: p = 2^-huge = 0, and 0 * u stays finite
        st.rb = st.n_rb;
        st.cur_trow = st.n_trow;
        st.in_b = row >= e.alt_from;
        st.cur_mtile = tc.m_tile;
      }
      if (!st.ahead) {   // first tile of this CTA: nobody fetched ahead (n_* describe THIS tile: fetch() ran for it)
        const uint64_t pol = sm100::l2_policy_evict_first();
        const __half* qrow = granule(e, st.n_trow, tc.n_tile, grp);
#pragma unroll
        for (int c = 0; c < (kSlots < kChunks ? kSlots : kChunks); ++c) ldq(qrow + c * 16, st.qb[c], pol);
        st.ref_ahead = __ldg(e.refs + (int64_t)(tc.n_tile * kGroups + grp) * e.ld_refs + st.n_trow);
      }
      __syncwarp();
    }

    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap* tmC,
                                                uint32_t tmem_acc, int, int epi_warp, int lane, uint8_t* smem, State& st) {
      const float* buf = reinterpret_cast<const float*>(smem + kStageBytes) + epi_warp * kColsW;
      const int q = epi_quarter();
      const int grp = epi_warp >> 2;
      const int row0 = tc.m_tile * BM + q * 32;
      const uint64_t pol = sm100::l2_policy_evict_first();
      const uint32_t ts = tmem_acc + ((uint32_t)(q * 32) << 16) + grp * kColsW;
      // By now fetch() has run for the next tile (if there is one): n_trow / n_mtile / n_ntile name it.  Without a next
      // tile they still name this one - the look-ahead loads below then re-read valid memory and are never used.
      const __half* qrow = granule(e, st.cur_trow, tc.n_tile, grp);
      const __half* qnext = granule(e, st.n_trow, st.n_ntile, grp);
      // q' = qt * nsc with nsc = -(cw/tau_s) * 2^(ref - rb); dead entries (cw = 0) must not see 0 * inf
      const float nsc = (st.cwt != 0.f) ? -(st.cwt * fast_ex2(st.ref_ahead - st.rb)) : 0.f;
      // ... and the next tile's granule maximum starts its trip now: a whole tile of lead
      st.ref_ahead = __ldg(e.refs + (int64_t)(st.n_ntile * kGroups + grp) * e.ld_refs + st.n_trow);
      if (e.run_rows && st.n_mtile == tc.m_tile) {
        // contiguous-run schedule: the tile after next continues this thread's row -> pull its granule towards L2
        const int col2 = (st.n_ntile + 1) * BN + grp * kColsW;
        if (col2 < p.N) {
          prefetch_l2(qnext + BN);
          if (kColsW > 64) prefetch_l2(qnext + BN + 64);
        }
      }
      // packed fp32 pairs through the sm_100 float2 intrinsics (FFMA2 / FADD2 / FMUL2): the register allocator pairs
      // the values itself - hand-packed 64-bit asm operands cost ~5 IMAD.MOV per pair in this loop
      const float2 as2 = make_float2(e.as2, e.as2), nl2 = make_float2(st.nl, st.nl), cw2 = make_float2(st.cwt, st.cwt),
                   nsc2 = make_float2(nsc, nsc);
      float2 lacc = make_float2(0.f, 0.f);
      uint8_t* wbase = smem + epi_warp * (kBufs * 4096);
      uint8_t* wbuf = wbase;
      WarpStage stg;
      uint32_t sa[16], sb[16];
      sm100::tmem_ld16_nowait(ts, sa);
      sm100::tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t (&cur)[16] = (c & 1) ? sb : sa;
        uint32_t (&nxt)[16] = (c & 1) ? sa : sb;
        if (c + 1 < kChunks) sm100::tmem_ld16_nowait(ts + (c + 1) * 16, nxt);
        sm100::pin16(cur);
        uint32_t (&qv)[8] = st.qb[c % kSlots];
        const float* cb = buf + c * 16;
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = lds128(cb + j);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 s2 = make_float2(__uint_as_float(cur[j + 2 * h]), __uint_as_float(cur[j + 2 * h + 1]));
            const float2 c2 = h ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y);
            const float2 u2 = __fadd2_rn(__ffma2_rn(s2, as2, c2), nl2);
            const float2 pp = make_float2(fast_ex2(u2.x), fast_ex2(u2.y));
            const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qv[j / 2 + h]));
            const float2 nq = __fmul2_rn(qf, nsc2);
            const float2 g2 = __ffma2_rn(cw2, pp, nq);
            lacc = __ffma2_rn(nq, u2, lacc);
            __nv_bfloat162 hh = __floats2bfloat162_rn(g2.x, g2.y);
            packed[j / 2 + h] = *reinterpret_cast<uint32_t*>(&hh);
          }
        }
        // refill the slot just drained: with a later chunk of this tile, or - once those are all in flight - with
        // chunk c + kSlots - kChunks of the next tile
        if (c + kSlots < kChunks) ldq(qrow + (c + kSlots) * 16, qv, pol);
        else ldq(qnext + (c + kSlots - kChunks) * 16, qv, pol);
        if ((c & 3) == 0) {   // the staging buffer about to be refilled must no longer be read by the previous store
          wbuf = wbase + st.flip * 4096;
          if (lane == 0) sm100::tma_store_wait_read<kBufs - 1>();
          __syncwarp();
          stg.init(wbuf, lane);
        }
        stg.put((c & 3) * 2, packed[0], packed[1], packed[2], packed[3]);
        stg.put((c & 3) * 2 + 1, packed[4], packed[5], packed[6], packed[7]);
        if ((c & 3) == 3) {
          sm100::fence_proxy_async_smem();
          __syncwarp();
          const int col0 = tc.n_tile * BN + grp * kColsW + (c >> 2) * 64;
          if (lane == 0) {
#if DINOX_EXP_G2 & 2   // experiment: G is staged but never stored
            if (false) {}
#elif DINOX_STREAM_EVICT_FIRST
            if (row0 < p.M && col0 < p.N) sm100::tma_store_3d_hint(tmC, wbuf, col0, row0, 0, pol);
#else
            if (row0 < p.M && col0 < p.N) sm100::tma_store_3d(tmC, wbuf, col0, row0, 0);
#endif
            sm100::tma_store_commit();   // always: wait_read<kBufs-1> counts groups
          }
          if (e.db2_partial && !(DINOX_EXP_G2 & 4)) {
            // column sums of the staged 32 x 64 tile: lane l owns columns 2l, 2l+1; row r of the tile is one
            // conflict-free 128-byte wavefront (16-byte piece j of row r lives at ((j ^ (r & 7)) << 4))
            const uint32_t cbase = stg.base + (lane & 3) * 4;
            const uint32_t piece = (uint32_t)lane >> 2;
            float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const uint32_t v = lds32(cbase + r * 128 + ((piece ^ (uint32_t)(r & 7)) << 4));
              d2 = __fadd2_rn(d2, make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)));
            }
            const int col = col0 + 2 * lane;
            if (row0 < p.M && col < p.N) {
              const float d0 = d2.x, d1 = d2.y;
              float* o = e.db2_partial + (int64_t)(row0 >> 5) * p.N + col;
              if (col + 1 < p.N && (p.N & 1) == 0) *reinterpret_cast<float2*>(o) = make_float2(d0, d1);
              else { o[0] = d0; if (col + 1 < p.N) o[1] = d1; }
            }
          }
          if (kBufs > 1) st.flip ^= 1;
        }
        if (c + 1 < kChunks) sm100::tmem_wait_ld();
      }
      st.ahead = true;
      if (st.in_b) st.loss_b -= lacc.x + lacc.y; else st.loss_a -= lacc.x + lacc.y;   // the accumulator holds -cwt q u
    }
  };
};
using EpiGradR = EpiGradRT<DINOX_RB_EPI_WARPS>;
template <bool kRun> using EpiTeachQ = EpiTeachQT<kRun, DINOX_RB_EPI_WARPS>;

// =============================================================================================
// Epilogue 3 (pass 2, "transposed"): TMEM lanes = prototypes k, columns = entries e.
// Sub-GEMM 0 = student logits  S[k,e] = W2s[k,:] . Hs[e,:],  sub-GEMM 1 = teacher T[k,e].
//   u = S*as2 - lse2[e] + cs2[k]            log2 of the student softmax prob,  p = 2^u
//   w = T*at2 - rb2[e]  + ct2[k]            log2 of the teacher prob,          q = 2^w
//   G[k,e]  = cw[e]/tau_s * (p - q)                       -> bf16, Gt (K, ldg)
//   loss   -= cw[e] * q * ln2 * u                         (-q ln p)
//   db2[k] += G[k,e]                                      (fp32, before rounding)
// Per-entry constants (-lse2, -rb2, cw/tau_s; zero weight for padding / out-of-range entries) are
// staged in shared memory once per tile.  Issue budget per element: 2 FFMA + 2 FADD + 2 MUFU
// (logits -> probs), FMUL + 2 FFMA + FADD (gradient, loss, bias grad), 1/2 F2FP, 3/4 LDS.128.
// =============================================================================================
#ifndef DINOX_G_EVICT_FIRST
#define DINOX_G_EVICT_FIRST 0
#endif
struct EpiGradT {
  static constexpr bool kUsesTmaStore = true;
  static constexpr int kEpiWarps = DINOX_EPI_WARPS;
  static constexpr int kGroups = kEpiWarps / 4;        // entry-column groups per TMEM lane quarter
  static constexpr int kCols = 128 / kGroups;          // entries per warp and tile (64 or 32)
  static constexpr int kRowBytes = kCols * 2;          // staging row: 128 B (SWIZZLE_128B) or 64 B (SWIZZLE_64B)
  static constexpr int kWarpBuf = 32 * kRowBytes;      // 4 KB or 2 KB per warp
// one staging buffer per warp leaves room for one more operand stage (5 instead of 4 without CTA pairs:
// -3.5 %; neutral with pairs); the store of tile i has long drained when tile i+1 reaches its first STS
#ifndef DINOX_GRAD_STAGING_BUFS
#define DINOX_GRAD_STAGING_BUFS 1
#endif
  static constexpr int kBufs = DINOX_GRAD_STAGING_BUFS; // staging buffers per warp (TMA store source)
  // Every epilogue warp is self-contained: its own staging buffers and its own copy of the per-entry
  // constants of its kCols entries (each warp loads them itself: a few hundred redundant bytes per
  // tile instead of two CTA-wide named barriers per tile that made every warp wait for the slowest).
  static constexpr int kStageBytes = kEpiWarps * kBufs * kWarpBuf;
  static constexpr int kConstFloats = 4 * kCols;       // per warp: -lse2 | -rb2 | cw/tau_s | -cw/tau_s
  static constexpr int kEpiSmemBytes = kStageBytes + kEpiWarps * kConstFloats * 4;
  struct Params {
    float as2, at2, inv_tau_s;
    const float* cs2;       // (K)
    const float* ct2;       // (K) for entries < alt_from
    const float* ct2_alt;   // (K) for entries >= alt_from (iBOT patch centre); may be NULL
    int alt_from;           // multiple of BN
    const float* lse2;      // (E) student LSE (log2) per entry
    const float* rb2;       // (E) teacher row bias (log2) per entry
    const float* cw;        // (E) entry weight (norm * group weight), 0 for padding
    float* db2_partial;     // (kGroups*num_n_tiles, M) or NULL
    float* loss_partial;    // (gridDim.x * kEpiWarps * 2): per CTA and epilogue warp, {entries < alt_from, >= alt_from}
  };
  struct State {
    float loss_a = 0.f, loss_b = 0.f;
    // raw prefetched values of the NEXT tile (no arithmetic on them until prologue(): a dependent
    // instruction right after the load would stall the warp for the whole global-memory latency)
    float nl[kCols / 32], nr[kCols / 32], cw[kCols / 32];   // this lane's entries of the warp's column group
    float cs, ct;            // per-prototype offsets of this thread's TMEM lane
    float cs_cur, ct_cur;    // ... of the tile being processed (latched by prologue)
    int flip = 0;
  };
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams&, int epi_warp, int lane, State& st) {
    const float a = warp_sum(st.loss_a), b = warp_sum(st.loss_b);
    const float sc = -DINOX_LN2 / e.inv_tau_s;   // the tiles accumulate (cw/tau_s) * q * u
    if (lane == 0) {
      e.loss_partial[((int64_t)blockIdx.x * kEpiWarps + epi_warp) * 2 + 0] = a * sc;
      e.loss_partial[((int64_t)blockIdx.x * kEpiWarps + epi_warp) * 2 + 1] = b * sc;
      sm100::tma_store_wait_all<0>();
    }
  }
  // this thread's staging row: 16-byte piece j lives at row*kRowBytes + ((j ^ swizzle(row)) << 4)
  struct Stage {
    uint32_t row, sw;
    __device__ __forceinline__ void init(uint8_t* buf, int lane) {
      row = sm100::smem_u32(buf) + lane * kRowBytes;
      sw = kRowBytes == 128 ? ((lane & 7) << 4) : (((lane >> 1) & 3) << 4);
    }
    __device__ __forceinline__ void put(int piece, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + ((piece << 4) ^ sw)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    }
  };
  template <int BN>
  struct Impl {
    static __device__ __forceinline__ void fetch(const Params& e, const CoreParams& p, TileCoord tc, int epi_warp, int lane,
                                                 State& st) {
      const int ent0 = tc.n_tile * BN + (epi_warp >> 2) * kCols;
#pragma unroll
      for (int j = 0; j < kCols / 32; ++j) {
        const int ent = ent0 + j * 32 + lane;
        const bool ok = ent < p.N;
        st.nl[j] = ok ? __ldg(e.lse2 + ent) : 0.f;
        st.nr[j] = ok ? __ldg(e.rb2 + ent) : 0.f;
        st.cw[j] = ok ? __ldg(e.cw + ent) : 0.f;
      }
      const int k = tc.m_tile * BM + epi_quarter() * 32 + lane;
      const bool alt = e.ct2_alt && tc.n_tile * BN >= e.alt_from;
      st.cs = k < p.M ? __ldg(e.cs2 + k) : 0.f;
      st.ct = k < p.M ? __ldg((alt ? e.ct2_alt : e.ct2) + k) : 0.f;
    }
    static __device__ __forceinline__ void prologue(const Params& e, const CoreParams&, TileCoord, int, int epi_warp, int lane,
                                                    uint8_t* smem, State& st) {
      float* buf = reinterpret_cast<float*>(smem + kStageBytes) + epi_warp * kConstFloats;
      __syncwarp();   // every lane is done reading the previous tile's constants
#pragma unroll
      for (int j = 0; j < kCols / 32; ++j) {
        buf[j * 32 + lane] = -st.nl[j];
        buf[kCols + j * 32 + lane] = -st.nr[j];
        buf[2 * kCols + j * 32 + lane] = st.cw[j] * e.inv_tau_s;
        buf[3 * kCols + j * 32 + lane] = -(st.cw[j] * e.inv_tau_s);
      }
      st.cs_cur = st.cs; st.ct_cur = st.ct;
      __syncwarp();
    }

    // 16 entries of one prototype row: logits -> (p, q) -> gradient, loss and bias-gradient terms;
    // the bf16 gradients go to pieces 2*c16, 2*c16+1 of this thread's staging row.
    // The fp32 arithmetic runs on packed pairs (Blackwell fma/add/mul.f32x2: two IEEE operations per
    // issued instruction - same results as the scalar form, ~40 % fewer issue slots):
    //   u = S*as2 + nl + cs ; w = T*at2 + nr + ct ; p = 2^u ; q = 2^w
    //   nq = (-cw)*q ; g = cw*p + nq ; loss_acc += nq*u (= -cw q u) ; db2_acc += g
    static __device__ __forceinline__ void chunk16(const Params& e, const uint32_t (&sr)[16], const uint32_t (&tr)[16],
                                                   const float* cb, float cs, float ct, const Stage& stg, int c16,
                                                   uint64_t& lacc, uint64_t& dacc) {
      uint32_t packed[8];
      const uint64_t as2 = pack2(e.as2, e.as2), at2 = pack2(e.at2, e.at2), cs2 = pack2(cs, cs), ct2 = pack2(ct, ct);
#if DINOX_EXP_EPI_MODE == 1   // experiment: TMEM loads only, trivial math
      {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) a += __uint_as_float(sr[j]) + __uint_as_float(tr[j]);
        lacc = add2(lacc, pack2(a, 0.f));
        return;
      }
#endif
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 nl = lds128(cb + j), nr = lds128(cb + kCols + j), cw = lds128(cb + 2 * kCols + j), ncw = lds128(cb + 3 * kCols + j);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t s2 = pack2u(sr[j + 2 * h], sr[j + 2 * h + 1]), t2 = pack2u(tr[j + 2 * h], tr[j + 2 * h + 1]);
          const uint64_t nl2 = h ? pack2(nl.z, nl.w) : pack2(nl.x, nl.y), nr2 = h ? pack2(nr.z, nr.w) : pack2(nr.x, nr.y);
          const uint64_t c2 = h ? pack2(cw.z, cw.w) : pack2(cw.x, cw.y), n2 = h ? pack2(ncw.z, ncw.w) : pack2(ncw.x, ncw.y);
          const uint64_t u2 = add2(fma2(s2, as2, nl2), cs2);
          const uint64_t w2 = add2(fma2(t2, at2, nr2), ct2);
          float ux, uy, wx, wy;
          unpack2(u2, ux, uy); unpack2(w2, wx, wy);
          const uint64_t pp = pack2(fast_ex2(ux), fast_ex2(uy)), qq = pack2(fast_ex2(wx), fast_ex2(wy));
          const uint64_t nq = mul2(n2, qq);
          const uint64_t g2 = fma2(c2, pp, nq);
          lacc = fma2(nq, u2, lacc);
          dacc = add2(dacc, g2);
          float gx, gy;
          unpack2(g2, gx, gy);
          __nv_bfloat162 hh = __floats2bfloat162_rn(gx, gy);
          packed[j / 2 + h] = *reinterpret_cast<uint32_t*>(&hh);
        }
      }
      stg.put(2 * c16, packed[0], packed[1], packed[2], packed[3]);
      stg.put(2 * c16 + 1, packed[4], packed[5], packed[6], packed[7]);
    }

    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap* tmC,
                                                uint32_t tmem_acc, int acc_stage, int epi_warp, int lane, uint8_t* smem,
                                                State& st) {
      static_assert(BN == 128, "EpiGradT is written for 128-entry tiles");
      const float* cb = reinterpret_cast<const float*>(smem + kStageBytes) + epi_warp * kConstFloats;
      const int q = epi_quarter();
      const int grp = epi_warp >> 2;
      const int k0 = tc.m_tile * BM + q * 32;
      const int k = k0 + lane;   // prototype
      const bool kok = k < p.M;
      const bool alt = e.ct2_alt && tc.n_tile * BN >= e.alt_from;
      const float cs = st.cs_cur, ct = st.ct_cur;
      const uint32_t ts = tmem_acc + ((uint32_t)(q * 32) << 16) + grp * kCols;
      const uint32_t tt = ts + BN;
      uint8_t* wbuf = smem + (epi_warp * kBufs + st.flip) * kWarpBuf;
      Stage stg;
      stg.init(wbuf, lane);
      uint64_t lacc = pack2(0.f, 0.f), dacc = pack2(0.f, 0.f);   // packed (even, odd) partial sums
      uint32_t sa[16], ta[16], sb[16], tb[16];
      sm100::tmem_ld16_nowait(ts, sa);
      sm100::tmem_ld16_nowait(tt, ta);
      // the staging buffer about to be refilled must no longer be read by the store issued kBufs tiles ago
      if (lane == 0) sm100::tma_store_wait_read<kBufs - 1>();
      __syncwarp();
      sm100::tmem_wait_ld();
#if DINOX_EXP_EPI_MODE == 2   // experiment: math only (one TMEM load per tile, registers reused)
      chunk16(e, sa, ta, cb, cs, ct, stg, 0, lacc, dacc);
      sm100::pin16(sa); sm100::pin16(ta);
      chunk16(e, sa, ta, cb + 16, cs, ct, stg, 1, lacc, dacc);
      sm100::pin16(sa); sm100::pin16(ta);
      chunk16(e, sa, ta, cb + 32, cs, ct, stg, 2, lacc, dacc);
      sm100::pin16(sa); sm100::pin16(ta);
      chunk16(e, sa, ta, cb + 48, cs, ct, stg, 3, lacc, dacc);
      if (false) {
#else
      {
#endif
      sm100::tmem_ld16_nowait(ts + 16, sb);
      sm100::tmem_ld16_nowait(tt + 16, tb);
      sm100::pin16(sa); sm100::pin16(ta);
      chunk16(e, sa, ta, cb, cs, ct, stg, 0, lacc, dacc);
      sm100::tmem_wait_ld();
      if (kCols == 64) {
        sm100::tmem_ld16_nowait(ts + 32, sa);
        sm100::tmem_ld16_nowait(tt + 32, ta);
      }
      sm100::pin16(sb); sm100::pin16(tb);
      chunk16(e, sb, tb, cb + 16, cs, ct, stg, 1, lacc, dacc);
      if (kCols == 64) {
        sm100::tmem_wait_ld();
        sm100::tmem_ld16_nowait(ts + 48, sb);
        sm100::tmem_ld16_nowait(tt + 48, tb);
        sm100::pin16(sa); sm100::pin16(ta);
        chunk16(e, sa, ta, cb + 32, cs, ct, stg, 2, lacc, dacc);
        sm100::tmem_wait_ld();
        sm100::pin16(sb); sm100::pin16(tb);
        chunk16(e, sb, tb, cb + 48, cs, ct, stg, 3, lacc, dacc);
      }
      }
      // G tile rows [k0, k0+32) x entries [ent0, ent0+kCols) leave as one TMA store (clipped at K and E)
      sm100::fence_proxy_async_smem();
      __syncwarp();
      const int ent0 = tc.n_tile * BN + grp * kCols;
      if (lane == 0) {
#if DINOX_G_EVICT_FIRST
        if (k0 < p.M && ent0 < p.N) sm100::tma_store_3d_hint(tmC, wbuf, ent0, k0, 0, sm100::l2_policy_evict_first());
#else
        if (k0 < p.M && ent0 < p.N) sm100::tma_store_3d(tmC, wbuf, ent0, k0, 0);
#endif
        sm100::tma_store_commit();   // always: wait_read<kBufs-1> counts groups
      }
      if (kBufs > 1) st.flip ^= 1;
      if (kok) {
        float l0, l1, d0, d1;
        unpack2(lacc, l0, l1); unpack2(dacc, d0, d1);
        if (e.db2_partial) e.db2_partial[(int64_t)(tc.n_tile * kGroups + grp) * p.M + k] = d0 + d1;
        if (alt) st.loss_b -= l0 + l1; else st.loss_a -= l0 + l1;   // the accumulators hold -cw q u
      }
    }
  };
};

// =============================================================================================
// Epilogue 4 (Gram anchoring): per image, sub-GEMM 0 = student Gram tile, sub-GEMM 1 = teacher
// Gram tile (tokens already L2-normalised, bf16).  delta = Gs - Gt never leaves the SM in fp32:
//   loss_partial += sum delta^2 ;  delta -> bf16 (B, T', ldd) as the operand of the backward GEMM.
// scripts/phase5_big_run.py:727 (bmm) + :738 (mse_loss) fused.
// =============================================================================================
struct EpiGramDiff {
  static constexpr bool kUsesTmaStore = false;
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiSmemBytes = 0;
  struct Params {
    __nv_bfloat16* delta;   // (B, T', ldd) or NULL (forward only)
    int64_t ldd, batch_stride;
    float* loss_partial;    // (gridDim.x * 8)
  };
  struct State {
    float loss = 0.f;
  };
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams&, int epi_warp, int lane, State& st) {
    const float a = warp_sum(st.loss);
    if (lane == 0) e.loss_partial[(int64_t)blockIdx.x * 8 + epi_warp] = a;
  }
  template <int BN>
  struct Impl {
    static __device__ __forceinline__ void fetch(const Params&, const CoreParams&, TileCoord, int, int, State&) {}
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int, int, uint8_t*, State&) {}
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap*,
                                                uint32_t tmem_acc, int, int epi_warp, int lane, uint8_t*, State& st) {
      static_assert(BN == 128, "EpiGramDiff is written for 128-wide tiles");
      const int q = epi_quarter();
      const int half = epi_warp >> 2;
      const int i = tc.m_tile * BM + q * 32 + lane;
      const bool iok = i < p.M;
      const uint32_t ts = tmem_acc + ((uint32_t)(q * 32) << 16) + half * 64;
      const uint32_t tt = ts + BN;
      float loss = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float sv[32], tv[32];
        sm100::tmem_ld32x2(ts + c * 32, tt + c * 32, sv, tv);
        const int j0 = tc.n_tile * BN + half * 64 + c * 32;
        if (j0 >= p.N) break;
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = (iok && j0 + j < p.N) ? (sv[j] - tv[j]) : 0.f;
          loss = fmaf(d, d, loss);
          sv[j] = d;
        }
        if (e.delta && iok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(sv[2 * j], sv[2 * j + 1]);
            packed[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          __nv_bfloat16* o = e.delta + tc.batch * e.batch_stride + (int64_t)i * e.ldd + j0;
          if (j0 + 32 <= e.ldd) {  // padding columns [N, ldd) receive zeros
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (j0 + j < e.ldd) o[j] = __float2bfloat16_rn(sv[j]);
          }
        }
      }
      st.loss += loss;
    }
  };
};

// per-CTA loss partials of pass 2, interleaved {a, b}: out[0] (+)= sum a, out[1] (+)= sum b (fixed order)
__global__ void __launch_bounds__(1024) pair_sum_kernel(const float* __restrict__ x, int64_t n_pairs,
                                                         float* __restrict__ out, int accumulate, int write_total = 0) {
  __shared__ float red[64];
  float a = 0.f, b = 0.f;
  for (int64_t i = threadIdx.x; i < n_pairs; i += 1024) { a += x[2 * i]; b += x[2 * i + 1]; }
  a = block_sum<1024>(a, red);
  b = block_sum<1024>(b, red);
  if (threadIdx.x == 0) {
    out[0] = a + (accumulate ? out[0] : 0.f);
    out[1] = b + (accumulate ? out[1] : 0.f);
    if (write_total) out[2] = out[0] + out[1];
  }
}

// single CTA deterministic sum of n floats (optionally scaled) into *out
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n, float scale, float* __restrict__ out,
                                                    int accumulate) {
  __shared__ float red[64];
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) a += x[i];
  a = block_sum<1024>(a, red);
  if (threadIdx.x == 0) *out = a * scale + (accumulate ? *out : 0.f);
}

// =============================================================================================
// kernels
// =============================================================================================
template <int BN, class Epi>
struct EpiAdapter {
  static constexpr bool kUsesTmaStore = Epi::kUsesTmaStore;
  static constexpr int kEpiWarps = Epi::kEpiWarps;
  static constexpr int kEpiSmemBytes = Epi::kEpiSmemBytes;
  using Params = typename Epi::Params;
  using State = typename Epi::State;
  static __device__ __forceinline__ void fetch(const Params& e, const CoreParams& p, TileCoord tc, int w, int l, State& st) {
    Epi::template Impl<BN>::fetch(e, p, tc, w, l, st);
  }
  static __device__ __forceinline__ void prologue(const Params& e, const CoreParams& p, TileCoord tc, int a, int w, int l,
                                                  uint8_t* s, State& st) {
    Epi::template Impl<BN>::prologue(e, p, tc, a, w, l, s, st);
  }
  static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, const CUtensorMap* tmC,
                                              uint32_t tm, int a, int w, int l, uint8_t* s, State& st) {
    Epi::template Impl<BN>::tile(e, p, tc, tmC, tm, a, w, l, s, st);
  }
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams& p, int w, int l, State& st) {
    Epi::finish(e, p, w, l, st);
  }
};

// NOUT output tensor maps (1 everywhere except the reduce-scatter GEMM: one map per data-parallel rank's shard)
constexpr int kMaxOwners = 8;
template <int NOUT>
struct alignas(64) OutMaps {
  CUtensorMap m[NOUT];
};

template <int BN, int NSPLIT, int NSUB, int CL, class Epi, int RES = kResNone, int NOUT = 1>
__global__ void __launch_bounds__((2 + Epi::kEpiWarps) * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
            const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
            const __grid_constant__ OutMaps<NOUT> tmC, const CoreParams p, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  gemm_body<BN, NSPLIT, NSUB, CL, EpiAdapter<BN, Epi>, RES>(p, ep, &tmA0, &tmB0, &tmA1, &tmB1, &tmC.m[0], smem_raw);
}

static int env_flag_early(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// number of CTAs for `tiles` super tiles of a CL-cluster kernel: one CTA per SM, whole clusters
static inline int launch_grid(int64_t super_tiles, int cl) {
  int64_t clusters = num_sms() / cl;
  if (clusters > super_tiles) clusters = super_tiles;
  if (clusters < 1) clusters = 1;
  return (int)(clusters * cl);
}

struct Operand {
  const void* ptr;
  int64_t rows;      // M (or N) extent
  int64_t ld;        // leading dimension in elements of the stored matrix
  int mn_major;      // 0: stored (rows, K) ; 1: stored (K, rows)
  int64_t batch_stride = 0;  // elements between consecutive problems (batched launches)
};

// output of the TMA-store epilogues: (slabs, M, N) with leading dimension ld and slab stride
// Ordered split for a GEMM of `tiles` super tiles on `ncl` clusters with `kblocks` k-blocks per tile: which tiles are cut
// (from `first` on) into how many parts.  Cost model in units of one whole tile: waves of equal items; every extra part
// of a tile costs one more epilogue pass over the output (`kPartCost`).  Returns parts = 0 when nothing beats whole tiles.
struct OrderedSplit { int first = 0, parts = 0; };
static OrderedSplit plan_ordered_split(int64_t tiles, int64_t ncl, int64_t kblocks) {
  OrderedSplit best;
  if (tiles <= 0 || ncl <= 1 || tiles % ncl == 0) return best;
  constexpr double kPartCost = 0.004;
  // parts of one tile that land in the same round run side by side and their epilogues hand over one after the
  // other: an item must be long (>= 32 k-blocks of MMA work) next to an epilogue pass for that chain to stay hidden
  constexpr int kMinKb = 32, kMaxParts = 16;
  double best_cost = (double)((tiles + ncl - 1) / ncl);
  const int64_t full = tiles / ncl * ncl, left = tiles - full;
  for (int variant = 0; variant < 2; ++variant) {
    // 0: only the tiles of the last, partial wave are cut; 1: every tile is cut
    const int64_t first = variant == 0 ? full : 0, cut = tiles - first;
    if (variant == 0 && full == 0) continue;
    for (int parts = 2; parts <= kMaxParts; ++parts) {
      if ((kblocks + parts - 1) / parts < kMinKb) break;
      if (((kblocks + parts - 1) / parts) * (parts - 1) >= kblocks) continue;   // an empty last part
      const double waves = (double)((cut * parts + ncl - 1) / ncl) / parts;
      const double cost = (double)(first / ncl) + waves + kPartCost * parts * ((double)cut / tiles);
      if (cost < best_cost - 1e-9) { best_cost = cost; best.first = (int)first; best.parts = parts; }
    }
  }
  (void)left;
  return best;
}

struct OutDesc {
  void* ptr = nullptr;
  int is_bf16 = 0;
  int64_t ld = 0, slab_stride = 0, slabs = 1;
  int row_bytes = 128;   // staging row of the epilogue: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  int64_t cols = 0;      // column extent of the map when it differs from N (padded rows), 0 = N
  // sharded output (reduce-scatter GEMM): owners > 0 buffers of rows_per_owner rows each; ptr is ignored
  int owners = 0;
  int64_t rows_per_owner = 0;
  void* owner_ptr[8] = {};
  // ordered split allowed (plain fp32 TMA-store GEMMs): the zeroed counters the epilogue's order_flags point to hold
  // at least order_flag_count entries (4 per 128-row tile of the output)
  int64_t order_flag_count = 0;
};

static int make_operand_tmap(CUtensorMap* tm, const Operand& o, int64_t K, int tile_rows, int64_t batches, const char* what) {
  if (batches > 1) {
    if (!o.mn_major) return make_tmap_bf16_3d(tm, o.ptr, batches, o.rows, K, o.ld, o.batch_stride, tile_rows, what);
    return make_tmap_bf16_3d(tm, o.ptr, batches, K, o.rows, o.ld, o.batch_stride, BK, what);
  }
  if (!o.mn_major) return make_tmap_bf16_2d(tm, o.ptr, o.rows, K, o.ld, tile_rows, what);
  return make_tmap_bf16_2d(tm, o.ptr, K, o.rows, o.ld, BK, what);  // box = 64 k-rows x 64 mn-elements
}

template <int BN, int NSPLIT, int NSUB, int CL, class Epi, int RES = kResNone, int NOUT = 1>
static int launch(const Operand& a0, const Operand& b0, const Operand* a1, const Operand* b1, int64_t M, int64_t N,
                  int64_t K, int m_fastest, const typename Epi::Params& ep, const OutDesc& od, cudaStream_t stream,
                  const char* name, int64_t batches = 1, int64_t splits = 1) {
  DINOX_REQUIRE(M > 0 && N > 0 && K > 0, DINOX_E_BADARG, "%s: empty problem", name);
  DINOX_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), DINOX_E_BADARG, "%s: dimension too large", name);
  DINOX_REQUIRE(batches >= 1 && batches < (1 << 20), DINOX_E_BADARG, "%s: bad batch count", name);
  DINOX_REQUIRE(splits >= 1 && (splits == 1 || batches == 1), DINOX_E_BADARG, "%s: split-K cannot be batched", name);
  DINOX_REQUIRE(!(b0.mn_major && ((BN / NSPLIT / CL) % 64) != 0), DINOX_E_UNSUPPORTED,
                "%s: MN-major B needs 64-element atoms per CTA at this tile shape", name);
  CUtensorMap tA0, tB0, tA1, tB1;
  OutMaps<NOUT> tC;
  int rc;
  if ((rc = make_operand_tmap(&tA0, a0, K, BM, batches, "A"))) return rc;
  if ((rc = make_operand_tmap(&tB0, b0, K, BN / NSPLIT / CL, batches, "B"))) return rc;   // one box per (N sub-tile, CTA)
  tA1 = tA0; tB1 = tB0;
  if (NSUB == 2) {
    DINOX_REQUIRE(a1 && b1 && a1->mn_major == a0.mn_major && b1->mn_major == b0.mn_major, DINOX_E_BADARG,
                  "%s: second operand pair missing or layout mismatch", name);
    if ((rc = make_operand_tmap(&tA1, *a1, K, BM, batches, "A1"))) return rc;
    if ((rc = make_operand_tmap(&tB1, *b1, K, BN / NSPLIT / CL, batches, "B1"))) return rc;
  }
  for (int i = 0; i < NOUT; ++i) tC.m[i] = tA0;
  if (NOUT > 1) {
    DINOX_REQUIRE(Epi::kUsesTmaStore && od.owners >= 1 && od.owners <= NOUT && od.rows_per_owner > 0 &&
                      od.rows_per_owner % BM == 0 && od.rows_per_owner * od.owners >= M && splits == 1 && batches == 1,
                  DINOX_E_BADARG, "%s: sharded output needs 1..%d owners with a multiple of %d rows each", name, NOUT, BM);
    for (int i = 0; i < od.owners; ++i) {
      DINOX_REQUIRE(od.owner_ptr[i], DINOX_E_BADARG, "%s: null shard pointer", name);
      if ((rc = make_tmap_out_3d(&tC.m[i], od.owner_ptr[i], od.is_bf16 != 0, 1, od.rows_per_owner, N, od.ld,
                                 od.rows_per_owner * od.ld, od.row_bytes, "C shard"))) return rc;
    }
  } else if (Epi::kUsesTmaStore) {
    DINOX_REQUIRE(od.ptr, DINOX_E_BADARG, "%s: output descriptor missing", name);
    if ((rc = make_tmap_out_3d(&tC.m[0], od.ptr, od.is_bf16 != 0, od.slabs, M, od.cols ? od.cols : N, od.ld, od.slab_stride, od.row_bytes, "C"))) return rc;
  }
  CoreParams p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.num_m_tiles = (int)((M + BM - 1) / BM);
  p.num_n_tiles = (int)((N + BN - 1) / BN);
  p.num_k_blocks = (int)((K + BK - 1) / BK);
  p.a_mn_major = a0.mn_major; p.b_mn_major = b0.mn_major; p.m_fastest = m_fastest;
  p.batches = (int)batches;
  p.splits = (int)splits;
  p.kb_per_split = (int)((p.num_k_blocks + splits - 1) / splits);
  p.rows_per_owner = NOUT > 1 ? (int)od.rows_per_owner : 0;
  int64_t os_items = 0;
  if (od.order_flag_count > 0 && splits == 1 && batches == 1 && RES == kResNone && NOUT == 1 && NSUB == 1) {
    const int64_t tiles = (int64_t)((p.num_m_tiles + CL - 1) / CL) * p.num_n_tiles;
    const OrderedSplit os = plan_ordered_split(tiles, num_sms() / CL, p.num_k_blocks);
    if (os.parts >= 2 && (int64_t)p.num_m_tiles * p.num_n_tiles * 4 <= od.order_flag_count) {
      p.os_first = os.first; p.os_parts = os.parts; p.os_kb = (p.num_k_blocks + os.parts - 1) / os.parts;
      os_items = os.first + (tiles - os.first) * os.parts;
    }
  }
  DINOX_REQUIRE(splits == 1 || (int64_t)p.kb_per_split * (splits - 1) < p.num_k_blocks, DINOX_E_BADARG,
                "%s: %lld splits leave an empty K range", name, (long long)splits);
  DINOX_REQUIRE(RES != kResA || (p.num_k_blocks <= kResKBlocks && !a0.mn_major && splits == 1), DINOX_E_UNSUPPORTED,
                "%s: resident-A mode needs a K-major A operand with K <= %d", name, kResKBlocks * BK);
  DINOX_REQUIRE(RES != kResB || (p.num_k_blocks <= kResKBlocks && !b0.mn_major && splits == 1 && batches == 1),
                DINOX_E_UNSUPPORTED, "%s: resident-B mode needs a K-major B operand with K <= %d", name, kResKBlocks * BK);
  auto kern = gemm_kernel<BN, NSPLIT, NSUB, CL, Epi, RES, NOUT>;
  constexpr int smem = smem_bytes<BN, CL, Epi, RES, NSUB>();
  static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
  // the opt-in shared-memory size is a per-device function attribute
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    DINOX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t super = os_items > 0 ? os_items
                                    : (int64_t)((p.num_m_tiles + CL - 1) / CL) * p.num_n_tiles * (splits > 1 ? splits : batches);
  DINOX_REQUIRE(super < (1ll << 31), DINOX_E_BADARG, "%s: too many tiles", name);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)launch_grid(super, CL));
  cfg.blockDim = dim3((2 + Epi::kEpiWarps) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int pdl = env_flag_early("DINOX_PDL", 0);
  if (pdl) {   // let the CTAs start their set-up while the previous kernel of the stream drains
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tA0, tB0, tA1, tB1, tC, p, ep);
  if (e != cudaSuccess) {
    set_error("%s: cudaLaunchKernelEx failed: %s", name, cudaGetErrorString(e));
    return DINOX_E_CUDA;
  }
  return check_launch(name, stream);
}

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// CTA-pair (cta_group::2) mode per kernel family, DINOX_PAIR bit mask for A/B measurements (default: all on):
//   1 = plain store GEMMs   dW2 0.44 -> 0.34 ms, dH 0.43 -> 0.31 ms, layer 1: halved B-operand fill per SM
//   2 = pass 1 (head_stats) 0.303 -> 0.288 ms      4 = pass 2 (head_grad) 0.942 -> 0.911 ms
enum { kPairStore = 1, kPairStats = 2, kPairGrad = 4 };
static bool pair_enabled(int family) {
  static int v = -1;
  if (v < 0) v = env_flag("DINOX_PAIR", kPairStore | kPairStats | kPairGrad);
  return (v & family) != 0;
}
// Resident-A mode (DINOX_RESA bit mask, same family bits): a CTA walks a contiguous run of tiles that share
// the M tile and keeps that tile's A rows (all k-blocks, K <= 384) in shared memory, so only B streams.
#ifndef DINOX_RESA_DEFAULT
#define DINOX_RESA_DEFAULT 2   /* pass 1: 0.286 -> 0.272 ms; pass 2 (bit 4) loses operand stages: 0.89 -> 1.08 ms */
#endif
static bool resa_enabled(int family, int64_t K) {
  static int v = -1;
  if (v < 0) v = env_flag("DINOX_RESA", DINOX_RESA_DEFAULT);
  return (v & family) != 0 && K <= kResKBlocks * BK;
}

// Schedule of the two read-back kernels (DINOX_RB_SCHED bit mask: 1 = teacher pass, 2 = pass 2): bit set = W2 tile
// resident + column sweep over the M tiles; clear = rows of an M tile resident + contiguous prototype runs (the pass-1
// schedule).  Measured at C2 (ms): teacher 0.355 (columns) / 0.384 (runs); pass 2 0.654 (columns) / 0.589 (runs).
#ifndef DINOX_RB_SCHED_DEFAULT
#define DINOX_RB_SCHED_DEFAULT 3
#endif
static bool readback_cols(int which, int64_t K) {
  static int v = -1;
  if (v < 0) v = env_flag("DINOX_RB_SCHED", DINOX_RB_SCHED_DEFAULT);
  return (v & which) != 0 && K <= kResKBlocks * BK;
}

#ifndef DINOX_EXP_NSPLIT256
#define DINOX_EXP_NSPLIT256 1   // experiment knob: 2 issues the 256-wide tile as two N=128 MMAs (A read twice from smem)
#endif
// tile-shape dispatch of the plain GEMM: widest tile without N waste; clusters of 2 share the B tile
struct StoreArgs {
  void* out;
  int64_t ldo;
  int out_bf16, accumulate;
  float alpha;
  const float* alpha_dev;
  const float* bias_n;
  int64_t slab_stride;   // output elements between batches / splits
  uint32_t* order_flags = nullptr;   // ordered split allowed: zeroed counters, order_flag_count of them
  int64_t order_flag_count = 0;
};

template <class Epi>
static int launch_store_t(int64_t M, int64_t N, const Operand& a, const Operand& b, int64_t K, int m_fastest,
                          const typename Epi::Params& ep, const OutDesc& od, cudaStream_t stream, int64_t batches,
                          int64_t splits) {
  const bool cl2 = M > BM && pair_enabled(kPairStore);   // a single M tile has nobody to pair with
  if (N % 384 == 0) {
    return cl2 ? launch<384, 3, 1, 2, Epi>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, od, stream, "gemm_bf16<384,pair>", batches, splits)
               : launch<384, 3, 1, 1, Epi>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, od, stream, "gemm_bf16<384>", batches, splits);
  }
  if (N % 256 == 0 || N > 2048) {
    return cl2 ? launch<256, 1, 1, 2, Epi>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, od, stream, "gemm_bf16<256,pair>", batches, splits)
               : launch<256, DINOX_EXP_NSPLIT256, 1, 1, Epi>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, od, stream, "gemm_bf16<256>", batches, splits);
  }
  return cl2 ? launch<128, 1, 1, 2, Epi>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, od, stream, "gemm_bf16<128,pair>", batches, splits)
             : launch<128, 1, 1, 1, Epi>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, od, stream, "gemm_bf16<128>", batches, splits);
}

static int launch_store(int64_t M, int64_t N, const Operand& a, const Operand& b, int64_t K, int m_fastest,
                        const StoreArgs& sa, cudaStream_t stream, int64_t batches, int64_t splits) {
  static int direct = -1;
  if (direct < 0) direct = env_flag("DINOX_DIRECT_STORE", 0);
  if ((sa.out_bf16 && sa.accumulate) || direct) {
    DINOX_REQUIRE(splits == 1, DINOX_E_UNSUPPORTED, "gemm_bf16: split-K needs the TMA-store epilogue");
    EpiStoreDirect::Params ep{sa.out, sa.ldo, sa.out_bf16, sa.accumulate, sa.alpha, sa.alpha_dev, sa.bias_n, sa.slab_stride};
    return launch_store_t<EpiStoreDirect>(M, N, a, b, K, m_fastest, ep, OutDesc{}, stream, batches, 1);
  }
  EpiStore::Params ep{sa.out_bf16, sa.accumulate, sa.alpha, sa.alpha_dev, sa.bias_n};
  OutDesc od;
  od.ptr = sa.out; od.is_bf16 = sa.out_bf16; od.ld = sa.ldo; od.slab_stride = sa.slab_stride;
  od.slabs = splits > 1 ? splits : batches;
  if (sa.order_flags && !sa.out_bf16 && splits == 1 && batches == 1) {
    ep.order_flags = sa.order_flags;
    od.order_flag_count = sa.order_flag_count;
  }
  return launch_store_t<EpiStore>(M, N, a, b, K, m_fastest, ep, od, stream, batches, splits);
}

}  // namespace gemm
}  // namespace dinox

extern "C" {
using namespace dinox;
using namespace dinox::gemm;

static int gemm_common_checks(const void* A, const void* B, void* C, int64_t N, int64_t ldc, int out_dtype, const char* name) {
  DINOX_REQUIRE(A && B && C, DINOX_E_BADARG, "%s: null pointer", name);
  DINOX_REQUIRE(out_dtype == DINOX_F32 || out_dtype == DINOX_BF16, DINOX_E_BADARG, "%s: out dtype must be f32 or bf16", name);
  DINOX_REQUIRE(aligned16(C) && (ldc * (out_dtype == DINOX_F32 ? 4 : 2)) % 16 == 0, DINOX_E_ALIGN,
                "%s: C / ldc not 16-byte aligned", name);
  DINOX_REQUIRE(ldc >= N, DINOX_E_BADARG, "%s: ldc < N", name);
  return require_sm100();
}

int dinox_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                    int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int out_dtype, int accumulate,
                    float alpha, const float* alpha_dev, const float* bias_n, int m_fastest,
                    dinox_stream_t stream) {
  int rc = gemm_common_checks(A, B, C, N, ldc, out_dtype, "gemm_bf16");
  if (rc) return rc;
  Operand a{A, M, lda, a_mn_major ? 1 : 0}, b{B, N, ldb, b_mn_major ? 1 : 0};
  StoreArgs sa{C, ldc, out_dtype == DINOX_BF16, accumulate, alpha, alpha_dev, bias_n, 0};
  return launch_store(M, N, a, b, K, m_fastest, sa, stream, 1, 1);
}

/* host-side views of the balanced schedule (no GPU needed; tests/test_host_logic.py): the plan for a tile / cluster /
 * k-block count, and the item list every cluster walks (TileWalker::init_ordered) */
int dinox_plan_ordered_split(int64_t tiles, int64_t clusters, int64_t kblocks, int* first, int* parts) {
  DINOX_REQUIRE(first && parts, DINOX_E_BADARG, "plan_ordered_split: null pointer");
  const OrderedSplit os = plan_ordered_split(tiles, clusters, kblocks);
  *first = os.first;
  *parts = os.parts;
  return DINOX_OK;
}

int64_t dinox_debug_walk_ordered(int num_m_super, int num_n_tiles, int m_fastest, int os_first, int os_parts, int clusters,
                                 int cl, int32_t* out /* (cap, 6): item, cluster, m_tile of CTA 0, n_tile, kpart, kparts */,
                                 int64_t cap) {
  if (!out || num_m_super <= 0 || num_n_tiles <= 0 || clusters <= 0 || cl <= 0 || os_parts < 2) return -1;
  CoreParams p{};
  p.num_m_tiles = num_m_super * cl; p.num_n_tiles = num_n_tiles; p.m_fastest = m_fastest;
  p.batches = 1; p.splits = 1; p.os_first = os_first; p.os_parts = os_parts;
  int64_t n = 0;
  for (int cid = 0; cid < clusters; ++cid) {
    TileWalker w;
    w.init_ordered(p, num_m_super, cid, clusters);
    for (; w.valid(); w.next()) {
      const TileCoord tc = w.coord(p, cl, 0);
      if (n < cap) {
        int32_t* o = out + n * 6;
        o[0] = w.o_i; o[1] = cid; o[2] = tc.m_tile; o[3] = tc.n_tile; o[4] = tc.kpart; o[5] = tc.kparts;
      }
      ++n;
    }
  }
  return n;
}

/* fp32 GEMM whose tile count does not fill whole waves of the persistent grid: the tiles of the partial wave (or all
 * tiles when there are fewer tiles than clusters) are cut along K into parts that accumulate into C in a FIXED order
 * (part j waits for part j - 1 of its tile through the counters in `flags`), see CoreParams::os_*. */
size_t dinox_gemm_bf16_balanced_workspace_bytes(int64_t M, int64_t N) {
  if (M <= 0 || N <= 0) return 0;
  return (size_t)((M + BM - 1) / BM) * ((N + 127) / 128) * 4 * sizeof(uint32_t);
}

int dinox_gemm_bf16_balanced(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                             int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int accumulate, float alpha,
                             const float* alpha_dev, const float* bias_n, int m_fastest, void* flags,
                             dinox_stream_t stream) {
  int rc = gemm_common_checks(A, B, C, N, ldc, DINOX_F32, "gemm_bf16_balanced");
  if (rc) return rc;
  DINOX_REQUIRE(flags && (reinterpret_cast<uintptr_t>(flags) & 3u) == 0, DINOX_E_BADARG, "gemm_bf16_balanced: flags missing");
  Operand a{A, M, lda, a_mn_major ? 1 : 0}, b{B, N, ldb, b_mn_major ? 1 : 0};
  StoreArgs sa{C, ldc, 0, accumulate, alpha, alpha_dev, bias_n, 0};
  sa.order_flags = reinterpret_cast<uint32_t*>(flags);
  sa.order_flag_count = (int64_t)(dinox_gemm_bf16_balanced_workspace_bytes(M, N) / sizeof(uint32_t));
  return launch_store(M, N, a, b, K, m_fastest, sa, stream, 1, 1);
}

int dinox_gemm_bf16_splitk(const void* A, const void* B, float* C_partials, int64_t M, int64_t N, int64_t K,
                           int64_t lda, int64_t ldb, int64_t ldc, int64_t split_stride, int splits, int a_mn_major,
                           int b_mn_major, float alpha, const float* alpha_dev, int m_fastest,
                           dinox_stream_t stream) {
  int rc = gemm_common_checks(A, B, C_partials, N, ldc, DINOX_F32, "gemm_bf16_splitk");
  if (rc) return rc;
  DINOX_REQUIRE(splits >= 1 && splits <= 64, DINOX_E_BADARG, "gemm_bf16_splitk: splits must be in [1, 64]");
  DINOX_REQUIRE(split_stride >= M * ldc && (split_stride * 4) % 16 == 0, DINOX_E_BADARG,
                "gemm_bf16_splitk: split_stride smaller than one (M, ldc) slab or misaligned");
  Operand a{A, M, lda, a_mn_major ? 1 : 0}, b{B, N, ldb, b_mn_major ? 1 : 0};
  StoreArgs sa{C_partials, ldc, 0, 0, alpha, alpha_dev, nullptr, split_stride};
  return launch_store(M, N, a, b, K, m_fastest, sa, stream, 1, splits);
}

/* C = alpha * A @ B^T with the ROWS of C sharded over `owners` buffers: tile rows [o*rows_per_owner, ...) are
 * reduce-added (cp.reduce.async.bulk.tensor .add.f32, performed in the owner's L2) into shard_ptrs[o], which may be
 * peer-mapped memory of another GPU (NVLink).  Every data-parallel rank launching this with the same shard table
 * performs the reduce-scatter of dW2 inside the GEMM epilogue, tile by tile, instead of a collective afterwards. */
int dinox_gemm_bf16_reduce_scatter(const void* A, const void* B, float* const* shard_ptrs, int owners,
                                   int64_t rows_per_owner, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                                   int64_t ldc, int a_mn_major, int b_mn_major, float alpha, const float* alpha_dev,
                                   const float* add_local, int64_t ld_local, float add_scale, dinox_stream_t stream) {
  DINOX_REQUIRE(A && B && shard_ptrs && owners >= 1 && owners <= kMaxOwners, DINOX_E_BADARG,
                "gemm_bf16_reduce_scatter: 1..%d shards", kMaxOwners);
  DINOX_REQUIRE(rows_per_owner > 0 && rows_per_owner % BM == 0 && rows_per_owner * owners >= M, DINOX_E_BADARG,
                "gemm_bf16_reduce_scatter: rows_per_owner must be a multiple of %d covering M", BM);
  DINOX_REQUIRE(ldc >= N && (ldc * 4) % 16 == 0, DINOX_E_ALIGN, "gemm_bf16_reduce_scatter: ldc misaligned");
  DINOX_REQUIRE(M > BM, DINOX_E_UNSUPPORTED, "gemm_bf16_reduce_scatter: needs more than one M tile");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a{A, M, lda, a_mn_major ? 1 : 0}, b{B, N, ldb, b_mn_major ? 1 : 0};
  DINOX_REQUIRE(!add_local || (ld_local >= N && ld_local % 4 == 0 && aligned16(add_local)), DINOX_E_ALIGN,
                "gemm_bf16_reduce_scatter: add_local / ld_local misaligned");
  EpiStore::Params ep{0, 1, alpha, alpha_dev, nullptr};
  ep.add_src = add_local; ep.ld_add = ld_local; ep.add_scale = add_scale;
  OutDesc od;
  od.is_bf16 = 0; od.ld = ldc; od.owners = owners; od.rows_per_owner = rows_per_owner;
  for (int i = 0; i < owners; ++i) {
    DINOX_REQUIRE(shard_ptrs[i] && aligned16(shard_ptrs[i]), DINOX_E_ALIGN, "gemm_bf16_reduce_scatter: shard %d null or misaligned", i);
    od.owner_ptr[i] = shard_ptrs[i];
  }
  if (N % 384 == 0)
    return launch<384, 3, 1, 2, EpiStore, kResNone, kMaxOwners>(a, b, nullptr, nullptr, M, N, K, 0, ep, od, stream, "gemm_bf16_rs<384,pair>");
  if (N % 256 == 0)
    return launch<256, 1, 1, 2, EpiStore, kResNone, kMaxOwners>(a, b, nullptr, nullptr, M, N, K, 0, ep, od, stream, "gemm_bf16_rs<256,pair>");
  set_error("gemm_bf16_reduce_scatter: N must be a multiple of 256 or 384 (got %lld)", (long long)N);
  return DINOX_E_UNSUPPORTED;
}

int dinox_gemm_splitk_plan(int64_t M, int64_t N, int64_t K) {
  // number of K splits that fills whole waves of the persistent grid (one CTA per SM) for a GEMM
  // with few output tiles and a long reduction; 1 when the tile count already fills the machine
  if (M <= 0 || N <= 0 || K <= 0) return 1;
  const int64_t bn = (N % 384 == 0) ? 384 : ((N % 256 == 0 || N > 2048) ? 256 : 128);
  // CTA pairs walk super tiles of two M tiles, one pair per two SMs
  const bool pair = M > BM && pair_enabled(kPairStore);
  const int64_t tiles = (((M + 127) / 128 + (pair ? 1 : 0)) / (pair ? 2 : 1)) * ((N + bn - 1) / bn);
  const int64_t kb = (K + 63) / 64;
  const int sms = num_sms() / (pair ? 2 : 1);
  if (tiles >= 3 * (int64_t)sms || kb < 16) return 1;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 32 && s * 8 <= kb; ++s) {
    const int64_t kbs = (kb + s - 1) / s;
    if (kbs * (s - 1) >= kb) continue;                  // an empty last split
    const int64_t waves = (tiles * s + sms - 1) / sms;
    // time ~ waves * (k-blocks per split + fixed per-tile epilogue cost in k-block units)
    const double cost = (double)waves * ((double)kbs + 6.0);
    if (cost < best_cost * 0.97) { best_cost = cost; best = s; }
  }
  return best;
}

/* diagnostics: how many CTA-pair clusters of the pass-1 kernel can be co-resident (74 = every SM) */
int dinox_debug_max_active_clusters(int cluster_size) {
  auto kern = cluster_size == 2 ? gemm_kernel<256, 1, 1, 2, EpiStats> : gemm_kernel<256, 1, 1, 1, EpiStats>;
  constexpr int smem = smem_bytes<256, 2, EpiStats>() > smem_bytes<256, 1, EpiStats>() ? smem_bytes<256, 2, EpiStats>()
                                                                                      : smem_bytes<256, 1, EpiStats>();
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3((2 + EpiStats::kEpiWarps) * 32);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -2; }
  return n;
}

size_t dinox_head_stats_workspace_bytes(int64_t rows, int64_t K) {
  if (rows <= 0 || K <= 0) return 0;
  const int64_t n_tiles = (K + 255) / 256;
  return (size_t)rows * EpiStats::kGroups * (n_tiles + 1) * sizeof(float2);   // + 1: run-mode slot bound
}

int dinox_head_stats(const void* H, const void* W2, int64_t rows, int64_t K, int64_t D, int64_t ldh, int64_t ldw,
                     float inv_tau, const float* col2, float* lse_nat, float* lse2, void* workspace,
                     dinox_stream_t stream) {
  DINOX_REQUIRE(H && W2 && workspace && (lse_nat || lse2), DINOX_E_BADARG, "head_stats: null pointer");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a{H, rows, ldh, 0}, b{W2, K, ldw, 0};
  EpiStats::Params ep{inv_tau * DINOX_LOG2E, col2, reinterpret_cast<float2*>(workspace), 1, 1, 1};
  const bool cl2 = rows > BM && pair_enabled(kPairStats);
  if (resa_enabled(kPairStats, D)) {
    // prototype tiles fastest in contiguous runs: the H rows of one M tile stay in shared memory and the row
    // statistics stay in registers over a run (one partial per row and cluster)
    const int cl = cl2 ? 2 : 1;
    const int64_t num_n = (K + 255) / 256, num_m_super = ((rows + BM - 1) / BM + cl - 1) / cl, super = num_m_super * num_n;
    const int ncl = launch_grid(super, cl) / cl;
    const int per = (int)((super + ncl - 1) / ncl);
    const int run_slots = (int)((num_n + per - 1) / per) + 1;
    // run_slots <= num_n + 1: inside the workspace dinox_head_stats_workspace_bytes() asks for
    EpiStatsRun::Params er{ep.scale2, ep.col2, ep.partial, cl, per, run_slots};
    rc = cl2 ? launch<256, 1, 1, 2, EpiStatsRun, kResA>(a, b, nullptr, nullptr, rows, K, D, /*m_fastest=*/0, er, OutDesc{}, stream, "head_stats<pair,resA>")
             : launch<256, 1, 1, 1, EpiStatsRun, kResA>(a, b, nullptr, nullptr, rows, K, D, /*m_fastest=*/0, er, OutDesc{}, stream, "head_stats<resA>");
    if (rc) return rc;
    stats_merge_run_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(er.partial, rows, EpiStats::kGroups, run_slots, (int)num_n,
                                                                               per, cl, lse_nat, lse2);
    return check_launch("stats_merge_run_kernel", stream);
  }
  rc = cl2 ? launch<256, 1, 1, 2, EpiStats>(a, b, nullptr, nullptr, rows, K, D, /*m_fastest=*/1, ep, OutDesc{}, stream, "head_stats<pair>")
           : launch<256, 1, 1, 1, EpiStats>(a, b, nullptr, nullptr, rows, K, D, /*m_fastest=*/1, ep, OutDesc{}, stream, "head_stats");
  if (rc) return rc;
  const int n_part = EpiStats::kGroups * (int)((K + 255) / 256);
  stats_merge_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const float2*>(workspace), rows, n_part,
                                                                     lse_nat, lse2);
  return check_launch("stats_merge_kernel", stream);
}

size_t dinox_head_grad_workspace_bytes(int64_t K, int64_t E) {
  if (K <= 0 || E <= 0) return 0;
  return (size_t)(1024 * EpiGradT::kEpiWarps * 2) * sizeof(float) + 256;  // per-CTA partials, grid <= 1024
}

/* rows of the db2_partial buffer of dinox_head_grad: one per (128-entry tile, column group of the epilogue) */
int64_t dinox_head_grad_db2_rows(int64_t E) { return E <= 0 ? 0 : EpiGradT::kGroups * ((E + 127) / 128); }

int dinox_head_grad(const void* W2s, const void* W2t, const void* HsE, const void* HtE, int64_t K, int64_t D,
                    int64_t E, int64_t ldw_s, int64_t ldw_t, int64_t ldh_s, int64_t ldh_t, float inv_tau_s,
                    float inv_tau_t, const float* cs2, const float* ct2, const float* ct2_alt, int64_t alt_from,
                    const float* lse2_e, const float* rb2_e, const float* cw_e, void* Gt, int64_t ldg,
                    float* db2_partial, float* loss_out, int loss_accumulate, void* workspace,
                    dinox_stream_t stream) {
  DINOX_REQUIRE(W2s && W2t && HsE && HtE && cs2 && ct2 && lse2_e && rb2_e && cw_e && Gt && loss_out && workspace,
                DINOX_E_BADARG, "head_grad: null pointer");
  DINOX_REQUIRE(ldg >= E && (ldg * 2) % 16 == 0 && aligned16(Gt), DINOX_E_ALIGN, "head_grad: Gt / ldg misaligned");
  DINOX_REQUIRE(alt_from % 128 == 0, DINOX_E_BADARG, "head_grad: alt_from must be a multiple of 128");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a0{W2s, K, ldw_s, 0}, b0{HsE, E, ldh_s, 0}, a1{W2t, K, ldw_t, 0}, b1{HtE, E, ldh_t, 0};
  EpiGradT::Params ep;
  ep.as2 = inv_tau_s * DINOX_LOG2E; ep.at2 = inv_tau_t * DINOX_LOG2E; ep.inv_tau_s = inv_tau_s;
  ep.cs2 = cs2; ep.ct2 = ct2; ep.ct2_alt = ct2_alt; ep.alt_from = (int)alt_from;
  ep.lse2 = lse2_e; ep.rb2 = rb2_e; ep.cw = cw_e;
  ep.db2_partial = db2_partial; ep.loss_partial = reinterpret_cast<float*>(workspace);
  OutDesc od;
  od.ptr = Gt; od.is_bf16 = 1; od.ld = ldg; od.slab_stride = K * ldg; od.slabs = 1;
  od.row_bytes = EpiGradT::kRowBytes;
  // entry tiles fastest: the 2 x (K-tile of W2) operands stay put while HsE/HtE (L2-resident) stream
  const bool cl2 = K > BM && pair_enabled(kPairGrad);
  if (cl2 && resa_enabled(kPairGrad, D))   // entries of an N tile resident (student + teacher), W2 tiles stream
    rc = launch<128, 1, 2, 2, EpiGradT, kResB>(a0, b0, &a1, &b1, K, E, D, /*m_fastest=*/1, ep, od, stream, "head_grad<pair,resB>");
  else
    rc = cl2 ? launch<128, 1, 2, 2, EpiGradT>(a0, b0, &a1, &b1, K, E, D, /*m_fastest=*/0, ep, od, stream, "head_grad<pair>")
             : launch<128, 1, 2, 1, EpiGradT>(a0, b0, &a1, &b1, K, E, D, /*m_fastest=*/0, ep, od, stream, "head_grad");
  if (rc) return rc;
  const int cl = cl2 ? 2 : 1;
  const int grid = launch_grid((((K + 127) / 128 + cl - 1) / cl) * ((E + 127) / 128), cl);
  pair_sum_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float*>(workspace), (int64_t)grid * EpiGradT::kEpiWarps, loss_out, loss_accumulate);
  return check_launch("pair_sum_kernel", stream);
}

/* ---- teacher, one pass: statistics + un-normalised fp16 probabilities (see EpiTeachQT) ---- */
size_t dinox_head_teacher_workspace_bytes(int64_t rows, int64_t K) {
  if (rows <= 0 || K <= 0) return 0;
  return (size_t)rows * EpiTeachQ<false>::kGroups * ((K + kRbTile - 1) / kRbTile + 1) * sizeof(float2);   // + 1: run-mode slot bound
}
/* granules (rows of `refs`) per 256-prototype tile: 128 / 64 prototypes per granule with 8 / 16 epilogue warps */
int dinox_head_teacher_granules_per_tile(void) { return EpiTeachQ<false>::kGroups; }
/* prototypes per tile of the read-back pair: rows of qt are padded to a multiple of it */
int dinox_head_teacher_tile_cols(void) { return kRbTile; }

int dinox_head_teacher(const void* H, const void* W2, int64_t rows, int64_t K, int64_t D, int64_t ldh, int64_t ldw,
                       float inv_tau, const float* col2, const float* col2_alt, int64_t alt_from_row, void* qt,
                       int64_t ldq, float* refs, int64_t ld_refs, float* lse_nat, float* lse2, void* workspace,
                       dinox_stream_t stream) {
  DINOX_REQUIRE(H && W2 && qt && refs && workspace, DINOX_E_BADARG, "head_teacher: null pointer");
  DINOX_REQUIRE(rows > 0 && K > 0 && D > 0, DINOX_E_BADARG, "head_teacher: empty problem");
  const int64_t num_n = (K + kRbTile - 1) / kRbTile;
  DINOX_REQUIRE(ldq >= num_n * kRbTile && aligned16(qt) && (ldq * 2) % 16 == 0, DINOX_E_ALIGN,
                "head_teacher: qt rows must be padded to whole prototype tiles (ldq >= %lld)", (long long)(num_n * kRbTile));
  DINOX_REQUIRE(ld_refs >= rows, DINOX_E_BADARG, "head_teacher: ld_refs < rows");
  DINOX_REQUIRE(!col2_alt || alt_from_row % BM == 0, DINOX_E_BADARG, "head_teacher: alt_from_row must be a multiple of 128");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a{H, rows, ldh, 0}, b{W2, K, ldw, 0};
  OutDesc od;
  od.ptr = qt; od.is_bf16 = 1; od.ld = ldq; od.slab_stride = rows * ldq; od.slabs = 1; od.cols = num_n * kRbTile;
  const int alt_mtile = col2_alt ? (int)(alt_from_row / BM) : (1 << 30);
  const bool cl2 = rows > BM && pair_enabled(kPairStats);
  if (cl2 && readback_cols(1, D)) {
    // W2 tile resident (a CTA of a pair holds half of it: 96 KB), every cluster sweeps the M tiles of its prototype tile in step with the others: W2 is read
    // from HBM exactly once per launch (the probability stream would evict it from L2 between two M-tile runs)
    EpiTeachQ<false>::Params ep{inv_tau * DINOX_LOG2E, col2, col2_alt, alt_mtile, reinterpret_cast<float2*>(workspace), 1, 1, 1, refs, ld_refs};
    rc = launch<kRbTile, 1, 1, 2, EpiTeachQ<false>, kResB>(a, b, nullptr, nullptr, rows, K, D, 1, ep, od, stream, "head_teacher<pair,resB>");
    if (rc) return rc;
    if (lse_nat || lse2) {
      stats_merge_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const float2*>(workspace), rows,
                                                                         EpiTeachQ<false>::kGroups * (int)num_n, lse_nat, lse2);
      return check_launch("stats_merge_kernel", stream);
    }
    return DINOX_OK;
  }
  if (cl2 && resa_enabled(kPairStats, D)) {   // (a lone CTA has no room for a resident operand beside the staging buffers)
    const int cl = 2;
    const int64_t num_m_super = ((rows + BM - 1) / BM + cl - 1) / cl, super = num_m_super * num_n;
    const int ncl = launch_grid(super, cl) / cl;
    const int per = (int)((super + ncl - 1) / ncl);
    const int run_slots = (int)((num_n + per - 1) / per) + 1;
    EpiTeachQ<true>::Params er{inv_tau * DINOX_LOG2E, col2, col2_alt, alt_mtile, reinterpret_cast<float2*>(workspace), cl, per,
                                run_slots, refs, ld_refs};
    rc = launch<kRbTile, 1, 1, 2, EpiTeachQ<true>, kResA>(a, b, nullptr, nullptr, rows, K, D, 0, er, od, stream, "head_teacher<pair,resA>");
    if (rc) return rc;
    if (lse_nat || lse2) {
      stats_merge_run_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(er.partial, rows, EpiTeachQ<true>::kGroups, run_slots,
                                                                                 (int)num_n, per, cl, lse_nat, lse2);
      return check_launch("stats_merge_run_kernel", stream);
    }
    return DINOX_OK;
  }
  EpiTeachQ<false>::Params ep{inv_tau * DINOX_LOG2E, col2, col2_alt, alt_mtile, reinterpret_cast<float2*>(workspace), 1, 1, 1, refs, ld_refs};
  rc = cl2 ? launch<kRbTile, 1, 1, 2, EpiTeachQ<false>>(a, b, nullptr, nullptr, rows, K, D, 1, ep, od, stream, "head_teacher<pair>")
           : launch<kRbTile, 1, 1, 1, EpiTeachQ<false>>(a, b, nullptr, nullptr, rows, K, D, 1, ep, od, stream, "head_teacher");
  if (rc) return rc;
  if (lse_nat || lse2) {
    stats_merge_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const float2*>(workspace), rows,
                                                                       EpiTeachQ<false>::kGroups * (int)num_n, lse_nat, lse2);
    return check_launch("stats_merge_kernel", stream);
  }
  return DINOX_OK;
}

/* ---- pass 2, student logits only, teacher probabilities read back (see EpiGradR) ---- */
size_t dinox_head_grad2_workspace_bytes(int64_t E, int64_t K) {
  if (K <= 0 || E <= 0) return 0;
  return (size_t)(1024 * EpiGradR::kEpiWarps * 2) * sizeof(float) + 256;   // per-CTA loss partials, grid <= 1024
}
int64_t dinox_head_grad2_db2_rows(int64_t E) { return E <= 0 ? 0 : (E + 31) / 32; }

int dinox_head_grad2(const void* HsE, const void* W2s, int64_t E, int64_t K, int64_t D, int64_t ldh, int64_t ldw,
                     float inv_tau_s, const float* cs2, const float* lse2_e, const float* cw_e, const float* rb2_e,
                     const int32_t* trow_e, const int32_t* srow_e, const void* qt, int64_t ldq, const float* refs,
                     int64_t ld_refs, int64_t alt_from, void* G, int64_t ldg, float* db2_partial, float* loss_out,
                     int loss_accumulate, void* workspace, uint32_t* ticket, dinox_stream_t stream) {
  DINOX_REQUIRE(HsE && W2s && cs2 && lse2_e && cw_e && rb2_e && trow_e && qt && refs && G && loss_out && workspace,
                DINOX_E_BADARG, "head_grad2: null pointer");
  DINOX_REQUIRE(E > 0 && K > 0 && D > 0, DINOX_E_BADARG, "head_grad2: empty problem");
  const int64_t num_n = (K + kRbTile - 1) / kRbTile;
  DINOX_REQUIRE(ldq >= num_n * kRbTile && (ldq * 2) % 32 == 0 && (reinterpret_cast<uintptr_t>(qt) & 31u) == 0, DINOX_E_ALIGN,
                "head_grad2: qt must be 32-byte aligned with rows padded to whole prototype tiles");
  DINOX_REQUIRE(ldg >= K && (ldg * 2) % 16 == 0 && aligned16(G), DINOX_E_ALIGN, "head_grad2: G / ldg misaligned");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a{HsE, E, ldh, 0}, b{W2s, K, ldw, 0};
  EpiGradR::Params ep;
  ep.as2 = inv_tau_s * DINOX_LOG2E; ep.inv_tau_s = inv_tau_s;
  ep.cs2 = cs2; ep.lse2 = lse2_e; ep.cw = cw_e; ep.rb2 = rb2_e; ep.trow = trow_e; ep.srow = srow_e;
  ep.ticket = ticket; ep.loss_out = loss_out; ep.loss_accumulate = loss_accumulate;
  ep.qt = reinterpret_cast<const __half*>(qt); ep.ldq = ldq; ep.refs = refs; ep.ld_refs = ld_refs;
  ep.alt_from = (int)(alt_from < 0 ? 0 : (alt_from > (1ll << 30) ? (1ll << 30) : alt_from));
  ep.db2_partial = db2_partial; ep.loss_partial = reinterpret_cast<float*>(workspace);
  OutDesc od;
  od.ptr = G; od.is_bf16 = 1; od.ld = ldg; od.slab_stride = E * ldg; od.slabs = 1;
  const bool cl2 = E > BM && pair_enabled(kPairGrad);
  const int cl = cl2 ? 2 : 1;
  ep.run_rows = (cl2 && !readback_cols(2, D) && resa_enabled(kPairStats, D)) ? 1 : 0;
  ep.m_step = (cl2 && readback_cols(2, D)) ? 2 : 0;
  if (cl2 && readback_cols(2, D))   // W2 tile resident, clusters sweep the entry tiles in step (W2 read from HBM once)
    rc = launch<kRbTile, 1, 1, 2, EpiGradR, kResB>(a, b, nullptr, nullptr, E, K, D, 1, ep, od, stream, "head_grad2<pair,resB>");
  else if (cl2 && resa_enabled(kPairStats, D))   // the entries of an M tile resident, prototype tiles walked in one contiguous run
    rc = launch<kRbTile, 1, 1, 2, EpiGradR, kResA>(a, b, nullptr, nullptr, E, K, D, 0, ep, od, stream, "head_grad2<pair,resA>");
  else
    rc = cl2 ? launch<kRbTile, 1, 1, 2, EpiGradR>(a, b, nullptr, nullptr, E, K, D, 1, ep, od, stream, "head_grad2<pair>")
             : launch<kRbTile, 1, 1, 1, EpiGradR>(a, b, nullptr, nullptr, E, K, D, 1, ep, od, stream, "head_grad2");
  if (rc || ticket) return rc;
  const int grid = launch_grid((((E + BM - 1) / BM + cl - 1) / cl) * num_n, cl);
  pair_sum_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float*>(workspace), (int64_t)grid * EpiGradR::kEpiWarps, loss_out,
                                          loss_accumulate, /*write_total=*/1);
  return check_launch("pair_sum_kernel", stream);
}

int dinox_gemm_bf16_batched(const void* A, const void* B, void* C, int64_t batches, int64_t M, int64_t N, int64_t K,
                            int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c,
                            int a_mn_major, int b_mn_major, int out_dtype, int accumulate, float alpha,
                            const float* alpha_dev, dinox_stream_t stream) {
  DINOX_REQUIRE(batches >= 1, DINOX_E_BADARG, "gemm_bf16_batched: bad arguments");
  int rc = gemm_common_checks(A, B, C, N, ldc, out_dtype, "gemm_bf16_batched");
  if (rc) return rc;
  const int es = out_dtype == DINOX_F32 ? 4 : 2;
  DINOX_REQUIRE((stride_c * es) % 16 == 0, DINOX_E_ALIGN, "gemm_bf16_batched: stride_c misaligned");
  Operand a{A, M, lda, a_mn_major ? 1 : 0, stride_a}, b{B, N, ldb, b_mn_major ? 1 : 0, stride_b};
  StoreArgs sa{C, ldc, out_dtype == DINOX_BF16, accumulate, alpha, alpha_dev, nullptr, stride_c};
  return launch_store(M, N, a, b, K, 1, sa, stream, batches, 1);
}

size_t dinox_gram_diff_workspace_bytes(int64_t batches, int64_t tokens) {
  if (batches <= 0 || tokens <= 0) return 0;
  return (size_t)(1024 * 8) * sizeof(float) + 256;  // per-CTA partials, grid <= 1024
}

int dinox_gram_diff(const void* xn_s, const void* xn_t, int64_t batches, int64_t tokens, int64_t D, void* delta,
                    int64_t ldd, float loss_scale, float* loss_out, void* workspace, dinox_stream_t stream) {
  DINOX_REQUIRE(xn_s && xn_t && loss_out && workspace && batches >= 1 && tokens > 0 && D > 0, DINOX_E_BADARG,
                "gram_diff: bad arguments");
  DINOX_REQUIRE(!delta || (aligned16(delta) && ldd % 8 == 0 && ldd >= tokens), DINOX_E_ALIGN, "gram_diff: delta / ldd misaligned");
  int rc = require_sm100();
  if (rc) return rc;
  // a single image still goes through the 3-D path (batches = 1 uses 2-D maps over (tokens, D))
  Operand a0{xn_s, tokens, D, 0, tokens * D}, a1{xn_t, tokens, D, 0, tokens * D};
  EpiGramDiff::Params ep{reinterpret_cast<__nv_bfloat16*>(delta), ldd, tokens * ldd, reinterpret_cast<float*>(workspace)};
  rc = launch<128, 1, 2, 1, EpiGramDiff>(a0, a0, &a1, &a1, tokens, tokens, D, 1, ep, OutDesc{}, stream, "gram_diff", batches);
  if (rc) return rc;
  const int64_t mt = (tokens + 127) / 128;
  sum_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float*>(workspace), (int64_t)launch_grid(batches * mt * mt, 1) * 8,
                                     loss_scale, loss_out, 0);
  return check_launch("sum_kernel", stream);
}

}  // extern "C"

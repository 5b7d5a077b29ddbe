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

namespace dinox {
namespace gemm {

// =============================================================================================
// Epilogue 1: store   C = alpha * acc (+ bias[n])  ->  fp32 | bf16, optional C += ...
// =============================================================================================
struct EpiStore {
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
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int, int, uint8_t*) {}
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, int, uint32_t tmem_acc,
                                                int, int, int lane, uint8_t*, State&) {
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
// =============================================================================================
struct EpiStats {
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiSmemBytes = 0;
  struct Params {
    float scale2;
    const float* col2;   // (N) log2-unit column offsets, may be NULL
    float2* partial;     // (M, 2*num_n_tiles)
  };
  struct State {};
  static __device__ __forceinline__ void finish(const Params&, const CoreParams&, int, int, State&) {}
  template <int BN>
  struct Impl {
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int, int, uint8_t*) {}
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, int, uint32_t tmem_acc,
                                                int, int epi_warp, int lane, uint8_t*, State&) {
      const int q = epi_quarter();
      const int half = epi_warp >> 2;
      const int row = tc.m_tile * BM + q * 32 + lane;
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + half * (BN / 2);
      float m = -INFINITY, s = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        float v[32];
        sm100::tmem_ld32(taddr + c * 32, v);
        const int col0 = tc.n_tile * BN + half * (BN / 2) + c * 32;
        if (col0 >= p.N) break;
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool ok = (col0 + j < p.N);
          const float cb = (e.col2 && ok) ? __ldg(e.col2 + col0 + j) : 0.f;
          v[j] = ok ? fmaf(v[j], e.scale2, cb) : -INFINITY;
          cm = fmaxf(cm, v[j]);
        }
        const float mn = fmaxf(m, cm);
        float acc = s * fast_ex2(m - mn);   // m = -inf -> 0
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += fast_ex2(v[j] - mn);
        s = acc;
        m = mn;
      }
      if (row < p.M) e.partial[(int64_t)row * (2 * p.num_n_tiles) + tc.n_tile * 2 + half] = make_float2(m, s);
    }
  };
};

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

// =============================================================================================
// Epilogue 3 (pass 2, "transposed"): TMEM lanes = prototypes k, columns = entries e.
// Sub-GEMM 0 = student logits  S[k,e] = W2s[k,:] . Hs[e,:],  sub-GEMM 1 = teacher T[k,e].
//   p = 2^(S*as2 + cs2[k] - lse2[e])       student softmax prob
//   q = 2^(T*at2 + ct2[k] - rb2[e])        teacher prob (centre or Sinkhorn biases)
//   G[k,e]  = cw[e] * inv_tau_s * (p - q)                 -> bf16, Gt (K, ldg)
//   loss   += cw[e] * q * ln2 * (lse2[e] - s2)            (-q ln p)
//   db2[k] += G[k,e]                                      (fp32, before rounding)
// =============================================================================================
struct EpiGradT {
  static constexpr int kEpiWarps = 8;
  static constexpr int kEpiSmemBytes = 2 * 3 * 128 * 4;
  struct Params {
    float as2, at2, inv_tau_s;
    const float* cs2;       // (K)
    const float* ct2;       // (K) for entries < alt_from
    const float* ct2_alt;   // (K) for entries >= alt_from (iBOT patch centre); may be NULL
    int alt_from;           // multiple of BN
    const float* lse2;      // (E) student LSE (log2) per entry
    const float* rb2;       // (E) teacher row bias (log2) per entry
    const float* cw;        // (E) entry weight (norm * group weight), 0 for padding
    __nv_bfloat16* gt;      // (K, ldg)
    int64_t ldg;
    float* db2_partial;     // (2*num_n_tiles, M) or NULL
    float* loss_partial;    // (gridDim.x * 8 * 2): per CTA, per epilogue warp, {entries < alt_from, >= alt_from}
  };
  struct State {
    float loss_a = 0.f, loss_b = 0.f;
  };
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams&, int epi_warp, int lane, State& st) {
    const float a = warp_sum(st.loss_a), b = warp_sum(st.loss_b);
    if (lane == 0) {
      e.loss_partial[((int64_t)blockIdx.x * 8 + epi_warp) * 2 + 0] = a * DINOX_LN2;
      e.loss_partial[((int64_t)blockIdx.x * 8 + epi_warp) * 2 + 1] = b * DINOX_LN2;
    }
  }
  template <int BN>
  struct Impl {
    static __device__ __forceinline__ void prologue(const Params& e, const CoreParams& p, TileCoord tc, int acc_stage,
                                                    int epi_warp, int lane, uint8_t* smem) {
      float* buf = reinterpret_cast<float*>(smem) + acc_stage * 3 * 128;
      const int i = epi_warp * 32 + lane;  // 0..255
      if (i < BN) {
        const int ent = tc.n_tile * BN + i;
        const bool ok = ent < p.N;
        buf[i] = ok ? e.lse2[ent] : 0.f;
        buf[128 + i] = ok ? e.rb2[ent] : 0.f;
        buf[256 + i] = ok ? e.cw[ent] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, int t, uint32_t tmem_acc,
                                                int acc_stage, int epi_warp, int lane, uint8_t* smem, State& st) {
      static_assert(BN == 128, "EpiGradT is written for 128-entry tiles");
      const float* buf = reinterpret_cast<const float*>(smem) + acc_stage * 3 * 128;
      const int q = epi_quarter();
      const int half = epi_warp >> 2;
      const int k = tc.m_tile * BM + q * 32 + lane;   // prototype
      const bool kok = k < p.M;
      const float* ctp = (e.ct2_alt && tc.n_tile * BN >= e.alt_from) ? e.ct2_alt : e.ct2;
      const float cs = kok ? __ldg(e.cs2 + k) : 0.f;
      const float ct = kok ? __ldg(ctp + k) : 0.f;
      const uint32_t ts = tmem_acc + ((uint32_t)(q * 32) << 16) + half * 64;
      const uint32_t tt = ts + BN;
      float loss = 0.f, db2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float sv[32], tv[32];
        sm100::tmem_ld32x2(ts + c * 32, tt + c * 32, sv, tv);
        const int e0 = half * 64 + c * 32;
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float lse2 = buf[e0 + j], rb2 = buf[128 + e0 + j], cw = buf[256 + e0 + j];
          const float s2 = fmaf(sv[j], e.as2, cs);
          const float pp = fast_ex2(s2 - lse2);
          const float qq = fast_ex2(fmaf(tv[j], e.at2, ct) - rb2);
          const float g = cw * e.inv_tau_s * (pp - qq);
          loss = fmaf(cw * qq, lse2 - s2, loss);
          db2 += g;
          sv[j] = g;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(sv[2 * j], sv[2 * j + 1]);
          packed[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        const int ent0 = tc.n_tile * BN + e0;
        if (kok) {
          __nv_bfloat16* o = e.gt + (int64_t)k * e.ldg + ent0;
          if (ent0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(o + j * 8) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (ent0 + j < p.N) o[j] = __float2bfloat16_rn(sv[j]);
          }
        }
      }
      if (!kok) { loss = 0.f; db2 = 0.f; }
      if (e.db2_partial && kok) e.db2_partial[(int64_t)(tc.n_tile * 2 + half) * p.M + k] = db2;
      if (e.ct2_alt && tc.n_tile * BN >= e.alt_from) st.loss_b += loss; else st.loss_a += loss;
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
    static __device__ __forceinline__ void prologue(const Params&, const CoreParams&, TileCoord, int, int, int, uint8_t*) {}
    static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, int t, uint32_t tmem_acc,
                                                int, int epi_warp, int lane, uint8_t*, State& st) {
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
                                                         float* __restrict__ out, int accumulate) {
  __shared__ float red[64];
  float a = 0.f, b = 0.f;
  for (int64_t i = threadIdx.x; i < n_pairs; i += 1024) { a += x[2 * i]; b += x[2 * i + 1]; }
  a = block_sum<1024>(a, red);
  b = block_sum<1024>(b, red);
  if (threadIdx.x == 0) {
    out[0] = a + (accumulate ? out[0] : 0.f);
    out[1] = b + (accumulate ? out[1] : 0.f);
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
  static constexpr int kEpiWarps = Epi::kEpiWarps;
  using Params = typename Epi::Params;
  using State = typename Epi::State;
  static __device__ __forceinline__ void prologue(const Params& e, const CoreParams& p, TileCoord tc, int a, int w, int l, uint8_t* s) {
    Epi::template Impl<BN>::prologue(e, p, tc, a, w, l, s);
  }
  static __device__ __forceinline__ void tile(const Params& e, const CoreParams& p, TileCoord tc, int t, uint32_t tm, int a, int w, int l, uint8_t* s, State& st) {
    Epi::template Impl<BN>::tile(e, p, tc, t, tm, a, w, l, s, st);
  }
  static __device__ __forceinline__ void finish(const Params& e, const CoreParams& p, int w, int l, State& st) {
    Epi::finish(e, p, w, l, st);
  }
};

template <int BN, int NSUB, class Epi>
__global__ void __launch_bounds__((2 + Epi::kEpiWarps) * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
            const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
            const CoreParams p, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  gemm_body<BN, NSUB, EpiAdapter<BN, Epi>>(p, ep, &tmA0, &tmB0, &tmA1, &tmB1, smem_raw);
}

static inline int launch_grid(int64_t tiles) {
  int grid = num_sms();
  return grid > tiles ? (int)tiles : grid;
}

struct Operand {
  const void* ptr;
  int64_t rows;      // M (or N) extent
  int64_t ld;        // leading dimension in elements of the stored matrix
  int mn_major;      // 0: stored (rows, K) ; 1: stored (K, rows)
  int64_t batch_stride = 0;  // elements between consecutive problems (batched launches)
};

static int make_operand_tmap(CUtensorMap* tm, const Operand& o, int64_t K, int tile_rows, int64_t batches, const char* what) {
  if (batches > 1) {
    if (!o.mn_major) return make_tmap_bf16_3d(tm, o.ptr, batches, o.rows, K, o.ld, o.batch_stride, tile_rows, what);
    return make_tmap_bf16_3d(tm, o.ptr, batches, K, o.rows, o.ld, o.batch_stride, BK, what);
  }
  if (!o.mn_major) return make_tmap_bf16_2d(tm, o.ptr, o.rows, K, o.ld, tile_rows, what);
  return make_tmap_bf16_2d(tm, o.ptr, K, o.rows, o.ld, BK, what);  // box = 64 k-rows x 64 mn-elements
}

template <int BN, int NSUB, class Epi>
static int launch(const Operand& a0, const Operand& b0, const Operand* a1, const Operand* b1, int64_t M, int64_t N,
                  int64_t K, int m_fastest, const typename Epi::Params& ep, cudaStream_t stream, const char* name,
                  int64_t batches = 1) {
  DINOX_REQUIRE(M > 0 && N > 0 && K > 0, DINOX_E_BADARG, "%s: empty problem", name);
  DINOX_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), DINOX_E_BADARG, "%s: dimension too large", name);
  CUtensorMap tA0, tB0, tA1, tB1;
  int rc;
  DINOX_REQUIRE(batches >= 1 && batches < (1 << 20), DINOX_E_BADARG, "%s: bad batch count", name);
  if ((rc = make_operand_tmap(&tA0, a0, K, BM, batches, "A"))) return rc;
  if ((rc = make_operand_tmap(&tB0, b0, K, BN, batches, "B"))) return rc;
  tA1 = tA0; tB1 = tB0;
  if (NSUB == 2) {
    DINOX_REQUIRE(a1 && b1 && a1->mn_major == a0.mn_major && b1->mn_major == b0.mn_major, DINOX_E_BADARG,
                  "%s: second operand pair missing or layout mismatch", name);
    if ((rc = make_operand_tmap(&tA1, *a1, K, BM, batches, "A1"))) return rc;
    if ((rc = make_operand_tmap(&tB1, *b1, K, BN, batches, "B1"))) return rc;
  }
  CoreParams p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.num_m_tiles = (int)((M + BM - 1) / BM);
  p.num_n_tiles = (int)((N + BN - 1) / BN);
  p.num_k_blocks = (int)((K + BK - 1) / BK);
  p.a_mn_major = a0.mn_major; p.b_mn_major = b0.mn_major; p.m_fastest = m_fastest;
  p.batches = (int)batches;
  auto kern = gemm_kernel<BN, NSUB, Epi>;
  constexpr int smem = smem_bytes<BN, Epi>();
  static bool attr_set = false;
  if (!attr_set) {
    DINOX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles * batches;
  DINOX_REQUIRE(tiles < (1ll << 31), DINOX_E_BADARG, "%s: too many tiles", name);
  int grid = launch_grid(tiles);
  kern<<<grid, (2 + Epi::kEpiWarps) * 32, smem, stream>>>(tA0, tB0, tA1, tB1, p, ep);
  return check_launch(name, stream);
}

}  // namespace gemm
}  // namespace dinox

extern "C" {
using namespace dinox;
using namespace dinox::gemm;

int dinox_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                    int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int out_dtype, int accumulate,
                    float alpha, const float* alpha_dev, const float* bias_n, int m_fastest,
                    dinox_stream_t stream) {
  DINOX_REQUIRE(A && B && C, DINOX_E_BADARG, "gemm_bf16: null pointer");
  DINOX_REQUIRE(out_dtype == DINOX_F32 || out_dtype == DINOX_BF16, DINOX_E_BADARG, "gemm_bf16: out dtype must be f32 or bf16");
  DINOX_REQUIRE(aligned16(C) && (ldc * (out_dtype == DINOX_F32 ? 4 : 2)) % 16 == 0, DINOX_E_ALIGN,
                "gemm_bf16: C / ldc not 16-byte aligned");
  DINOX_REQUIRE(ldc >= N, DINOX_E_BADARG, "gemm_bf16: ldc < N");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a{A, M, lda, a_mn_major ? 1 : 0}, b{B, N, ldb, b_mn_major ? 1 : 0};
  EpiStore::Params ep{C, ldc, out_dtype == DINOX_BF16, accumulate, alpha, alpha_dev, bias_n, 0};
  // widest tile that divides N without waste; ragged N falls back to 128-wide tiles
  if (N % 256 == 0 || N > 2048) return launch<256, 1, EpiStore>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, stream, "gemm_bf16<256>");
  if (N % 192 == 0) return launch<192, 1, EpiStore>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, stream, "gemm_bf16<192>");
  return launch<128, 1, EpiStore>(a, b, nullptr, nullptr, M, N, K, m_fastest, ep, stream, "gemm_bf16<128>");
}

size_t dinox_head_stats_workspace_bytes(int64_t rows, int64_t K) {
  if (rows <= 0 || K <= 0) return 0;
  const int64_t n_tiles = (K + 255) / 256;
  return (size_t)rows * 2 * n_tiles * sizeof(float2);
}

int dinox_head_stats(const void* H, const void* W2, int64_t rows, int64_t K, int64_t D, int64_t ldh, int64_t ldw,
                     float inv_tau, const float* col2, float* lse_nat, float* lse2, void* workspace,
                     dinox_stream_t stream) {
  DINOX_REQUIRE(H && W2 && workspace && (lse_nat || lse2), DINOX_E_BADARG, "head_stats: null pointer");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a{H, rows, ldh, 0}, b{W2, K, ldw, 0};
  EpiStats::Params ep{inv_tau * DINOX_LOG2E, col2, reinterpret_cast<float2*>(workspace)};
  rc = launch<256, 1, EpiStats>(a, b, nullptr, nullptr, rows, K, D, /*m_fastest=*/1, ep, stream, "head_stats");
  if (rc) return rc;
  const int n_part = 2 * (int)((K + 255) / 256);
  stats_merge_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const float2*>(workspace), rows, n_part,
                                                                     lse_nat, lse2);
  return check_launch("stats_merge_kernel", stream);
}

size_t dinox_head_grad_workspace_bytes(int64_t K, int64_t E) {
  if (K <= 0 || E <= 0) return 0;
  return (size_t)(1024 * 8 * 2) * sizeof(float) + 256;  // per-CTA partials, grid <= 1024
}

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
  ep.gt = reinterpret_cast<__nv_bfloat16*>(Gt); ep.ldg = ldg;
  ep.db2_partial = db2_partial; ep.loss_partial = reinterpret_cast<float*>(workspace);
  // entry tiles fastest: the 2 x (K-tile of W2) operands stay put while HsE/HtE (L2-resident) stream
  rc = launch<128, 2, EpiGradT>(a0, b0, &a1, &b1, K, E, D, /*m_fastest=*/0, ep, stream, "head_grad");
  if (rc) return rc;
  const int grid = launch_grid(((K + 127) / 128) * ((E + 127) / 128));
  pair_sum_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float*>(workspace), (int64_t)grid * 8, loss_out, loss_accumulate);
  return check_launch("pair_sum_kernel", stream);
}

int dinox_gemm_bf16_batched(const void* A, const void* B, void* C, int64_t batches, int64_t M, int64_t N, int64_t K,
                            int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c,
                            int a_mn_major, int b_mn_major, int out_dtype, int accumulate, float alpha,
                            const float* alpha_dev, dinox_stream_t stream) {
  DINOX_REQUIRE(A && B && C && batches >= 1, DINOX_E_BADARG, "gemm_bf16_batched: bad arguments");
  DINOX_REQUIRE(out_dtype == DINOX_F32 || out_dtype == DINOX_BF16, DINOX_E_BADARG, "gemm_bf16_batched: out dtype must be f32 or bf16");
  const int es = out_dtype == DINOX_F32 ? 4 : 2;
  DINOX_REQUIRE(aligned16(C) && (ldc * es) % 16 == 0 && (stride_c * es) % 16 == 0 && ldc >= N, DINOX_E_ALIGN,
                "gemm_bf16_batched: C / ldc / stride_c misaligned");
  int rc = require_sm100();
  if (rc) return rc;
  Operand a{A, M, lda, a_mn_major ? 1 : 0, stride_a}, b{B, N, ldb, b_mn_major ? 1 : 0, stride_b};
  EpiStore::Params ep{C, ldc, out_dtype == DINOX_BF16, accumulate, alpha, alpha_dev, nullptr, stride_c};
  if (N % 256 == 0 || N > 2048) return launch<256, 1, EpiStore>(a, b, nullptr, nullptr, M, N, K, 1, ep, stream, "gemm_bf16_batched<256>", batches);
  if (N % 192 == 0) return launch<192, 1, EpiStore>(a, b, nullptr, nullptr, M, N, K, 1, ep, stream, "gemm_bf16_batched<192>", batches);
  return launch<128, 1, EpiStore>(a, b, nullptr, nullptr, M, N, K, 1, ep, stream, "gemm_bf16_batched<128>", batches);
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
  rc = launch<128, 2, EpiGramDiff>(a0, a0, &a1, &a1, tokens, tokens, D, 1, ep, stream, "gram_diff", batches);
  if (rc) return rc;
  const int64_t mt = (tokens + 127) / 128;
  sum_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const float*>(workspace), (int64_t)launch_grid(batches * mt * mt) * 8,
                                     loss_scale, loss_out, 0);
  return check_launch("sum_kernel", stream);
}

}  // extern "C"

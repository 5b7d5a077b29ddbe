// Row / column statistics and cross-entropy on MATERIALISED logits.
// This is the drop-in path behind DINOLoss.forward(student_out, teacher_out, ...)
// (scripts/phase5_big_run.py:692-720) plus the Sinkhorn-Knopp (E2) passes.  All kernels are
// HBM/L2-bound streaming reductions: 128-bit loads where alignment allows, fp32 math in log2
// units (one FFMA + one MUFU.EX2 per element), warp-shuffle + shared-memory block reductions in
// a fixed order (deterministic; no fp32 atomics).
#include "common.cuh"
#include "sm100.cuh"
#include <stdlib.h>

namespace dinox {

constexpr int kRowThreads = 512;
constexpr int kMaxGlobalViews = 4;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&o)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[4]) {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&o)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = v;
  }
};
template <> struct Vec4<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&o)[4]) {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&v.x));
    float2 b = __half22float2(*reinterpret_cast<__half2*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&o)[4]) {
    __half2 a = __floats2half2_rn(o[0], o[1]);
    __half2 b = __floats2half2_rn(o[2], o[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = v;
  }
};

// loads 4 consecutive elements starting at column k (vector path) or up to `n` scalars (tail / unaligned)
template <typename T, bool kVec>
__device__ __forceinline__ void load4(const T* row, int64_t k, int64_t K, float (&o)[4], float fill) {
  if (kVec) {
    Vec4<T>::load(row + k, o);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (k + j < K) ? to_f32<T>(row[k + j]) : fill;
  }
}
template <bool kVec>
__device__ __forceinline__ void loadf4(const float* p, int64_t k, int64_t K, float (&o)[4], float fill) {
  if (p == nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = fill;
  } else {
    load4<float, kVec>(p, k, K, o, fill);
  }
}

// ---------------------------------------------------------------------------------------------
// rows_lse: one CTA per row.  lse[i] = ln sum_k exp(u[i,k]),  entropy[i] = lse - sum_k p_k u_k
// ---------------------------------------------------------------------------------------------
// kThreads = 512 while all rows fit in one wave of 4 CTAs per SM, else 256 (8 CTAs per SM: 640 rows at 512 threads
// were 1.08 waves - the second wave ran on 48 CTAs).  Two 16-byte loads per thread in flight.
template <typename T, bool kVec, int kThreads>
__global__ void __launch_bounds__(kThreads)
rows_lse_kernel(const T* __restrict__ x, int64_t K, int64_t ld, float scale2 /* inv_tau*log2e */,
                const float* __restrict__ colbias, float* __restrict__ lse, float* __restrict__ entropy) {
  __shared__ float red[64];
  const T* row = x + (int64_t)blockIdx.x * ld;
  float m = -INFINITY, s = 0.f, e = 0.f;
  const bool want_ent = entropy != nullptr;
  constexpr int kU = 2;   // 16-byte loads in flight per thread (four need 60 registers: 640 rows no longer fit one wave)
  for (int64_t k = (int64_t)threadIdx.x * 4; k < K; k += (int64_t)kThreads * 4 * kU) {
    float v[kU][4], cb[kU][4];
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      const int64_t ki = k + (int64_t)i * kThreads * 4;
      if (ki < K) {
        load4<T, kVec>(row, ki, K, v[i], 0.f);
        loadf4<kVec>(colbias, ki, K, cb[i], 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[i][j] = 0.f; cb[i][j] = 0.f; }
      }
    }
    float u[kU * 4];
    float mv = -INFINITY;
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      const int64_t ki = k + (int64_t)i * kThreads * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool in = kVec ? (ki < K) : (ki + j < K);
        u[i * 4 + j] = in ? fmaf(v[i][j], scale2, -cb[i][j] * DINOX_LOG2E) : -INFINITY;
        mv = fmaxf(mv, u[i * 4 + j]);
      }
    }
    float mn = fmaxf(m, mv);
    if (mn == -INFINITY) continue;
    float r = exp2f(m - mn);  // m == -inf -> 0
    s *= r;
    e *= r;
#pragma unroll
    for (int j = 0; j < kU * 4; ++j) {
      float p = exp2f(u[j] - mn);
      s += p;
      if (want_ent) e += (u[j] == -INFINITY) ? 0.f : p * u[j];
    }
    m = mn;
  }
  // block combine
  MaxSum ms{m, s};
  MaxSum tot = block_maxsum<kThreads>(ms, red);
  if (want_ent) {
    float escaled = (m == -INFINITY) ? 0.f : e * exp2f(m - tot.m);
    float etot = block_sum<kThreads>(escaled, red);
    if (threadIdx.x == 0) {
      float l2 = tot.m + log2f(tot.s);
      entropy[blockIdx.x] = DINOX_LN2 * (l2 - etot / tot.s);
    }
  }
  if (threadIdx.x == 0) lse[blockIdx.x] = DINOX_LN2 * (tot.m + log2f(tot.s));
}

// ---------------------------------------------------------------------------------------------
// column passes: each thread owns 4 adjacent columns and walks all rows
// ---------------------------------------------------------------------------------------------
constexpr int kColGroups = 16;   // row groups per block: 32 column-threads (4 columns each) x 16 row groups

// Block = 32 column-threads x kColGroups row groups.  Every row group walks its rows (stride kColGroups, two loads
// in flight) with a private online (max, sum 2^x) per column; the groups are merged through shared memory in a
// fixed order.  (One thread walking all rows of its columns serially was latency bound at 0.5 TB/s.)
template <typename T, bool kVec>
__global__ void __launch_bounds__(32 * kColGroups)
cols_lse_kernel(const T* __restrict__ x, int64_t rows, int64_t K, int64_t ld, float scale2,
                const float* __restrict__ rowbias, float* __restrict__ out) {
  __shared__ float sm[kColGroups][32][4], ss[kColGroups][32][4];
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int64_t k = ((int64_t)blockIdx.x * 32 + cx) * 4;
  float m[4], s[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { m[j] = -INFINITY; s[j] = 0.f; }
  auto fold = [&](const float (&v)[4], float rb) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = fmaf(v[j], scale2, -rb);
      const float mn = fmaxf(m[j], u);
      if (mn != -INFINITY) {
        s[j] = s[j] * exp2f(m[j] - mn) + exp2f(u - mn);
        m[j] = mn;
      }
    }
  };
  if (k < K) {
    int64_t i = gy;
    for (; i + kColGroups < rows; i += 2 * kColGroups) {
      float v0[4], v1[4];
      load4<T, kVec>(x + i * ld, k, K, v0, 0.f);
      load4<T, kVec>(x + (i + kColGroups) * ld, k, K, v1, 0.f);
      const float r0 = rowbias ? rowbias[i] * DINOX_LOG2E : 0.f, r1 = rowbias ? rowbias[i + kColGroups] * DINOX_LOG2E : 0.f;
      fold(v0, r0);
      fold(v1, r1);
    }
    for (; i < rows; i += kColGroups) {
      float v0[4];
      load4<T, kVec>(x + i * ld, k, K, v0, 0.f);
      fold(v0, rowbias ? rowbias[i] * DINOX_LOG2E : 0.f);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { sm[gy][cx][j] = m[j]; ss[gy][cx][j] = s[j]; }
  __syncthreads();
  if (gy == 0 && k < K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      MaxSum a{-INFINITY, 0.f};
      for (int g = 0; g < kColGroups; ++g) a = maxsum_merge(a, MaxSum{sm[g][cx][j], ss[g][cx][j]});   // fixed order
      if (k + j < K) out[k + j] = DINOX_LN2 * (a.m + log2f(a.s));
    }
  }
}

// column sums: block = 32 column-threads (4 columns each) x kRowGroups row groups; the row groups
// stride over the rows (4 loads in flight each) and are combined through shared memory in a fixed
// order.  Tall-skinny inputs (7424 x 384 head activations) still fill the SM this way.
constexpr int kColSumGroups = 16;

template <typename T, bool kVec>
__global__ void __launch_bounds__(32 * kColSumGroups)
cols_sum_kernel(const T* __restrict__ x, int64_t rows, int64_t K, int64_t ld, float* __restrict__ out, int64_t chunk,
                float scale = 1.f, const float* __restrict__ scale_dev = nullptr, int accumulate = 0) {
  __shared__ float red[kColSumGroups][32][4];
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  // blockIdx.y selects a chunk of rows; chunk sums land in row blockIdx.y of out (chunk >= rows: one chunk)
  x += (int64_t)blockIdx.y * chunk * ld;
  out += (int64_t)blockIdx.y * K;
  rows = rows - (int64_t)blockIdx.y * chunk < chunk ? rows - (int64_t)blockIdx.y * chunk : chunk;
  const int64_t k = ((int64_t)blockIdx.x * 32 + cx) * 4;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < K) {
    int64_t i = gy;
    for (; i + 3 * kColSumGroups < rows; i += 4 * kColSumGroups) {
      float v0[4], v1[4], v2[4], v3[4];
      load4<T, kVec>(x + (i + 0 * kColSumGroups) * ld, k, K, v0, 0.f);
      load4<T, kVec>(x + (i + 1 * kColSumGroups) * ld, k, K, v1, 0.f);
      load4<T, kVec>(x + (i + 2 * kColSumGroups) * ld, k, K, v2, 0.f);
      load4<T, kVec>(x + (i + 3 * kColSumGroups) * ld, k, K, v3, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; i < rows; i += kColSumGroups) {
      float v[4];
      load4<T, kVec>(x + i * ld, k, K, v, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[gy][cx][j] = a[j];
  __syncthreads();
  if (gy == 0 && k < K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < kColSumGroups; ++g) t += red[g][cx][j];
      if (k + j < K) {
        const float v = t * (scale * (scale_dev ? *scale_dev : 1.f));
        out[k + j] = accumulate ? out[k + j] + v : v;
      }
    }
  }
}

// Column sums of up to two row ranges of one matrix in ONE launch (the centre statistics of the fused head: sum of the
// teacher activations over the CLS rows and over the masked-patch rows).  Blocks sum 64-row chunks into a scratch
// slab; the LAST block of a (segment, column block) - found with a ticket counter - adds the chunk sums in chunk
// order, so the result does not depend on which block that is.  The ticket returns to 0: graph replays and
// later launches may reuse it without a memset.  Block (0, 0, seg) also copies the segment's row count.
struct SumSegs {
  int64_t begin[2], end[2];
};
constexpr int kSegChunk = 64;
template <typename T, bool kVec>
__global__ void __launch_bounds__(32 * kColSumGroups)
segment_cols_sum_kernel(const T* __restrict__ x, int64_t K, int64_t ld, SumSegs sg, float* __restrict__ out /* (nseg, K) */,
                        const float* __restrict__ counts_in, float* __restrict__ counts_out,
                        float* __restrict__ scratch /* (nseg, gridDim.y, K) */, unsigned* __restrict__ tickets) {
  __shared__ float red[kColSumGroups][32][4];
  __shared__ int is_last;
  const int cx = threadIdx.x & 31, gy = threadIdx.x >> 5;
  const int seg = blockIdx.z;
  const int64_t k = ((int64_t)blockIdx.x * 32 + cx) * 4;
  const int64_t r0 = sg.begin[seg] + (int64_t)blockIdx.y * kSegChunk;
  const int64_t r1 = r0 + kSegChunk < sg.end[seg] ? r0 + kSegChunk : sg.end[seg];
  const int nchunks = (int)((sg.end[seg] - sg.begin[seg] + kSegChunk - 1) / kSegChunk);
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < K) {
    for (int64_t i = r0 + gy; i < r1; i += kColSumGroups) {   // 64 rows / 16 groups: 4 independent loads per thread
      float v[4];
      load4<T, kVec>(x + i * ld, k, K, v, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[gy][cx][j] = a[j];
  __syncthreads();
  float* mine = scratch + ((int64_t)seg * gridDim.y + blockIdx.y) * K;
  if (gy == 0 && k < K && (int)blockIdx.y < nchunks) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < kColSumGroups; ++g) t += red[g][cx][j];
      if (k + j < K) mine[k + j] = t;
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && counts_out) counts_out[seg] = counts_in[seg];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* tk = tickets + seg * gridDim.x + blockIdx.x;
    const unsigned t = atomicAdd(tk, 1u);
    is_last = (t == gridDim.y - 1);
    if (is_last) *tk = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float b[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < K) {
    const float* base = scratch + (int64_t)seg * gridDim.y * K;
    for (int c = gy; c < nchunks; c += kColSumGroups) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < K) b[j] += __ldcg(base + (int64_t)c * K + k + j);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) red[gy][cx][j] = b[j];
  __syncthreads();
  if (gy == 0 && k < K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < kColSumGroups; ++g) t += red[g][cx][j];
      if (k + j < K) out[(int64_t)seg * K + k + j] = t;
    }
  }
}

__global__ void lse_combine_kernel(const float* __restrict__ g, int world, int64_t K, float add,
                                   float* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float m = -INFINITY;
  for (int r = 0; r < world; ++r) m = fmaxf(m, g[(int64_t)r * K + k]);
  float s = 0.f;
  for (int r = 0; r < world; ++r) s += (m == -INFINITY) ? 0.f : expf(g[(int64_t)r * K + k] - m);
  out[k] = m + logf(s) + add;
}

__global__ void center_ema_kernel(float* __restrict__ center, const float* __restrict__ colsum,
                                  float inv_rows, float m, float om, int64_t K) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  // reference op order (scripts/phase5_big_run.py:689): center*m + batch_center*(1-m)
  const float bc = colsum[k] * inv_rows;
  center[k] = __fadd_rn(__fmul_rn(center[k], m), __fmul_rn(bc, om));
}

__global__ void axpb_kernel(const float* __restrict__ a, float alpha, float beta,
                            float* __restrict__ out, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = fmaf(a[k], alpha, beta);
}

__global__ void axpby_kernel(const float* __restrict__ x, float alpha, const float* __restrict__ alpha_dev,
                             const float* __restrict__ y, float beta, float* __restrict__ out, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float a = alpha * (alpha_dev ? *alpha_dev : 1.f);
  if (k < n) out[k] = fmaf(x[k], a, (y ? y[k] : 0.f) * beta);
}

// ---------------------------------------------------------------------------------------------
// cross-entropy forward / backward over groups x K-splits
// ---------------------------------------------------------------------------------------------
constexpr int kCeThreads = 256;
// resident CTAs per SM the kernels are compiled for (register caps 85 / 64 / 128): the K-split is chosen against
// these so that the grid fills whole waves
constexpr int kCeFwdCtasPerSm = 3, kCeBwdCtasPerSm = 4, kOnePassCtasPerSm = 2;

struct CeArgs {
  int64_t groups, K, ld_s, ld_t, ld_g;
  int V, Vg, exclude_same, ksplit;
  float s2, t2;  // inv_tau * log2e
  const float* colbias_t;
  const float* rowbias_t;
  const float* lse_s;
  const float* group_w;
  float norm;
};

template <typename TS, typename TT, bool kVec>
__global__ void __launch_bounds__(kCeThreads, kCeFwdCtasPerSm)
ce_fwd_kernel(const TS* __restrict__ student, const TT* __restrict__ teacher, CeArgs a,
              float* __restrict__ partial /* [groups][ksplit][Vg][2] */) {
  __shared__ float red[64];
  const int64_t g = blockIdx.x;
  const int split = blockIdx.y;
  const int64_t per = ((a.K + a.ksplit - 1) / a.ksplit + 3) & ~int64_t(3);
  const int64_t k0 = split * per, k1 = (k0 + per < a.K) ? k0 + per : a.K;
  float rb2[kMaxGlobalViews];
#pragma unroll
  for (int q = 0; q < kMaxGlobalViews; ++q)
    rb2[q] = (q < a.Vg) ? a.rowbias_t[q * a.groups + g] * DINOX_LOG2E : 0.f;
  float qsum[kMaxGlobalViews], cross[kMaxGlobalViews];
#pragma unroll
  for (int q = 0; q < kMaxGlobalViews; ++q) { qsum[q] = 0.f; cross[q] = 0.f; }

  for (int64_t k = k0 + (int64_t)threadIdx.x * 4; k < k1; k += (int64_t)kCeThreads * 4) {
    float cb[4];
    loadf4<kVec>(a.colbias_t, k, k1, cb, 0.f);
    float qv[kMaxGlobalViews][4];
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) {
      if (q < a.Vg) {
        float t[4];
        load4<TT, kVec>(teacher + (q * a.groups + g) * a.ld_t, k, k1, t, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float u = fmaf(t[j], a.t2, -cb[j] * DINOX_LOG2E) - rb2[q];
          qv[q][j] = (!kVec && k + j >= k1) ? 0.f : exp2f(u);
        }
      }
    }
    float stot[4] = {0.f, 0.f, 0.f, 0.f};
    float sown[kMaxGlobalViews][4];
    // student views: the teacher-paired (global) ones are kept, all are summed; four row loads in flight
    for (int v0 = 0; v0 < a.V; v0 += 4) {
      float s[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v0 + u < a.V) load4<TS, kVec>(student + ((v0 + u) * a.groups + g) * a.ld_s, k, k1, s[u], 0.f);
        else { s[u][0] = s[u][1] = s[u][2] = s[u][3] = 0.f; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 4; ++j) stot[j] += s[u][j];
#pragma unroll
        for (int q = 0; q < kMaxGlobalViews; ++q)
          if (q == v0 + u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sown[q][j] = s[u][j];
          }
      }
    }
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) {
      if (q < a.Vg) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float other = a.exclude_same ? (stot[j] - sown[q][j]) : stot[j];
          qsum[q] += qv[q][j];
          cross[q] = fmaf(qv[q][j], other, cross[q]);
        }
      }
    }
  }
  for (int q = 0; q < a.Vg; ++q) {
    float qs = block_sum<kCeThreads>(qsum[q], red);
    float cr = block_sum<kCeThreads>(cross[q], red);
    if (threadIdx.x == 0) {
      float* p = partial + (((int64_t)g * a.ksplit + split) * a.Vg + q) * 2;
      p[0] = qs;
      p[1] = cr;
    }
  }
}

// single CTA, fixed summation order => deterministic loss
__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const float* __restrict__ partial, CeArgs a, float inv_tau_s, float* __restrict__ loss_out) {
  __shared__ float red[64];
  float acc = 0.f;
  for (int64_t g = threadIdx.x; g < a.groups; g += 1024) {
    float lg = 0.f;
    for (int q = 0; q < a.Vg; ++q) {
      float qs = 0.f, cr = 0.f;
      for (int sp = 0; sp < a.ksplit; ++sp) {
        const float* p = partial + ((g * a.ksplit + sp) * a.Vg + q) * 2;
        qs += p[0];
        cr += p[1];
      }
      float lse_sum = 0.f;
      for (int v = 0; v < a.V; ++v)
        if (!(a.exclude_same && v == q)) lse_sum += a.lse_s[v * a.groups + g];
      // sum over pairs of [ lse_v * sum_k q - sum_k q * us ] ; us = s*inv_tau_s
      lg += qs * lse_sum - cr * inv_tau_s;
    }
    acc += lg * (a.group_w ? a.group_w[g] : 1.f);
  }
  float tot = block_sum<1024>(acc, red);
  if (threadIdx.x == 0) *loss_out = tot * a.norm;
}

// ---------------------------------------------------------------------------------------------
// cross-entropy forward in ONE pass over the logits (softmax-centred teacher): the student LSEs, the teacher
// LSEs and the cross terms sum_k 2^(u_t[k]) * s[k] are all accumulated online (running maximum + rescale) while
// every logit is read exactly once - DINOLoss.forward (scripts/phase5_big_run.py:703-717) without the separate
// softmax / log_softmax passes.  The LSEs come out as by-products for the backward kernel.
//   partial[g][split][ q < Vg : (m, z, c) | v < V : (m, s) ]   all in log2 units, merged in split order
// ---------------------------------------------------------------------------------------------
constexpr int kOnePassMaxViews = 12;

__device__ __forceinline__ float ex2_approx(float x) {   // one MUFU.EX2; 2^-inf = 0, NaN stays NaN
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (max, sum, cross) of one row across the warp: rescale to the common maximum, then plain sums
__device__ __forceinline__ void warp_merge3(float& m, float& z, float& c) {
  const float M = warp_max(m);
  const float sc = (m == -INFINITY) ? 0.f : exp2f(m - M);
  z = warp_sum(z * sc);
  c = warp_sum(c * sc);
  m = M;
}

template <typename TS, typename TT, bool kVec, int kMaxV>
__global__ void __launch_bounds__(kCeThreads, kMaxV <= 4 ? kCeFwdCtasPerSm : kOnePassCtasPerSm)
ce_fwd_onepass_kernel(const TS* __restrict__ student, const TT* __restrict__ teacher, CeArgs a,
                      float* __restrict__ partial /* [groups][ksplit][3 Vg + 2 V] */) {
  static_assert(kMaxV % 4 == 0 && kMaxV >= kMaxGlobalViews, "views are walked four at a time");
  constexpr int kWarps = kCeThreads / 32;
  constexpr int kVals = 3 * kMaxGlobalViews + 2 * kMaxV;
  __shared__ float red[kWarps][kVals];
  const int64_t g = blockIdx.x;
  const int split = blockIdx.y;
  const int64_t per = ((a.K + a.ksplit - 1) / a.ksplit + 3) & ~int64_t(3);
  const int64_t k0 = split * per, k1 = (k0 + per < a.K) ? k0 + per : a.K;
  float tm[kMaxGlobalViews], tz[kMaxGlobalViews], tc[kMaxGlobalViews];
#pragma unroll
  for (int q = 0; q < kMaxGlobalViews; ++q) { tm[q] = -INFINITY; tz[q] = 0.f; tc[q] = 0.f; }
  float sm[kMaxV], ss[kMaxV];
#pragma unroll
  for (int v = 0; v < kMaxV; ++v) { sm[v] = -INFINITY; ss[v] = 0.f; }
  const uint32_t tvs = (uint32_t)(a.groups * a.ld_t), svs = (uint32_t)(a.groups * a.ld_s);   // elements between two views
  const uint32_t tg0 = (uint32_t)(g * a.ld_t + k0), sg0 = (uint32_t)(g * a.ld_s + k0);        // (view 0, group g, column k0)

  for (int64_t k = k0 + (int64_t)threadIdx.x * 4; k < k1; k += (int64_t)kCeThreads * 4) {
    // phase A: every load of this 4-column block is issued before any arithmetic (predicated, no branches), so one
    // trip to memory covers all rows - processing row by row behind its own load serialised five round trips
    float cb[4];
    loadf4<kVec>(a.colbias_t, k, k1, cb, 0.f);
    float tr[kMaxGlobalViews][4], sr[kMaxV][4];
    // row (q, g), column k = base + (g * ld + k) + q * view stride, all in 32-bit element offsets (the host checks that
    // both matrices hold fewer than 2^32 elements): one IMAD + one IMAD.WIDE per row - 64-bit products per row and
    // block were 250 of the loop's 880 instructions
    const uint32_t tk = tg0 + (uint32_t)(k - k0), sk = sg0 + (uint32_t)(k - k0);
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) {
      if (q < a.Vg) {
        if (kVec) Vec4<TT>::load(teacher + (tk + (uint32_t)q * tvs), tr[q]);
        else load4<TT, false>(teacher + (tg0 - (uint32_t)k0 + (uint32_t)q * tvs), k, k1, tr[q], 0.f);
      } else { tr[q][0] = tr[q][1] = tr[q][2] = tr[q][3] = 0.f; }
    }
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) {
      if (v < a.V) {
        if (kVec) Vec4<TS>::load(student + (sk + (uint32_t)v * svs), sr[v]);
        else load4<TS, false>(student + (sg0 - (uint32_t)k0 + (uint32_t)v * svs), k, k1, sr[v], 0.f);
      } else { sr[v][0] = sr[v][1] = sr[v][2] = sr[v][3] = 0.f; }
    }
    // phase B.  Student rows: the maximum is taken on the raw logits (inv_tau_s > 0, so it commutes with the scale)
    // and the scale rides in the FFMA in front of each exponential: ~25 instructions per row and block instead of ~70
    // with one libm exp2f per element (MUFU.EX2 through ex2.approx.ftz: 2^-22 relative, far inside the 1e-5 budget)
    float stot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) {
      if (v < a.V) {
        float raw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          stot[j] += sr[v][j];
          raw[j] = (!kVec && k + j >= k1) ? -INFINITY : sr[v][j];
        }
        const float mn = fmaxf(sm[v], a.s2 * fmaxf(fmaxf(raw[0], raw[1]), fmaxf(raw[2], raw[3])));
        const float mref = (mn == -INFINITY) ? 0.f : mn;   // nothing finite yet: every term below is 2^-inf = 0
        float acc = ss[v] * ex2_approx(sm[v] - mref);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += ex2_approx(fmaf(raw[j], a.s2, -mref));
        ss[v] = acc;
        sm[v] = mn;
      }
    }
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) {
      if (q < a.Vg) {
        float u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          u[j] = (!kVec && k + j >= k1) ? -INFINITY : fmaf(tr[q][j], a.t2, -cb[j] * DINOX_LOG2E);
        const float mn = fmaxf(tm[q], fmaxf(fmaxf(u[0], u[1]), fmaxf(u[2], u[3])));
        const float mref = (mn == -INFINITY) ? 0.f : mn;
        const float r = ex2_approx(tm[q] - mref);
        float z = tz[q] * r, c = tc[q] * r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // the student views this teacher view pairs with: all, or all but its own (row q of the student block)
          const float other = a.exclude_same ? (stot[j] - sr[q < kMaxV ? q : 0][j]) : stot[j];
          const float p = ex2_approx(u[j] - mref);
          z += p;
          c = fmaf(p, other, c);
        }
        tz[q] = z; tc[q] = c; tm[q] = mn;
      }
    }
  }
  // block merge: warp shuffles, then one thread per row walks the warps in order
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < kMaxGlobalViews; ++q) {
    if (q < a.Vg) {
      warp_merge3(tm[q], tz[q], tc[q]);
      if (lane == 0) { red[w][3 * q] = tm[q]; red[w][3 * q + 1] = tz[q]; red[w][3 * q + 2] = tc[q]; }
    }
  }
#pragma unroll
  for (int v = 0; v < kMaxV; ++v) {
    if (v < a.V) {
      float dummy = 0.f;
      warp_merge3(sm[v], ss[v], dummy);
      if (lane == 0) { red[w][3 * kMaxGlobalViews + 2 * v] = sm[v]; red[w][3 * kMaxGlobalViews + 2 * v + 1] = ss[v]; }
    }
  }
  __syncthreads();
  float* out = partial + ((int64_t)g * a.ksplit + split) * (3 * a.Vg + 2 * a.V);
  const int t = threadIdx.x;
  if (t < a.Vg) {
    float M = -INFINITY;
    for (int i = 0; i < kWarps; ++i) M = fmaxf(M, red[i][3 * t]);
    float z = 0.f, c = 0.f;
    for (int i = 0; i < kWarps; ++i) {
      const float m = red[i][3 * t];
      const float sc = (m == -INFINITY) ? 0.f : exp2f(m - M);
      z += red[i][3 * t + 1] * sc;
      c += red[i][3 * t + 2] * sc;
    }
    out[3 * t] = M; out[3 * t + 1] = z; out[3 * t + 2] = c;
  } else if (t >= 32 && t < 32 + a.V) {
    const int v = t - 32;
    float M = -INFINITY;
    for (int i = 0; i < kWarps; ++i) M = fmaxf(M, red[i][3 * kMaxGlobalViews + 2 * v]);
    float s = 0.f;
    for (int i = 0; i < kWarps; ++i) {
      const float m = red[i][3 * kMaxGlobalViews + 2 * v];
      s += red[i][3 * kMaxGlobalViews + 2 * v + 1] * ((m == -INFINITY) ? 0.f : exp2f(m - M));
    }
    out[3 * a.Vg + 2 * v] = M; out[3 * a.Vg + 2 * v + 1] = s;
  }
}

// one CTA.  Phase 1: one thread per (group, row) merges that row's K-splits in split order (online max / rescale) and
// writes the natural-log LSE - the backward kernel's inputs - and, for teacher rows, the normalised cross term c / z.
// Phase 2: one thread per group forms the pair sums.  Fixed summation order throughout => deterministic.
__global__ void __launch_bounds__(1024)
ce_onepass_finalize_kernel(const float* __restrict__ partial, CeArgs a, float inv_tau_s, float* __restrict__ loss_out,
                           float* lse_s_out, float* rowbias_t_out, float* cz /* (Vg, groups) scratch */) {
  __shared__ float red[64];
  const int stride = 3 * a.Vg + 2 * a.V;
  const int R = a.V + a.Vg;
  const int64_t items = a.groups * R;
  for (int64_t it = threadIdx.x; it < items; it += 1024) {
    const int64_t g = it / R;
    const int r = (int)(it - g * R);
    const float* pg = partial + g * a.ksplit * stride;
    if (r < a.V) {
      const float* p = pg + 3 * a.Vg + 2 * r;
      float M = -INFINITY, S = 0.f;
#pragma unroll 4
      for (int sp = 0; sp < a.ksplit; ++sp) {
        const float m = p[sp * stride], sv = p[sp * stride + 1];
        const float mn = fmaxf(M, m);
        const float mref = (mn == -INFINITY) ? 0.f : mn;
        S = S * exp2f(M - mref) + sv * exp2f(m - mref);
        M = mn;
      }
      lse_s_out[r * a.groups + g] = DINOX_LN2 * (M + log2f(S));
    } else {
      const int q = r - a.V;
      const float* p = pg + 3 * q;
      float M = -INFINITY, Z = 0.f, C = 0.f;
#pragma unroll 4
      for (int sp = 0; sp < a.ksplit; ++sp) {
        const float m = p[sp * stride], zv = p[sp * stride + 1], cv = p[sp * stride + 2];
        const float mn = fmaxf(M, m);
        const float mref = (mn == -INFINITY) ? 0.f : mn;
        const float r0 = exp2f(M - mref), r1 = exp2f(m - mref);
        Z = Z * r0 + zv * r1;
        C = C * r0 + cv * r1;
        M = mn;
      }
      rowbias_t_out[q * a.groups + g] = DINOX_LN2 * (M + log2f(Z));
      cz[q * a.groups + g] = C / Z;   // sum_k q[k] * (sum of the paired student rows)[k], q = 2^(u - M) / Z
    }
  }
  __syncthreads();   // one CTA: the global writes above are visible to the whole block
  float acc = 0.f;
  for (int64_t g = threadIdx.x; g < a.groups; g += 1024) {
    float lse_tot = 0.f;
    for (int v = 0; v < a.V; ++v) lse_tot += lse_s_out[v * a.groups + g];
    float lg = 0.f;
    for (int q = 0; q < a.Vg; ++q) {
      // sum over the pairs of this teacher view of [ lse_v - (1/tau_s) sum_k q[k] s_v[k] ]
      const float lse_sum = a.exclude_same ? lse_tot - lse_s_out[q * a.groups + g] : lse_tot;
      lg += lse_sum - cz[q * a.groups + g] * inv_tau_s;
    }
    acc += lg * (a.group_w ? a.group_w[g] : 1.f);
  }
  const float tot = block_sum<1024>(acc, red);
  if (threadIdx.x == 0) *loss_out = tot * a.norm;
}

// ---------------------------------------------------------------------------------------------
// The same one-pass forward as a PERSISTENT, shared-memory staged stream (the default when alignment allows):
// one CTA per SM; a producer thread copies the next 2048-column block of every row of a (group, K-split) work item
// into shared memory with 1-D bulk copies (cp.async.bulk, mbarrier complete_tx) while 512 consumer threads run the
// online LSE / cross-term arithmetic on the previous block.  Loads need no registers, run kStages blocks ahead and
// straight across work-item boundaries, so no CTA start-up latency is exposed after the first block; the consumers
// hold only the running (max, sum, cross) state of the rows (96 registers per thread in all).  Same partial layout and finalize kernel as the register form.
// ---------------------------------------------------------------------------------------------
constexpr int kStreamConsumers = 512;
constexpr int kStreamThreads = kStreamConsumers + 32;
constexpr int kStreamCols = kStreamConsumers * 4;       // columns per stage
constexpr int kStreamMaxStages = 4;
constexpr int kStreamSmemBudget = 220 * 1024;           // of the 227 KB a CTA may own

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sm100::smem_u32(dst)), "l"(src), "r"(bytes), "r"(sm100::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumers) : "memory"); }

struct StreamArgs {
  int stages;            // 2..kStreamMaxStages
  int64_t per;           // columns per K-split (multiple of 8)
  int64_t items;         // groups * ksplit
  uint32_t row_s, row_t; // bytes of one staged student / teacher row (kStreamCols elements)
  uint32_t off_t, off_cb, stage_bytes;   // teacher rows / column offsets inside a stage; bytes per stage
};

template <typename TS, typename TT, int kMaxV>
__global__ void __launch_bounds__(kStreamThreads, 1)
ce_fwd_onepass_stream_kernel(const TS* __restrict__ student, const TT* __restrict__ teacher, CeArgs a, StreamArgs sa,
                             float* __restrict__ partial /* [groups][ksplit][3 Vg + 2 V] */) {
  static_assert(kMaxV % 4 == 0 && kMaxV >= kMaxGlobalViews, "views");
  constexpr int kWarps = kStreamConsumers / 32;
  constexpr int kVals = 3 * kMaxGlobalViews + 2 * kMaxV;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[kStreamMaxStages], empty[kStreamMaxStages];
  __shared__ float red[2][kWarps][kVals];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStreamMaxStages; ++i) { sm100::mbar_init(&full[i], 1); sm100::mbar_init(&empty[i], kWarps); }
    sm100::fence_barrier_init();
  }
  __syncthreads();
  const int stride = 3 * a.Vg + 2 * a.V;

  if (threadIdx.x >= kStreamConsumers) {
    // ---------------- producer: one thread walks (item, block) and keeps `stages` blocks in flight ----------------
    if (threadIdx.x != kStreamConsumers) return;
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t item = blockIdx.x; item < sa.items; item += gridDim.x) {
      const int64_t g = item / a.ksplit;
      const int split = (int)(item - g * a.ksplit);
      const int64_t k0 = split * sa.per, k1 = (k0 + sa.per < a.K) ? k0 + sa.per : a.K;
      for (int64_t k = k0; k < k1; k += kStreamCols) {
        const uint32_t cols = (uint32_t)((k1 - k < kStreamCols) ? (k1 - k) : kStreamCols);
        sm100::mbar_wait(&empty[stage], phase ^ 1u, 31);
        uint8_t* st = smem + (size_t)stage * sa.stage_bytes;
        const uint32_t bs = cols * (uint32_t)sizeof(TS), bt = cols * (uint32_t)sizeof(TT);
        sm100::mbar_expect_tx(&full[stage], (uint32_t)a.V * bs + (uint32_t)a.Vg * bt + (a.colbias_t ? cols * 4u : 0u));
        for (int v = 0; v < a.V; ++v)
          bulk_g2s(st + (size_t)v * sa.row_s, student + (v * a.groups + g) * a.ld_s + k, bs, &full[stage]);
        for (int q = 0; q < a.Vg; ++q)
          bulk_g2s(st + sa.off_t + (size_t)q * sa.row_t, teacher + (q * a.groups + g) * a.ld_t + k, bt, &full[stage]);
        if (a.colbias_t) bulk_g2s(st + sa.off_cb, a.colbias_t + k, cols * 4u, &full[stage]);
        if (++stage == sa.stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  // Arithmetic in packed fp32 pairs (FFMA2 / FADD2, sm_100); the running maximum of a row is only touched when a block
  // raises it (rare after the first blocks), otherwise a row costs one LDS.128, three FMNMX, one compare and, per pair
  // of prototypes, one FFMA2 + two MUFU.EX2 + one FADD2.  -1e30 stands for "no maximum yet": 2^(x + 1e30) never
  // appears because the first finite block raises the maximum before any term is formed, and -inf inputs give 0.
  constexpr float kNoMax = -1.0e30f;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int stage = 0;
  uint32_t phase = 0;
  int flip = 0;
  const float2 s22 = make_float2(a.s2, a.s2), t22 = make_float2(a.t2, a.t2);
  for (int64_t item = blockIdx.x; item < sa.items; item += gridDim.x) {
    const int64_t g = item / a.ksplit;
    const int split = (int)(item - g * a.ksplit);
    const int64_t k0 = split * sa.per, k1 = (k0 + sa.per < a.K) ? k0 + sa.per : a.K;
    float tm[kMaxGlobalViews];
    float2 tz[kMaxGlobalViews], tc[kMaxGlobalViews];
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) { tm[q] = kNoMax; tz[q] = make_float2(0.f, 0.f); tc[q] = make_float2(0.f, 0.f); }
    float sm[kMaxV];
    float2 ss[kMaxV];
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) { sm[v] = kNoMax; ss[v] = make_float2(0.f, 0.f); }

    for (int64_t k = k0; k < k1; k += kStreamCols) {
      const uint32_t cols = (uint32_t)((k1 - k < kStreamCols) ? (k1 - k) : kStreamCols);
      sm100::mbar_wait(&full[stage], phase, 32);
      const uint8_t* st = smem + (size_t)stage * sa.stage_bytes;
      if ((uint32_t)t * 4u < cols) {
        float2 stot0 = make_float2(0.f, 0.f), stot1 = make_float2(0.f, 0.f);
        const uint8_t* srow = st + (size_t)t * (4 * sizeof(TS));
#pragma unroll
        for (int v = 0; v < kMaxV; ++v) {
          if (v < a.V) {
            float raw[4];
            Vec4<TS>::load(reinterpret_cast<const TS*>(srow), raw);
            srow += sa.row_s;
            const float2 r0 = make_float2(raw[0], raw[1]), r1 = make_float2(raw[2], raw[3]);
            stot0 = __fadd2_rn(stot0, r0);
            stot1 = __fadd2_rn(stot1, r1);
            const float cand = a.s2 * fmaxf(fmaxf(raw[0], raw[1]), fmaxf(raw[2], raw[3]));
            if (cand > sm[v]) {   // the block raises this thread's maximum of the row: rescale the running sums
              const float r = ex2_approx(sm[v] - cand);   // first block: 2^(-1e30 - cand) = 0
              ss[v] = __fmul2_rn(ss[v], make_float2(r, r));
              sm[v] = cand;
            }
            const float2 nm = make_float2(-sm[v], -sm[v]);
            const float2 x0 = __ffma2_rn(r0, s22, nm), x1 = __ffma2_rn(r1, s22, nm);
            ss[v] = __fadd2_rn(ss[v], make_float2(ex2_approx(x0.x), ex2_approx(x0.y)));
            ss[v] = __fadd2_rn(ss[v], make_float2(ex2_approx(x1.x), ex2_approx(x1.y)));
          }
        }
        float2 cb0 = make_float2(0.f, 0.f), cb1 = make_float2(0.f, 0.f);
        if (a.colbias_t) {
          float cb[4];
          Vec4<float>::load(reinterpret_cast<const float*>(st + sa.off_cb) + 4 * t, cb);
          cb0 = make_float2(-cb[0] * DINOX_LOG2E, -cb[1] * DINOX_LOG2E);
          cb1 = make_float2(-cb[2] * DINOX_LOG2E, -cb[3] * DINOX_LOG2E);
        }
        const uint8_t* trow = st + sa.off_t + (size_t)t * (4 * sizeof(TT));
        const uint8_t* orow = st + (size_t)t * (4 * sizeof(TS));
#pragma unroll
        for (int q = 0; q < kMaxGlobalViews; ++q) {
          if (q < a.Vg) {
            float tr[4], own[4] = {0.f, 0.f, 0.f, 0.f};
            Vec4<TT>::load(reinterpret_cast<const TT*>(trow), tr);
            if (a.exclude_same) Vec4<TS>::load(reinterpret_cast<const TS*>(orow), own);
            trow += sa.row_t;
            orow += sa.row_s;
            const float2 u0 = __ffma2_rn(make_float2(tr[0], tr[1]), t22, cb0), u1 = __ffma2_rn(make_float2(tr[2], tr[3]), t22, cb1);
            const float cand = fmaxf(fmaxf(u0.x, u0.y), fmaxf(u1.x, u1.y));
            if (cand > tm[q]) {
              const float r = ex2_approx(tm[q] - cand);
              const float2 r2 = make_float2(r, r);
              tz[q] = __fmul2_rn(tz[q], r2);
              tc[q] = __fmul2_rn(tc[q], r2);
              tm[q] = cand;
            }
            const float2 nm = make_float2(-tm[q], -tm[q]);
            const float2 e0 = __fadd2_rn(u0, nm), e1 = __fadd2_rn(u1, nm);
            const float2 p0 = make_float2(ex2_approx(e0.x), ex2_approx(e0.y)), p1 = make_float2(ex2_approx(e1.x), ex2_approx(e1.y));
            // the student views this teacher view pairs with: all, or all but its own (row q of the student block)
            const float2 o0 = __fadd2_rn(stot0, make_float2(-own[0], -own[1])), o1 = __fadd2_rn(stot1, make_float2(-own[2], -own[3]));
            tz[q] = __fadd2_rn(tz[q], __fadd2_rn(p0, p1));
            tc[q] = __ffma2_rn(p0, o0, tc[q]);
            tc[q] = __ffma2_rn(p1, o1, tc[q]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&empty[stage]);   // this warp has read the stage
      if (++stage == sa.stages) { stage = 0; phase ^= 1u; }
    }
    // merge of the work item: warp shuffles, one thread per row walks the warps in order (fixed order)
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) {
      if (q < a.Vg) {
        float m = tm[q], z = tz[q].x + tz[q].y, c = tc[q].x + tc[q].y;
        warp_merge3(m, z, c);
        if (lane == 0) { red[flip][w][3 * q] = m; red[flip][w][3 * q + 1] = z; red[flip][w][3 * q + 2] = c; }
      }
    }
#pragma unroll
    for (int v = 0; v < kMaxV; ++v) {
      if (v < a.V) {
        float m = sm[v], sv = ss[v].x + ss[v].y, dummy = 0.f;
        warp_merge3(m, sv, dummy);
        if (lane == 0) { red[flip][w][3 * kMaxGlobalViews + 2 * v] = m; red[flip][w][3 * kMaxGlobalViews + 2 * v + 1] = sv; }
      }
    }
    consumer_sync();   // red[flip] complete; red[flip ^ 1] (previous item) has been read by everyone who passes here
    float* out = partial + item * stride;
    if (t < a.Vg) {
      float M = -INFINITY;
      for (int i = 0; i < kWarps; ++i) M = fmaxf(M, red[flip][i][3 * t]);
      float z = 0.f, c = 0.f;
      for (int i = 0; i < kWarps; ++i) {
        const float sc = exp2f(red[flip][i][3 * t] - M);   // maxima are finite (>= -1e30): no inf - inf
        z += red[flip][i][3 * t + 1] * sc;
        c += red[flip][i][3 * t + 2] * sc;
      }
      out[3 * t] = M; out[3 * t + 1] = z; out[3 * t + 2] = c;
    } else if (t >= 32 && t < 32 + a.V) {
      const int v = t - 32;
      float M = -INFINITY;
      for (int i = 0; i < kWarps; ++i) M = fmaxf(M, red[flip][i][3 * kMaxGlobalViews + 2 * v]);
      float sv = 0.f;
      for (int i = 0; i < kWarps; ++i)
        sv += red[flip][i][3 * kMaxGlobalViews + 2 * v + 1] * exp2f(red[flip][i][3 * kMaxGlobalViews + 2 * v] - M);
      out[3 * a.Vg + 2 * v] = M; out[3 * a.Vg + 2 * v + 1] = sv;
    }
    flip ^= 1;
  }
}

// kBatch student rows are loaded together with the teacher rows before any arithmetic (kBatch = 12 covers every row
// of up to 12 views in ONE trip to memory per 4-column block; more views take further batches), then written back
// row by row.  kBatch = 4 is the small-register form for <= 4 views.
template <typename TS, typename TT, bool kVec, int kBatch>
__global__ void __launch_bounds__(kCeThreads, kBatch <= 4 ? kCeBwdCtasPerSm : kOnePassCtasPerSm)
ce_bwd_kernel(const TS* __restrict__ student, const TT* __restrict__ teacher, CeArgs a,
              const float* __restrict__ upstream, TS* __restrict__ grad) {
  const int64_t g = blockIdx.x;
  const int split = blockIdx.y;
  const int64_t per = ((a.K + a.ksplit - 1) / a.ksplit + 3) & ~int64_t(3);
  const int64_t k0 = split * per, k1 = (k0 + per < a.K) ? k0 + per : a.K;
  const float w = (a.group_w ? a.group_w[g] : 1.f) * a.norm * (*upstream) * (a.s2 * DINOX_LN2);
  float rb2[kMaxGlobalViews];
#pragma unroll
  for (int q = 0; q < kMaxGlobalViews; ++q)
    rb2[q] = (q < a.Vg) ? a.rowbias_t[q * a.groups + g] * DINOX_LOG2E : 0.f;
  const TT* trow0 = teacher + g * a.ld_t;
  const TS* srow0 = student + g * a.ld_s;
  TS* grow0 = grad + g * a.ld_g;
  const int64_t tvs = a.groups * a.ld_t, svs = a.groups * a.ld_s, gvs = a.groups * a.ld_g;   // view strides
  float lse2_first[kBatch];   // log2 LSEs of the first batch of student rows: constant over the K sweep
#pragma unroll
  for (int u = 0; u < kBatch; ++u) lse2_first[u] = (u < a.V) ? a.lse_s[u * a.groups + g] * DINOX_LOG2E : 0.f;

  for (int64_t k = k0 + (int64_t)threadIdx.x * 4; k < k1; k += (int64_t)kCeThreads * 4) {
    // phase A: teacher rows, column offsets and the first batch of student rows - all loads before any arithmetic
    float cb[4];
    loadf4<kVec>(a.colbias_t, k, k1, cb, 0.f);
    float tr[kMaxGlobalViews][4];
    const TT* tq = trow0;   // row (q, g) = row (0, g) + q * view stride: no 64-bit multiplies in the loop
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) {
      if (q < a.Vg) load4<TT, kVec>(tq, k, k1, tr[q], 0.f);
      else { tr[q][0] = tr[q][1] = tr[q][2] = tr[q][3] = 0.f; }
      tq += tvs;
    }
    float sb[kBatch][4];
    const TS* sv = srow0;
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (u < a.V) load4<TS, kVec>(sv, k, k1, sb[u], 0.f);
      else { sb[u][0] = sb[u][1] = sb[u][2] = sb[u][3] = 0.f; }
      sv += svs;
    }
    // phase B
    float qv[kMaxGlobalViews][4];
    float qtot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < kMaxGlobalViews; ++q) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float u = fmaf(tr[q][j], a.t2, -cb[j] * DINOX_LOG2E) - rb2[q];
        qv[q][j] = (q < a.Vg) ? exp2f(u) : 0.f;
        qtot[j] += qv[q][j];
      }
    }
    TS* gv = grow0;
    for (int v0 = 0; v0 < a.V; v0 += kBatch) {
      float lse2[kBatch];
      if (v0 > 0) {   // more views than one batch holds
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int64_t r = (int64_t)(v0 + u) * a.groups + g;
          if (v0 + u < a.V) {
            load4<TS, kVec>(srow0 + (int64_t)(v0 + u) * svs, k, k1, sb[u], 0.f);
            lse2[u] = a.lse_s[r] * DINOX_LOG2E;
          } else {
            sb[u][0] = sb[u][1] = sb[u][2] = sb[u][3] = 0.f;
            lse2[u] = 0.f;
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < kBatch; ++u) lse2[u] = lse2_first[u];
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int v = v0 + u;
        if (v >= a.V) break;
        const bool own = a.exclude_same && v < a.Vg;
        const float nq = (float)(a.Vg - (own ? 1 : 0));
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float p = exp2f(fmaf(sb[u][j], a.s2, -lse2[u]));
          float target = qtot[j];
#pragma unroll
          for (int q = 0; q < kMaxGlobalViews; ++q)
            if (own && q == v) target -= qv[q][j];
          o[j] = w * (nq * p - target);
        }
        if (kVec) {
          Vec4<TS>::store(gv + k, o);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k + j < k1) gv[k + j] = from_f32<TS>(o[j]);
        }
        gv += gvs;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch helpers
// ---------------------------------------------------------------------------------------------
static inline size_t elt_size(int dtype) { return dtype == DINOX_F32 ? 4 : 2; }
static inline bool vec_ok(const void* p, int dtype, int64_t K, int64_t ld) {
  return (K % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) % (4 * elt_size(dtype))) == 0);
}
static inline bool fvec_ok(const float* p) { return p == nullptr || aligned16(p); }

// K-split of the (groups x splits) grids: the smallest split count whose grid fills whole waves of the resident CTAs
// best (at most 4 waves; each split keeps >= 1024 prototypes).  A fixed "2 CTAs per SM" target left 3/4 of the
// machine's threads unused (ce_fwd 2.3 TB/s, ce_bwd 2.9 TB/s at the C2 shapes); a fixed 8 per SM ran 2.05 waves.
static int pick_ksplit(int64_t groups, int64_t K, int ctas_per_sm) {
  const int64_t resident = (int64_t)num_sms() * ctas_per_sm;
  int64_t maxs = K / 1024;
  if (maxs < 1) maxs = 1;
  if (maxs > 64) maxs = 64;
  int best = 1;
  double best_eff = 0.0;
  for (int64_t ks = 1; ks <= maxs; ++ks) {
    const int64_t total = groups * ks;
    const int64_t waves = (total + resident - 1) / resident;
    if (waves > 4 && ks > 1) break;
    const double eff = (double)total / (double)(waves * resident);
    if (eff > best_eff + 0.02) { best_eff = eff; best = (int)ks; }
  }
  return best;
}

#define DISPATCH_T(dtype, T, ...)                                  \
  switch (dtype) {                                                 \
    case DINOX_F32: { using T = float; __VA_ARGS__; break; }       \
    case DINOX_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
    case DINOX_F16: { using T = __half; __VA_ARGS__; break; }      \
    default: set_error("unknown dtype code %d", dtype); return DINOX_E_BADARG; \
  }

}  // namespace dinox

extern "C" {
using namespace dinox;

int dinox_rows_lse(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld, float inv_tau,
                   const float* colbias, float* lse, float* entropy, dinox_stream_t stream) {
  DINOX_REQUIRE(x && lse && rows >= 0 && K > 0 && ld >= K, DINOX_E_BADARG, "rows_lse: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  if (rows == 0) return DINOX_OK;
  DINOX_REQUIRE(rows < (1ll << 31), DINOX_E_BADARG, "rows_lse: too many rows");
  const bool vec = vec_ok(x, dtype, K, ld) && fvec_ok(colbias);
  const float s2 = inv_tau * DINOX_LOG2E;
  const bool one_wave = rows * kRowThreads <= (int64_t)num_sms() * 2048;
  DISPATCH_T(dtype, T, {
    if (one_wave) {
      if (vec) rows_lse_kernel<T, true, kRowThreads><<<(unsigned)rows, kRowThreads, 0, stream>>>((const T*)x, K, ld, s2, colbias, lse, entropy);
      else rows_lse_kernel<T, false, kRowThreads><<<(unsigned)rows, kRowThreads, 0, stream>>>((const T*)x, K, ld, s2, colbias, lse, entropy);
    } else {
      if (vec) rows_lse_kernel<T, true, 256><<<(unsigned)rows, 256, 0, stream>>>((const T*)x, K, ld, s2, colbias, lse, entropy);
      else rows_lse_kernel<T, false, 256><<<(unsigned)rows, 256, 0, stream>>>((const T*)x, K, ld, s2, colbias, lse, entropy);
    }
  });
  return check_launch("rows_lse_kernel", stream);
}

int dinox_cols_lse(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld, float inv_tau,
                   const float* rowbias, float* out, dinox_stream_t stream) {
  DINOX_REQUIRE(x && out && rows > 0 && K > 0 && ld >= K, DINOX_E_BADARG, "cols_lse: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  const bool vec = vec_ok(x, dtype, K, ld);
  const unsigned grid = (unsigned)((K + 127) / 128);
  const float s2 = inv_tau * DINOX_LOG2E;
  DISPATCH_T(dtype, T, {
    if (vec) cols_lse_kernel<T, true><<<grid, 32 * kColGroups, 0, stream>>>((const T*)x, rows, K, ld, s2, rowbias, out);
    else cols_lse_kernel<T, false><<<grid, 32 * kColGroups, 0, stream>>>((const T*)x, rows, K, ld, s2, rowbias, out);
  });
  return check_launch("cols_lse_kernel", stream);
}

int dinox_lse_combine(const float* gathered, int world, int64_t K, float add, float* out,
                      dinox_stream_t stream) {
  DINOX_REQUIRE(gathered && out && world > 0 && K > 0, DINOX_E_BADARG, "lse_combine: bad arguments");
  lse_combine_kernel<<<(unsigned)((K + 255) / 256), 256, 0, stream>>>(gathered, world, K, add, out);
  return check_launch("lse_combine_kernel", stream);
}

int dinox_cols_sum(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld, float* out,
                   dinox_stream_t stream) {
  DINOX_REQUIRE(x && out && rows > 0 && K > 0 && ld >= K, DINOX_E_BADARG, "cols_sum: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  const bool vec = vec_ok(x, dtype, K, ld);
  const unsigned grid = (unsigned)((K + 127) / 128);
  DISPATCH_T(dtype, T, {
    if (vec) cols_sum_kernel<T, true><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, rows, K, ld, out, rows);
    else cols_sum_kernel<T, false><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, rows, K, ld, out, rows);
  });
  return check_launch("cols_sum_kernel", stream);
}

int64_t dinox_segment_cols_sum_chunks(int64_t rows0, int64_t rows1) {
  const int64_t m = rows0 > rows1 ? rows0 : rows1;
  return m > 0 ? (m + kSegChunk - 1) / kSegChunk : 1;
}

size_t dinox_segment_cols_sum_workspace_bytes(int64_t rows0, int64_t rows1, int64_t K) {
  // [tickets: 2 * ceil(K/128) unsigned, rounded up to 256 B | scratch (2, chunks, K) fp32]
  const size_t tk = ((size_t)(2 * ((K + 127) / 128)) * sizeof(unsigned) + 255) / 256 * 256;
  return tk + (size_t)2 * dinox_segment_cols_sum_chunks(rows0, rows1) * K * sizeof(float);
}

int dinox_segment_cols_sum(const void* x, int dtype, int64_t K, int64_t ld, int nseg, const int64_t* seg_begin,
                           const int64_t* seg_end, float* out, const float* counts_in, float* counts_out, void* workspace,
                           dinox_stream_t stream) {
  DINOX_REQUIRE(x && out && workspace && seg_begin && seg_end && (nseg == 1 || nseg == 2) && K > 0 && ld >= K &&
                    (!counts_out || counts_in),
                DINOX_E_BADARG, "segment_cols_sum: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  SumSegs sg{};
  for (int i = 0; i < nseg; ++i) {
    DINOX_REQUIRE(seg_begin[i] >= 0 && seg_end[i] >= seg_begin[i], DINOX_E_BADARG, "segment_cols_sum: bad segment");
    sg.begin[i] = seg_begin[i];
    sg.end[i] = seg_end[i];
  }
  const int64_t chunks = dinox_segment_cols_sum_chunks(sg.end[0] - sg.begin[0], nseg > 1 ? sg.end[1] - sg.begin[1] : 0);
  DINOX_REQUIRE(chunks <= 65535, DINOX_E_BADARG, "segment_cols_sum: too many rows");
  const size_t tk = ((size_t)(2 * ((K + 127) / 128)) * sizeof(unsigned) + 255) / 256 * 256;
  unsigned* tickets = reinterpret_cast<unsigned*>(workspace);
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + tk);
  const bool vec = vec_ok(x, dtype, K, ld);
  const dim3 grid((unsigned)((K + 127) / 128), (unsigned)chunks, (unsigned)nseg);
  DISPATCH_T(dtype, T, {
    if (vec) segment_cols_sum_kernel<T, true><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, K, ld, sg, out, counts_in, counts_out, scratch, tickets);
    else segment_cols_sum_kernel<T, false><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, K, ld, sg, out, counts_in, counts_out, scratch, tickets);
  });
  return check_launch("segment_cols_sum_kernel", stream);
}

int dinox_cols_sum_axpy(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld, float scale, const float* scale_dev,
                        float* out, int accumulate, dinox_stream_t stream) {
  DINOX_REQUIRE(x && out && rows > 0 && K > 0 && ld >= K, DINOX_E_BADARG, "cols_sum_axpy: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  const bool vec = vec_ok(x, dtype, K, ld);
  const unsigned grid = (unsigned)((K + 127) / 128);
  DISPATCH_T(dtype, T, {
    if (vec) cols_sum_kernel<T, true><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, rows, K, ld, out, rows, scale, scale_dev, accumulate);
    else cols_sum_kernel<T, false><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, rows, K, ld, out, rows, scale, scale_dev, accumulate);
  });
  return check_launch("cols_sum_kernel", stream);
}

int dinox_cols_sum_chunked(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld, int64_t chunk,
                           float* partial, dinox_stream_t stream) {
  DINOX_REQUIRE(x && partial && rows > 0 && K > 0 && ld >= K && chunk > 0, DINOX_E_BADARG, "cols_sum_chunked: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  const int64_t nchunks = (rows + chunk - 1) / chunk;
  DINOX_REQUIRE(nchunks <= 65535, DINOX_E_BADARG, "cols_sum_chunked: too many chunks");
  const bool vec = vec_ok(x, dtype, K, ld);
  const dim3 grid((unsigned)((K + 127) / 128), (unsigned)nchunks);
  DISPATCH_T(dtype, T, {
    if (vec) cols_sum_kernel<T, true><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, rows, K, ld, partial, chunk);
    else cols_sum_kernel<T, false><<<grid, 32 * kColSumGroups, 0, stream>>>((const T*)x, rows, K, ld, partial, chunk);
  });
  return check_launch("cols_sum_kernel<chunked>", stream);
}

int dinox_center_ema(float* center, const float* colsum, float inv_rows, float m, int64_t K,
                     dinox_stream_t stream) {
  DINOX_REQUIRE(center && colsum && K > 0, DINOX_E_BADARG, "center_ema: bad arguments");
  // (1 - m) evaluated in double then rounded, as Python does for `1 - self.center_momentum`
  const float om = (float)(1.0 - (double)m);
  center_ema_kernel<<<(unsigned)((K + 255) / 256), 256, 0, stream>>>(center, colsum, inv_rows, m, om, K);
  return check_launch("center_ema_kernel", stream);
}

int dinox_axpb(const float* a, float alpha, float beta, float* out, int64_t n, dinox_stream_t stream) {
  DINOX_REQUIRE(a && out && n > 0, DINOX_E_BADARG, "axpb: bad arguments");
  axpb_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, alpha, beta, out, n);
  return check_launch("axpb_kernel", stream);
}

int dinox_axpby(const float* x, float alpha, const float* alpha_dev, const float* y, float beta, float* out,
                int64_t n, dinox_stream_t stream) {
  DINOX_REQUIRE(x && out && n > 0, DINOX_E_BADARG, "axpby: bad arguments");
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, alpha, alpha_dev, y, beta, out, n);
  return check_launch("axpby_kernel", stream);
}

size_t dinox_ce_workspace_bytes(int64_t groups, int64_t K) {
  if (groups <= 0 || K <= 0) return 0;
  return (size_t)groups * 64 * kMaxGlobalViews * 2 * sizeof(float);
}

// measurement knob (read once): DINOX_CE_BWD_BATCH=4 forces the four-rows-at-a-time backward (64 registers, 4 CTAs
// per SM) for any view count.  Measured equal at the C2 shapes (73.7 us both; profiles/r02_hbm_kernels_tuning.txt), as
// was an L2 prefetch of the next column block in both kernels (no gain, removed).
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
static int ce_bwd_batch() { static const int v = env_int("DINOX_CE_BWD_BATCH", 0); return v; }

static int ce_args(CeArgs& a, const void* student, const void* teacher, int64_t groups, int V, int Vg,
                   int64_t K, int64_t ld_s, int64_t ld_t, float inv_tau_s, float inv_tau_t,
                   const float* colbias_t, const float* rowbias_t, const float* lse_s,
                   const float* group_w, float norm, int exclude_same, int ctas_per_sm) {
  DINOX_REQUIRE(student && teacher && rowbias_t && lse_s, DINOX_E_BADARG, "ce: null pointer");
  DINOX_REQUIRE(groups > 0 && K > 0 && V >= 1 && Vg >= 1 && Vg <= kMaxGlobalViews && Vg <= V,
                DINOX_E_BADARG, "ce: need groups>0, K>0, 1<=Vg<=%d, Vg<=V (got groups=%lld V=%d Vg=%d)",
                kMaxGlobalViews, (long long)groups, V, Vg);
  DINOX_REQUIRE(ld_s >= K && ld_t >= K, DINOX_E_BADARG, "ce: leading dimension < K");
  DINOX_REQUIRE(groups < (1ll << 31), DINOX_E_BADARG, "ce: too many groups");
  a.groups = groups; a.K = K; a.ld_s = ld_s; a.ld_t = ld_t; a.ld_g = ld_s;
  a.V = V; a.Vg = Vg; a.exclude_same = exclude_same ? 1 : 0;
  a.ksplit = pick_ksplit(groups, K, ctas_per_sm);
  a.s2 = inv_tau_s * DINOX_LOG2E; a.t2 = inv_tau_t * DINOX_LOG2E;
  a.colbias_t = colbias_t; a.rowbias_t = rowbias_t; a.lse_s = lse_s; a.group_w = group_w; a.norm = norm;
  return require_sm100();
}

int dinox_ce_fwd(const void* student, int s_dtype, const void* teacher, int t_dtype, int64_t groups,
                 int V, int Vg, int64_t K, int64_t ld_s, int64_t ld_t, float inv_tau_s, float inv_tau_t,
                 const float* colbias_t, const float* rowbias_t, const float* lse_s,
                 const float* group_w, float norm, int exclude_same, float* loss_out, void* workspace,
                 dinox_stream_t stream) {
  CeArgs a;
  int rc = ce_args(a, student, teacher, groups, V, Vg, K, ld_s, ld_t, inv_tau_s, inv_tau_t, colbias_t,
                   rowbias_t, lse_s, group_w, norm, exclude_same, kCeFwdCtasPerSm);
  if (rc) return rc;
  DINOX_REQUIRE(loss_out && workspace, DINOX_E_BADARG, "ce_fwd: null output/workspace");
  const bool vec = vec_ok(student, s_dtype, K, ld_s) && vec_ok(teacher, t_dtype, K, ld_t) && fvec_ok(colbias_t);
  dim3 grid((unsigned)groups, (unsigned)a.ksplit);
  float* partial = (float*)workspace;
  DISPATCH_T(s_dtype, TS, DISPATCH_T(t_dtype, TT, {
    if (vec) ce_fwd_kernel<TS, TT, true><<<grid, kCeThreads, 0, stream>>>((const TS*)student, (const TT*)teacher, a, partial);
    else ce_fwd_kernel<TS, TT, false><<<grid, kCeThreads, 0, stream>>>((const TS*)student, (const TT*)teacher, a, partial);
  }));
  rc = check_launch("ce_fwd_kernel", stream);
  if (rc) return rc;
  ce_finalize_kernel<<<1, 1024, 0, stream>>>(partial, a, inv_tau_s, loss_out);
  return check_launch("ce_finalize_kernel", stream);
}

}  // extern "C"

namespace dinox {
static int ce_stream() { static const int v = env_int("DINOX_CE_STREAM", 1); return v; }

// K-split of the persistent stream: work items = groups * splits are dealt round-robin to one CTA per SM; the smallest
// split count (each split keeps >= 2 blocks of kStreamCols columns) within 2 % of the best balance
static int pick_ksplit_stream(int64_t groups, int64_t K) {
  const int64_t nsm = num_sms();
  int64_t maxs = K / (2 * kStreamCols);
  if (maxs < 1) maxs = 1;
  if (maxs > 64) maxs = 64;
  int best = 1;
  double best_eff = 0.0;
  for (int64_t ks = 1; ks <= maxs; ++ks) {
    const int64_t items = groups * ks;
    const double eff = (double)items / (double)(((items + nsm - 1) / nsm) * nsm);
    if (eff > best_eff + 0.02) { best_eff = eff; best = (int)ks; }
  }
  return best;
}

template <typename TS, typename TT, int MAXV>
static int launch_onepass_stream(const void* student, const void* teacher, CeArgs& a, float* partial, cudaStream_t stream) {
  StreamArgs sa;
  sa.row_s = kStreamCols * (uint32_t)sizeof(TS);
  sa.row_t = kStreamCols * (uint32_t)sizeof(TT);
  sa.off_t = (uint32_t)a.V * sa.row_s;
  sa.off_cb = sa.off_t + (uint32_t)a.Vg * sa.row_t;
  sa.stage_bytes = sa.off_cb + (a.colbias_t ? kStreamCols * 4u : 0u);
  int stages = (int)(kStreamSmemBudget / sa.stage_bytes);
  if (stages > kStreamMaxStages) stages = kStreamMaxStages;
  if (stages < 2) return 1;   // not for this shape: the caller takes the register kernel
  sa.stages = stages;
  a.ksplit = pick_ksplit_stream(a.groups, a.K);
  sa.per = ((a.K + a.ksplit - 1) / a.ksplit + 7) & ~int64_t(7);
  sa.items = a.groups * a.ksplit;
  auto kern = ce_fwd_onepass_stream_kernel<TS, TT, MAXV>;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    DINOX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmemBudget));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t nsm = num_sms();
  const unsigned grid = (unsigned)(sa.items < nsm ? sa.items : nsm);
  kern<<<grid, kStreamThreads, (size_t)stages * sa.stage_bytes, stream>>>((const TS*)student, (const TT*)teacher, a, sa, partial);
  return 0;
}
}  // namespace dinox

extern "C" {
using namespace dinox;

int dinox_ce_onepass_max_views(void) { return kOnePassMaxViews; }

size_t dinox_ce_onepass_workspace_bytes(int64_t groups, int V, int Vg, int64_t K) {
  if (groups <= 0 || K <= 0 || V < 1 || Vg < 1) return 0;
  return ((size_t)groups * 64 * (size_t)(3 * Vg + 2 * V) + (size_t)groups * Vg) * sizeof(float);
}

int dinox_ce_fwd_onepass(const void* student, int s_dtype, const void* teacher, int t_dtype, int64_t groups,
                         int V, int Vg, int64_t K, int64_t ld_s, int64_t ld_t, float inv_tau_s, float inv_tau_t,
                         const float* colbias_t, const float* group_w, float norm, int exclude_same,
                         float* loss_out, float* lse_s_out, float* rowbias_t_out, void* workspace,
                         dinox_stream_t stream) {
  CeArgs a;
  DINOX_REQUIRE(lse_s_out && rowbias_t_out, DINOX_E_BADARG, "ce_fwd_onepass: null LSE outputs");
  int rc = ce_args(a, student, teacher, groups, V, Vg, K, ld_s, ld_t, inv_tau_s, inv_tau_t, colbias_t,
                   rowbias_t_out, lse_s_out, group_w, norm, exclude_same, V <= 4 ? kCeFwdCtasPerSm : kOnePassCtasPerSm);
  if (rc) return rc;
  DINOX_REQUIRE(inv_tau_s > 0.f, DINOX_E_BADARG, "ce_fwd_onepass: student temperature must be positive");
  DINOX_REQUIRE((double)V * (double)groups * (double)ld_s < 4294967296.0 && (double)Vg * (double)groups * (double)ld_t < 4294967296.0,
                DINOX_E_BADARG, "ce_fwd_onepass: logit matrices of 2^32 elements or more (use dinox_rows_lse + dinox_ce_fwd)");
  DINOX_REQUIRE(V <= kOnePassMaxViews, DINOX_E_BADARG, "ce_fwd_onepass: V=%d views > %d (use dinox_rows_lse + dinox_ce_fwd)",
                V, kOnePassMaxViews);
  DINOX_REQUIRE(loss_out && workspace, DINOX_E_BADARG, "ce_fwd_onepass: null output/workspace");
  const bool vec = vec_ok(student, s_dtype, K, ld_s) && vec_ok(teacher, t_dtype, K, ld_t) && fvec_ok(colbias_t);
  float* partial = (float*)workspace;
  // persistent shared-memory staged form: rows 16-byte aligned in both matrices (bulk copies), >= 2 stages fit
  const size_t es = elt_size(s_dtype), et = elt_size(t_dtype);
  const bool bulk_ok = vec && ce_stream() && aligned16(student) && aligned16(teacher) && (ld_s * es) % 16 == 0 &&
                       (ld_t * et) % 16 == 0 && (K * es) % 16 == 0 && (K * et) % 16 == 0;
  if (bulk_ok) {
    int r = 1;
#define DINOX_STREAM_LAUNCH(MAXV) \
    DISPATCH_T(s_dtype, TS, DISPATCH_T(t_dtype, TT, { r = launch_onepass_stream<TS, TT, MAXV>(student, teacher, a, partial, stream); }))
    if (V <= 4) { DINOX_STREAM_LAUNCH(4); } else { DINOX_STREAM_LAUNCH(kOnePassMaxViews); }
#undef DINOX_STREAM_LAUNCH
    if (r < 0) return r;
    if (r == 0) {
      rc = check_launch("ce_fwd_onepass_stream_kernel", stream);
      if (rc) return rc;
      ce_onepass_finalize_kernel<<<1, 1024, 0, stream>>>(partial, a, inv_tau_s, loss_out, lse_s_out, rowbias_t_out,
                                                         partial + (size_t)groups * 64 * (size_t)(3 * Vg + 2 * V));
      return check_launch("ce_onepass_finalize_kernel", stream);
    }
  }
  dim3 grid((unsigned)groups, (unsigned)a.ksplit);
#define DINOX_ONEPASS_LAUNCH(MAXV)                                                                                          \
  DISPATCH_T(s_dtype, TS, DISPATCH_T(t_dtype, TT, {                                                                         \
    if (vec) ce_fwd_onepass_kernel<TS, TT, true, MAXV><<<grid, kCeThreads, 0, stream>>>((const TS*)student, (const TT*)teacher, a, partial); \
    else ce_fwd_onepass_kernel<TS, TT, false, MAXV><<<grid, kCeThreads, 0, stream>>>((const TS*)student, (const TT*)teacher, a, partial);    \
  }))
  if (V <= 4) { DINOX_ONEPASS_LAUNCH(4); } else { DINOX_ONEPASS_LAUNCH(kOnePassMaxViews); }
#undef DINOX_ONEPASS_LAUNCH
  rc = check_launch("ce_fwd_onepass_kernel", stream);
  if (rc) return rc;
  ce_onepass_finalize_kernel<<<1, 1024, 0, stream>>>(partial, a, inv_tau_s, loss_out, lse_s_out, rowbias_t_out,
                                                       partial + (size_t)groups * 64 * (size_t)(3 * Vg + 2 * V));
  return check_launch("ce_onepass_finalize_kernel", stream);
}

int dinox_ce_bwd(const void* student, int s_dtype, const void* teacher, int t_dtype, int64_t groups,
                 int V, int Vg, int64_t K, int64_t ld_s, int64_t ld_t, float inv_tau_s, float inv_tau_t,
                 const float* colbias_t, const float* rowbias_t, const float* lse_s,
                 const float* group_w, float norm, int exclude_same, const float* upstream, void* grad,
                 int64_t ld_g, dinox_stream_t stream) {
  CeArgs a;
  int rc = ce_args(a, student, teacher, groups, V, Vg, K, ld_s, ld_t, inv_tau_s, inv_tau_t, colbias_t,
                   rowbias_t, lse_s, group_w, norm, exclude_same,
                   (V <= 4 || ce_bwd_batch() == 4) ? kCeBwdCtasPerSm : kOnePassCtasPerSm);
  if (rc) return rc;
  DINOX_REQUIRE(upstream && grad && ld_g >= K, DINOX_E_BADARG, "ce_bwd: null upstream/grad or ld_g < K");
  a.ld_g = ld_g;
  const bool vec = vec_ok(student, s_dtype, K, ld_s) && vec_ok(teacher, t_dtype, K, ld_t) &&
                   vec_ok(grad, s_dtype, K, ld_g) && fvec_ok(colbias_t);
  dim3 grid((unsigned)groups, (unsigned)a.ksplit);
#define DINOX_CE_BWD_LAUNCH(BATCH)                                                                                          \
  DISPATCH_T(s_dtype, TS, DISPATCH_T(t_dtype, TT, {                                                                         \
    if (vec) ce_bwd_kernel<TS, TT, true, BATCH><<<grid, kCeThreads, 0, stream>>>((const TS*)student, (const TT*)teacher, a, upstream, (TS*)grad); \
    else ce_bwd_kernel<TS, TT, false, BATCH><<<grid, kCeThreads, 0, stream>>>((const TS*)student, (const TT*)teacher, a, upstream, (TS*)grad);    \
  }))
  if (V <= 4 || ce_bwd_batch() == 4) { DINOX_CE_BWD_LAUNCH(4); } else { DINOX_CE_BWD_LAUNCH(kOnePassMaxViews); }
#undef DINOX_CE_BWD_LAUNCH
  return check_launch("ce_bwd_kernel", stream);
}

}  // extern "C"

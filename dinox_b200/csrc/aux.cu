// Small HBM-bound helper kernels around the tensor-core contractions of the loss head:
// row gather + cast to bf16 (CLS / masked-patch staging), GELU forward/backward of the projection
// head (zoo/arch.py:254, exact erf form), bf16 GEMV for the teacher-centre batch mean, indexed
// row gather-sum (dL/dh of entries -> rows), token L2-normalise forward/backward for Gram
// anchoring (scripts/phase5_big_run.py:726).  128-bit accesses, one warp per row.
#include "common.cuh"
#include <type_traits>

namespace dinox {

// ---------------------------------------------------------------------------------------------
// dst[r, :] = bf16( src[idx ? idx[r] : r, :] * scale )     src fp32 or bf16, row stride ld_src
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_cast_kernel(const T* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ idx,
                                   int64_t rows, int D, __nv_bfloat16* __restrict__ dst, int64_t ld_dst) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int64_t sr = idx ? idx[r] : r;
  __nv_bfloat16* d = dst + r * ld_dst;
  if (sr < 0) {  // padding entry: zero row
    for (int c = lane * 2; c < D; c += 64) *reinterpret_cast<__nv_bfloat162*>(d + c) = __floats2bfloat162_rn(0.f, 0.f);
    return;
  }
  const T* s = src + sr * ld_src;
  for (int c = lane * 2; c < D; c += 64) {
    float a = to_f32<T>(s[c]), b = to_f32<T>(s[c + 1]);
    *reinterpret_cast<__nv_bfloat162*>(d + c) = __floats2bfloat162_rn(a, b);
  }
}

// 16-byte vector flavour (D % 8 == 0, 16-byte aligned rows): each lane moves 8 elements per step,
// all loads of a row are issued before the first store
template <typename T>
__global__ void gather_cast_vec_kernel(const T* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ idx,
                                       int64_t rows, int D, __nv_bfloat16* __restrict__ dst, int64_t ld_dst) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int64_t sr = idx ? idx[r] : r;
  __nv_bfloat16* d = dst + r * ld_dst;
  constexpr int kMaxIter = 4;   // D <= 1024 per pass
  for (int c0 = 0; c0 < D; c0 += 256 * kMaxIter) {
    uint4 o[kMaxIter];
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int c = c0 + it * 256 + lane * 8;
      o[it] = make_uint4(0u, 0u, 0u, 0u);
      if (c < D && sr >= 0) {
        if (sizeof(T) == 4) {
          const float4 v0 = ldg_stream_f4(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + sr * ld_src + c));
          const float4 v1 = ldg_stream_f4(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + sr * ld_src + c + 4));
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v0.x, v0.y), h1 = __floats2bfloat162_rn(v0.z, v0.w);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v1.x, v1.y), h3 = __floats2bfloat162_rn(v1.z, v1.w);
          o[it] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                             *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        } else {
          const uint4 v = ldg_stream_u4(reinterpret_cast<const uint4*>(src + sr * ld_src + c));
          if (sizeof(T) == 2 && !std::is_same<T, __half>::value) {
            o[it] = v;   // bf16 -> bf16
          } else {
            const __half2* hp = reinterpret_cast<const __half2*>(&v);
            __nv_bfloat162 h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(hp[j]); h[j] = __floats2bfloat162_rn(f.x, f.y); }
            o[it] = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                               *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int c = c0 + it * 256 + lane * 8;
      if (c < D) *reinterpret_cast<uint4*>(d + c) = o[it];
    }
  }
}

// Two sources into one destination in ONE launch: rows [0, rows0) come from segment 0, rows [rows0, rows0 + rows1) from
// segment 1 (the staging of [CLS rows | masked-patch rows] of one branch of the fused head).  16-byte vector rows only.
struct GatherSeg {
  const void* src;
  int64_t ld_src;
  const int64_t* idx;
  int64_t rows;
};
template <typename T>
__global__ void gather_cast2_kernel(GatherSeg s0, GatherSeg s1, int D, __nv_bfloat16* __restrict__ dst, int64_t ld_dst) {
  int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= s0.rows + s1.rows) return;
  __nv_bfloat16* d = dst + r * ld_dst;
  const bool second = r >= s0.rows;
  const GatherSeg& sg = second ? s1 : s0;
  if (second) r -= s0.rows;
  const int64_t sr = sg.idx ? sg.idx[r] : r;
  const T* src = reinterpret_cast<const T*>(sg.src);
  for (int c = lane * 8; c < D; c += 256) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (sr >= 0) {
      if (sizeof(T) == 4) {
        const float* sp = reinterpret_cast<const float*>(src) + sr * sg.ld_src + c;
        const float4 v0 = ldg_stream_f4(reinterpret_cast<const float4*>(sp));
        const float4 v1 = ldg_stream_f4(reinterpret_cast<const float4*>(sp + 4));
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v0.x, v0.y), h1 = __floats2bfloat162_rn(v0.z, v0.w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v1.x, v1.y), h3 = __floats2bfloat162_rn(v1.z, v1.w);
        o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                       *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
      } else {
        const uint4 v = ldg_stream_u4(reinterpret_cast<const uint4*>(src + sr * sg.ld_src + c));
        if (!std::is_same<T, __half>::value) {
          o = v;
        } else {
          const __half2* hp = reinterpret_cast<const __half2*>(&v);
          __nv_bfloat162 h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(hp[j]); h[j] = __floats2bfloat162_rn(f.x, f.y); }
          o = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                         *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
        }
      }
    }
    *reinterpret_cast<uint4*>(d + c) = o;
  }
}

// out[r] = src[idx[r]] (fp32 scalars; idx < 0 -> fill)
__global__ void gather_f32_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t n,
                                  float fill, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = idx[i] >= 0 ? src[idx[i]] : fill;
}

// dst[idx[r], :] = src[r, :] (fp32 rows, D % 4 == 0): one warp per row.  The backward of the in-kernel
// token gather: masked positions are unique per crop, so rows never collide; idx < 0 is skipped.
__global__ void scatter_rows_kernel(const float* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ idx,
                                    int64_t rows, int D4, float* __restrict__ dst, int64_t ld_dst) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t t = idx[r];
  if (t < 0) return;
  const float4* s = reinterpret_cast<const float4*>(src + r * ld_src);
  float4* d = reinterpret_cast<float4*>(dst + t * ld_dst);
  for (int c = threadIdx.x & 31; c < D4; c += 32) d[c] = s[c];
}

// out[r, :] = float(src[idx[r], :]) (idx < 0 -> zero row): materialises index-named rows in fp32
template <typename T>
__global__ void gather_rows_f32_kernel(const T* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ idx,
                                       int64_t rows, int D, float* __restrict__ out, int64_t ld_out) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t t = idx ? idx[r] : r;
  float* o = out + r * ld_out;
  if (t < 0) { for (int c = threadIdx.x & 31; c < D; c += 32) o[c] = 0.f; return; }
  const T* x = src + t * ld_src;
  for (int c = threadIdx.x & 31; c < D; c += 32) o[c] = to_f32<T>(x[c]);
}

// dst[idx[r], :] += src[r, :] (unique idx, so no two rows collide: a plain read-modify-write is race free)
__global__ void scatter_add_rows_kernel(const float* __restrict__ src, int64_t ld_src, const int64_t* __restrict__ idx,
                                        int64_t rows, int D4, float* __restrict__ dst, int64_t ld_dst) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t t = idx[r];
  if (t < 0) return;
  const float4* s = reinterpret_cast<const float4*>(src + r * ld_src);
  float4* d = reinterpret_cast<float4*>(dst + t * ld_dst);
  for (int c = threadIdx.x & 31; c < D4; c += 32) {
    float4 a = d[c];
    const float4 b = s[c];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    d[c] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// GELU (erf):  h = bf16(gelu(a));   backward: da = bf16(dh * gelu'(a)), colsum(da) for db1
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__global__ void gelu_fwd_kernel(const float* __restrict__ a, int64_t n4, __nv_bfloat16* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(a)[i];
  __nv_bfloat162 lo = __floats2bfloat162_rn(gelu_f(v.x), gelu_f(v.y));
  __nv_bfloat162 hi = __floats2bfloat162_rn(gelu_f(v.z), gelu_f(v.w));
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  reinterpret_cast<uint2*>(h)[i] = o;
}

// da[r, c] = bf16(dh[r,c] * scale * gelu'(a[r,c])); each thread owns 4 columns of a kGeluSlab-row slab
// and emits one partial column sum per slab (fixed order => deterministic db1)
constexpr int kGeluSlab = 4;   // 16 rows per thread were a 16-deep chain of dependent loads (40 us at 8064 x 384)
__global__ void gelu_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ a, int64_t rows, int D,
                                const float* __restrict__ scale_dev, __nv_bfloat16* __restrict__ da,
                                float* __restrict__ colsum_partial /* (gridDim.y, D) */) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const float sc = scale_dev ? *scale_dev : 1.f;
  const int64_t r0 = (int64_t)blockIdx.y * kGeluSlab;
  const int64_t r1 = r0 + kGeluSlab < rows ? r0 + kGeluSlab : rows;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float4 g = *reinterpret_cast<const float4*>(dh + r * D + c);
    const float4 x = *reinterpret_cast<const float4*>(a + r * D + c);
    const float d0 = g.x * sc * gelu_grad_f(x.x), d1 = g.y * sc * gelu_grad_f(x.y);
    const float d2 = g.z * sc * gelu_grad_f(x.z), d3 = g.w * sc * gelu_grad_f(x.w);
    __nv_bfloat162 lo = __floats2bfloat162_rn(d0, d1), hi = __floats2bfloat162_rn(d2, d3);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(da + r * D + c) = o;
    s0 += d0; s1 += d1; s2 += d2; s3 += d3;
  }
  if (colsum_partial) *reinterpret_cast<float4*>(colsum_partial + (int64_t)blockIdx.y * D + c) = make_float4(s0, s1, s2, s3);
}

// The same with dh[r, :] = sum_{i in [ptr[r], ptr[r+1])} sum_{s < slabs} src[s*slab_stride + ent[i]*ld_src, :] formed on
// the fly (entries outer, split-K slabs inner: the order of gather_sum_kernel, so the two routes agree bit for bit):
// the per-row dL/dh of the fused head never exists in memory.
__global__ void gelu_bwd_gather_kernel(const float* __restrict__ src, int64_t ld_src, int slabs, int64_t slab_stride,
                                       const int64_t* __restrict__ ptr, const int64_t* __restrict__ ent,
                                       const float* __restrict__ a, int64_t rows, int D, const float* __restrict__ scale_dev,
                                       __nv_bfloat16* __restrict__ da, float* __restrict__ colsum_partial) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const float sc = scale_dev ? *scale_dev : 1.f;
  const int64_t r0 = (int64_t)blockIdx.y * kGeluSlab;
  const int64_t r1 = r0 + kGeluSlab < rows ? r0 + kGeluSlab : rows;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  // the CSR bounds of the slab's rows first (one round trip), then the rows one after the other
  int64_t bound[kGeluSlab + 1];
#pragma unroll
  for (int j = 0; j <= kGeluSlab; ++j) bound[j] = ptr[r0 + j < rows ? r0 + j : rows];
#pragma unroll
  for (int j = 0; j < kGeluSlab; ++j) {
    const int64_t r = r0 + j;
    if (r >= r1) break;
    const int64_t i0 = bound[j], i1 = bound[j + 1];
    const float4 x = *reinterpret_cast<const float4*>(a + r * D + c);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = i0; i < i1; ++i) {
      const float* base = src + ent[i] * ld_src + c;
      for (int sl = 0; sl < slabs; ++sl) {
        const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(base + sl * slab_stride));
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
    }
    const float d0 = g.x * sc * gelu_grad_f(x.x), d1 = g.y * sc * gelu_grad_f(x.y);
    const float d2 = g.z * sc * gelu_grad_f(x.z), d3 = g.w * sc * gelu_grad_f(x.w);
    __nv_bfloat162 lo = __floats2bfloat162_rn(d0, d1), hi = __floats2bfloat162_rn(d2, d3);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(da + r * D + c) = o;
    s0 += d0; s1 += d1; s2 += d2; s3 += d3;
  }
  if (colsum_partial) *reinterpret_cast<float4*>(colsum_partial + (int64_t)blockIdx.y * D + c) = make_float4(s0, s1, s2, s3);
}

// ---------------------------------------------------------------------------------------------
// out[k] = (sum_d W[k,d] * x[d]) * alpha + bias[k] * beta     W bf16 (K, D), x fp32 (D); warp per row
// ---------------------------------------------------------------------------------------------
__global__ void gemv_bf16_kernel(const __nv_bfloat16* __restrict__ W, int64_t ldw, const float* __restrict__ x,
                                 int64_t K, int D, float alpha, const float* __restrict__ bias, float beta,
                                 float* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const __nv_bfloat16* w = W + k * ldw;
  float acc = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 v = ldg_stream_u4(reinterpret_cast<const uint4*>(w + c));
    const float4 x0 = *reinterpret_cast<const float4*>(x + c), x1 = *reinterpret_cast<const float4*>(x + c + 4);
    acc = fmaf(__uint_as_float(v.x << 16), x0.x, acc); acc = fmaf(__uint_as_float(v.x & 0xffff0000u), x0.y, acc);
    acc = fmaf(__uint_as_float(v.y << 16), x0.z, acc); acc = fmaf(__uint_as_float(v.y & 0xffff0000u), x0.w, acc);
    acc = fmaf(__uint_as_float(v.z << 16), x1.x, acc); acc = fmaf(__uint_as_float(v.z & 0xffff0000u), x1.y, acc);
    acc = fmaf(__uint_as_float(v.w << 16), x1.z, acc); acc = fmaf(__uint_as_float(v.w & 0xffff0000u), x1.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[k] = acc * alpha + (bias ? bias[k] * beta : 0.f);
}

// nvec (<= 4) vectors against the same matrix in one pass over W: out[v][k] = alpha[v] * W[k,:].x[v] + beta*bias[k]
struct GemvAlphas { float a[4]; };
// optional in-place EMA of the result into a target vector: target[k] = m * target[k] + (1 - m) * value
struct GemvEma { float* target[4]; float m[4]; };
constexpr int kGemvRows = 4;
template <int NV>
__global__ void gemv_bf16_multi_kernel(const __nv_bfloat16* __restrict__ W, int64_t ldw, const float* __restrict__ X,
                                       int64_t K, int D, GemvAlphas alphas, const float* __restrict__ divisors,
                                       const float* __restrict__ bias, float beta, float* __restrict__ out, GemvEma ema) {
  const int lane = threadIdx.x & 31;
  // each warp owns kGemvRows consecutive rows and loads them together: one row of D = 384 is only 1.5 16-byte loads
  // per lane, which left the kernel latency bound at 1.7 TB/s
  const int64_t k0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kGemvRows;
  if (k0 >= K) return;
  float acc[kGemvRows][NV];
#pragma unroll
  for (int r = 0; r < kGemvRows; ++r)
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[r][v] = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    uint4 u[kGemvRows];
#pragma unroll
    for (int r = 0; r < kGemvRows; ++r)
      u[r] = (k0 + r < K) ? ldg_stream_u4(reinterpret_cast<const uint4*>(W + (k0 + r) * ldw + c)) : make_uint4(0, 0, 0, 0);
    float4 x0[NV], x1[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      x0[v] = *reinterpret_cast<const float4*>(X + (int64_t)v * D + c);
      x1[v] = *reinterpret_cast<const float4*>(X + (int64_t)v * D + c + 4);
    }
#pragma unroll
    for (int r = 0; r < kGemvRows; ++r) {
      const float wv[8] = {__uint_as_float(u[r].x << 16), __uint_as_float(u[r].x & 0xffff0000u), __uint_as_float(u[r].y << 16),
                           __uint_as_float(u[r].y & 0xffff0000u), __uint_as_float(u[r].z << 16), __uint_as_float(u[r].z & 0xffff0000u),
                           __uint_as_float(u[r].w << 16), __uint_as_float(u[r].w & 0xffff0000u)};
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float a = acc[r][v];
        a = fmaf(wv[0], x0[v].x, a); a = fmaf(wv[1], x0[v].y, a); a = fmaf(wv[2], x0[v].z, a); a = fmaf(wv[3], x0[v].w, a);
        a = fmaf(wv[4], x1[v].x, a); a = fmaf(wv[5], x1[v].y, a); a = fmaf(wv[6], x1[v].z, a); a = fmaf(wv[7], x1[v].w, a);
        acc[r][v] = a;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kGemvRows; ++r) {
    const int64_t k = k0 + r;
    if (k >= K) break;
    const float b = bias ? bias[k] * beta : 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float t = warp_sum(acc[r][v]);
      if (lane == 0) {
        const float val = t * (divisors ? alphas.a[v] / divisors[v] : alphas.a[v]) + b;
        if (out) out[(int64_t)v * K + k] = val;
        if (ema.target[v]) ema.target[v][k] = ema.target[v][k] * ema.m[v] + val * (1.f - ema.m[v]);   // :686-690 op order
      }
    }
  }
}

// dst[i] (+)= scale * sum_{s < slabs} src[s*slab_stride + i]   (fixed order; split-K partial slabs)
__global__ void sum_slabs_kernel(const float* __restrict__ src, int slabs, int64_t slab_stride, int64_t n4,
                                 const float* __restrict__ scale_dev, float scale, float* __restrict__ dst, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float sc = scale * (scale_dev ? *scale_dev : 1.f);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < slabs; ++s) {
    const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(src + s * slab_stride) + i);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  float4 o = make_float4(a.x * sc, a.y * sc, a.z * sc, a.w * sc);
  float4* d = reinterpret_cast<float4*>(dst) + i;
  if (accumulate) { const float4 old = *d; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
  *d = o;
}

// ---------------------------------------------------------------------------------------------
// dst[r, :] (+)= scale * sum_{i in [ptr[r], ptr[r+1])} sum_{s < slabs} src[s*slab_stride + ent[i]*ld, :]
// fp32, warp per row, fixed summation order (entries outer, split-K slabs inner)
// ---------------------------------------------------------------------------------------------
__global__ void gather_sum_kernel(const float* __restrict__ src, int64_t ld_src, int slabs, int64_t slab_stride,
                                  const int64_t* __restrict__ ptr,
                                  const int64_t* __restrict__ ent, int64_t rows, int D, const float* __restrict__ scale_dev,
                                  float scale, float* __restrict__ dst, int64_t ld_dst, int accumulate) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float sc = scale * (scale_dev ? *scale_dev : 1.f);
  const int64_t i0 = ptr[r], i1 = ptr[r + 1];
  for (int c = lane * 4; c < D; c += 128) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = i0; i < i1; ++i) {
      const float* base = src + ent[i] * ld_src + c;
      for (int sl = 0; sl < slabs; ++sl) {
        const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(base + sl * slab_stride));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    float4* d = reinterpret_cast<float4*>(dst + r * ld_dst + c);
    float4 o = make_float4(a.x * sc, a.y * sc, a.z * sc, a.w * sc);
    if (accumulate) { const float4 old = *d; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
    *d = o;
  }
}

// ---------------------------------------------------------------------------------------------
// Gram anchoring staging: token rows of feats (B, T, D) [fp32 | bf16] without CLS:
//   xn[b, t, :] = bf16( x / max(||x||, 1e-12) ),  inv_norm[b, t] = 1 / max(||x||, 1e-12)
// backward of the normalise:  dx = (dxn - xn_f * <xn_f, dxn>) * inv_norm  with xn_f = x*inv_norm,
// written to grad (B, T, D) fp32 at token t+1 (+)=, CLS row untouched (zeroed by the caller).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void normalize_tokens_kernel(const T* __restrict__ feats, int64_t stride_b, int64_t stride_t, int tokens,
                                        int skip, int D, int64_t n_rows, __nv_bfloat16* __restrict__ xn,
                                        float* __restrict__ inv_norm) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int64_t b = r / tokens, t = r % tokens;
  const T* x = feats + b * stride_b + (t + skip) * stride_t;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = to_f32<T>(x[c]); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  __nv_bfloat16* o = xn + r * D;
  for (int c = lane * 2; c < D; c += 64)
    *reinterpret_cast<__nv_bfloat162*>(o + c) = __floats2bfloat162_rn(to_f32<T>(x[c]) * inv, to_f32<T>(x[c + 1]) * inv);
  if (lane == 0) inv_norm[r] = inv;
}

template <typename T>
__global__ void normalize_bwd_kernel(const T* __restrict__ feats, int64_t stride_b, int64_t stride_t, int tokens,
                                     int skip, int D, int64_t n_rows, const float* __restrict__ dxn /* (n_rows, D) */,
                                     const float* __restrict__ inv_norm, const float* __restrict__ scale_dev, float scale,
                                     float* __restrict__ grad, int64_t gstride_b, int64_t gstride_t) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int64_t b = r / tokens, t = r % tokens;
  const T* x = feats + b * stride_b + (t + skip) * stride_t;
  const float* g = dxn + r * D;
  const float inv = inv_norm[r];
  const float sc = scale * (scale_dev ? *scale_dev : 1.f);
  float dot = 0.f;
  for (int c = lane; c < D; c += 32) dot = fmaf(to_f32<T>(x[c]) * inv, g[c], dot);
  dot = warp_sum(dot);
  float* o = grad + b * gstride_b + (t + skip) * gstride_t;
  for (int c = lane; c < D; c += 32) o[c] = (g[c] - to_f32<T>(x[c]) * inv * dot) * inv * sc;
  if (t < skip) {   // the skipped leading tokens (CLS) receive no gradient: row t of the image is written as zeros
    float* z = grad + b * gstride_b + t * gstride_t;
    for (int c = lane; c < D; c += 32) z[c] = 0.f;
  }
}

__global__ void fill_f32_kernel(float* __restrict__ p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// 16-byte stores, 4 per thread (the 100 MB gradient of W2 at the start of an accumulation window)
__global__ void fill_f32x4_kernel(float4* __restrict__ p, int64_t n4, float v) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x) * 4 + threadIdx.x;
  const float4 x = make_float4(v, v, v, v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = i0 + (int64_t)j * blockDim.x;
    if (i < n4) p[i] = x;
  }
}

}  // namespace dinox

// per-prototype offsets of the fused passes in log2 units, one launch: cs2 = b2s * as2; ct2 = (b2t - c) * at2;
// ct2p = (b2t - cp) * at2 (optional)
__global__ void head_offsets_kernel(const float* __restrict__ b2s, const float* __restrict__ b2t, const float* __restrict__ c,
                                    const float* __restrict__ cp, float as2, float at2, float* __restrict__ cs2,
                                    float* __restrict__ ct2, float* __restrict__ ct2p, int64_t K) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float bt = b2t[k];
  cs2[k] = b2s[k] * as2;
  ct2[k] = (bt - c[k]) * at2;
  if (ct2p) ct2p[k] = (bt - cp[k]) * at2;
}
// entry weights: out = base, except out[off + i] = mw[i] * scale for i < n (the iBOT entries' mask weights)
__global__ void entry_weights_kernel(const float* __restrict__ base, int64_t total, const float* __restrict__ mw, int64_t off,
                                     int64_t n, float scale, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  out[i] = (i >= off && i < off + n) ? mw[i - off] * scale : base[i];
}

struct ScalarTerms {
  const float* p[8];
  float w[8];
};
__global__ void scalar_combine_kernel(ScalarTerms t, int n, float scale, float* __restrict__ out, float* __restrict__ out_unscaled) {
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < n; ++i) a += t.w[i] * *t.p[i];   // fixed order
    out[0] = a * scale;
    if (out_unscaled) out_unscaled[0] = a;
  }
}
__global__ void scalar_fanout_kernel(const float* __restrict__ up, ScalarTerms t, int n, float scale, float* __restrict__ out) {
  if ((int)threadIdx.x < n) out[threadIdx.x] = *up * scale * t.w[threadIdx.x];
}

extern "C" {
using namespace dinox;

int dinox_gather_cast_bf16(const void* src, int src_dtype, int64_t ld_src, const int64_t* idx, int64_t rows,
                           int64_t D, void* dst, int64_t ld_dst, dinox_stream_t stream) {
  DINOX_REQUIRE(src && dst && rows >= 0 && D > 0 && D % 2 == 0 && ld_dst >= D, DINOX_E_BADARG, "gather_cast: bad arguments");
  DINOX_REQUIRE((reinterpret_cast<uintptr_t>(dst) % 4) == 0 && ld_dst % 2 == 0, DINOX_E_ALIGN, "gather_cast: dst misaligned");
  if (rows == 0) return DINOX_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const int es = src_dtype == DINOX_F32 ? 4 : 2;
  const bool vec = D % 8 == 0 && aligned16(src) && aligned16(dst) && (ld_src * es) % 16 == 0 && (ld_dst * 2) % 16 == 0;
#define DINOX_GC(T)                                                                                                \
  do {                                                                                                             \
    if (vec) gather_cast_vec_kernel<T><<<grid, 256, 0, stream>>>((const T*)src, ld_src, idx, rows, (int)D, (__nv_bfloat16*)dst, ld_dst); \
    else gather_cast_kernel<T><<<grid, 256, 0, stream>>>((const T*)src, ld_src, idx, rows, (int)D, (__nv_bfloat16*)dst, ld_dst);         \
  } while (0)
  if (src_dtype == DINOX_F32) DINOX_GC(float);
  else if (src_dtype == DINOX_BF16) DINOX_GC(__nv_bfloat16);
  else if (src_dtype == DINOX_F16) DINOX_GC(__half);
  else { set_error("gather_cast: unknown dtype %d", src_dtype); return DINOX_E_BADARG; }
#undef DINOX_GC
  return check_launch("gather_cast_kernel", stream);
}

int dinox_gather_cast_bf16_2(const void* src0, int64_t ld_src0, const int64_t* idx0, int64_t rows0, const void* src1,
                             int64_t ld_src1, const int64_t* idx1, int64_t rows1, int src_dtype, int64_t D, void* dst,
                             int64_t ld_dst, dinox_stream_t stream) {
  DINOX_REQUIRE(src0 && src1 && dst && rows0 >= 0 && rows1 >= 0 && D > 0 && ld_dst >= D, DINOX_E_BADARG,
                "gather_cast_2: bad arguments");
  const int es = src_dtype == DINOX_F32 ? 4 : 2;
  DINOX_REQUIRE(D % 8 == 0 && aligned16(src0) && aligned16(src1) && aligned16(dst) && (ld_src0 * es) % 16 == 0 &&
                    (ld_src1 * es) % 16 == 0 && (ld_dst * 2) % 16 == 0,
                DINOX_E_ALIGN, "gather_cast_2: rows must be 16-byte vectors (use dinox_gather_cast_bf16 per segment otherwise)");
  if (rows0 + rows1 == 0) return DINOX_OK;
  const unsigned grid = (unsigned)((rows0 + rows1 + 7) / 8);
  const GatherSeg s0{src0, ld_src0, idx0, rows0}, s1{src1, ld_src1, idx1, rows1};
  if (src_dtype == DINOX_F32) gather_cast2_kernel<float><<<grid, 256, 0, stream>>>(s0, s1, (int)D, (__nv_bfloat16*)dst, ld_dst);
  else if (src_dtype == DINOX_BF16) gather_cast2_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(s0, s1, (int)D, (__nv_bfloat16*)dst, ld_dst);
  else if (src_dtype == DINOX_F16) gather_cast2_kernel<__half><<<grid, 256, 0, stream>>>(s0, s1, (int)D, (__nv_bfloat16*)dst, ld_dst);
  else { set_error("gather_cast_2: unknown dtype %d", src_dtype); return DINOX_E_BADARG; }
  return check_launch("gather_cast2_kernel", stream);
}

int dinox_gather_f32(const float* src, const int64_t* idx, int64_t n, float fill, float* out, dinox_stream_t stream) {
  DINOX_REQUIRE(src && idx && out && n >= 0, DINOX_E_BADARG, "gather_f32: bad arguments");
  if (n == 0) return DINOX_OK;
  gather_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, idx, n, fill, out);
  return check_launch("gather_f32_kernel", stream);
}

int dinox_scatter_rows_f32(const float* src, int64_t ld_src, const int64_t* idx, int64_t rows, int64_t D, float* dst,
                           int64_t ld_dst, dinox_stream_t stream) {
  DINOX_REQUIRE(src && idx && dst && rows >= 0 && D > 0 && D % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0,
                DINOX_E_BADARG, "scatter_rows_f32: bad arguments (D, ld multiples of 4)");
  DINOX_REQUIRE(aligned16(src) && aligned16(dst), DINOX_E_ALIGN, "scatter_rows_f32: misaligned");
  if (rows == 0) return DINOX_OK;
  scatter_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(src, ld_src, idx, rows, (int)(D / 4), dst, ld_dst);
  return check_launch("scatter_rows_kernel", stream);
}

int dinox_gather_rows_f32(const void* src, int dtype, int64_t ld_src, const int64_t* idx, int64_t rows, int64_t D,
                          float* out, int64_t ld_out, dinox_stream_t stream) {
  DINOX_REQUIRE(src && out && rows >= 0 && D > 0, DINOX_E_BADARG, "gather_rows_f32: bad arguments");
  if (rows == 0) return DINOX_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (dtype == DINOX_F32) gather_rows_f32_kernel<float><<<grid, 256, 0, stream>>>((const float*)src, ld_src, idx, rows, (int)D, out, ld_out);
  else if (dtype == DINOX_BF16) gather_rows_f32_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)src, ld_src, idx, rows, (int)D, out, ld_out);
  else if (dtype == DINOX_F16) gather_rows_f32_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)src, ld_src, idx, rows, (int)D, out, ld_out);
  else { set_error("gather_rows_f32: dtype must be f32, bf16 or f16"); return DINOX_E_BADARG; }
  return check_launch("gather_rows_f32_kernel", stream);
}

int dinox_scatter_add_rows_f32(const float* src, int64_t ld_src, const int64_t* idx, int64_t rows, int64_t D, float* dst,
                               int64_t ld_dst, dinox_stream_t stream) {
  DINOX_REQUIRE(src && idx && dst && rows >= 0 && D > 0 && D % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0,
                DINOX_E_BADARG, "scatter_add_rows_f32: bad arguments (D, ld multiples of 4)");
  DINOX_REQUIRE(aligned16(src) && aligned16(dst), DINOX_E_ALIGN, "scatter_add_rows_f32: misaligned");
  if (rows == 0) return DINOX_OK;
  scatter_add_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(src, ld_src, idx, rows, (int)(D / 4), dst, ld_dst);
  return check_launch("scatter_add_rows_kernel", stream);
}

int dinox_gelu_fwd(const float* a, int64_t n, void* h_bf16, dinox_stream_t stream) {
  DINOX_REQUIRE(a && h_bf16 && n > 0 && n % 4 == 0, DINOX_E_BADARG, "gelu_fwd: n must be a positive multiple of 4");
  DINOX_REQUIRE(aligned16(a) && (reinterpret_cast<uintptr_t>(h_bf16) % 8) == 0, DINOX_E_ALIGN, "gelu_fwd: misaligned");
  gelu_fwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream>>>(a, n / 4, (__nv_bfloat16*)h_bf16);
  return check_launch("gelu_fwd_kernel", stream);
}

size_t dinox_gelu_bwd_workspace_bytes(int64_t rows, int64_t D) {
  return (size_t)((rows + kGeluSlab - 1) / kGeluSlab) * D * sizeof(float);
}

int dinox_gelu_bwd(const float* dh, const float* a, int64_t rows, int64_t D, const float* scale_dev, void* da_bf16,
                   float* colsum_partial, dinox_stream_t stream) {
  DINOX_REQUIRE(dh && a && da_bf16 && rows > 0 && D > 0 && D % 4 == 0, DINOX_E_BADARG, "gelu_bwd: bad arguments");
  dim3 grid((unsigned)((D / 4 + 127) / 128), (unsigned)((rows + kGeluSlab - 1) / kGeluSlab));
  gelu_bwd_kernel<<<grid, 128, 0, stream>>>(dh, a, rows, (int)D, scale_dev, (__nv_bfloat16*)da_bf16, colsum_partial);
  return check_launch("gelu_bwd_kernel", stream);
}

int dinox_gelu_bwd_gather(const float* src, int64_t ld_src, int slabs, int64_t slab_stride, const int64_t* ptr,
                          const int64_t* ent, const float* a, int64_t rows, int64_t D, const float* scale_dev, void* da_bf16,
                          float* colsum_partial, dinox_stream_t stream) {
  DINOX_REQUIRE(src && ptr && ent && a && da_bf16 && rows > 0 && D > 0 && D % 4 == 0 && slabs >= 1 && ld_src % 4 == 0 &&
                    slab_stride % 4 == 0 && aligned16(src),
                DINOX_E_BADARG, "gelu_bwd_gather: bad arguments");
  dim3 grid((unsigned)((D / 4 + 127) / 128), (unsigned)((rows + kGeluSlab - 1) / kGeluSlab));
  gelu_bwd_gather_kernel<<<grid, 128, 0, stream>>>(src, ld_src, slabs, slab_stride, ptr, ent, a, rows, (int)D, scale_dev,
                                                   (__nv_bfloat16*)da_bf16, colsum_partial);
  return check_launch("gelu_bwd_gather_kernel", stream);
}

int dinox_gemv_bf16(const void* W, int64_t ldw, const float* x, int64_t K, int64_t D, float alpha, const float* bias,
                    float beta, float* out, dinox_stream_t stream) {
  DINOX_REQUIRE(W && x && out && K > 0 && D > 0 && D % 8 == 0 && ldw % 8 == 0, DINOX_E_BADARG, "gemv_bf16: bad arguments (D, ldw multiples of 8)");
  DINOX_REQUIRE(aligned16(W) && aligned16(x), DINOX_E_ALIGN, "gemv_bf16: misaligned");
  gemv_bf16_kernel<<<(unsigned)((K + 7) / 8), 256, 0, stream>>>((const __nv_bfloat16*)W, ldw, x, K, (int)D, alpha, bias, beta, out);
  return check_launch("gemv_bf16_kernel", stream);
}

static int gemv_multi_impl(const void* W, int64_t ldw, const float* X, int nvec, int64_t K, int64_t D,
                           const float* alphas_host, const float* divisors_dev, const float* bias, float beta, float* out,
                           const GemvEma& ema, dinox_stream_t stream);

int dinox_gemv_bf16_multi(const void* W, int64_t ldw, const float* X, int nvec, int64_t K, int64_t D,
                          const float* alphas_host, const float* divisors_dev, const float* bias, float beta, float* out,
                          dinox_stream_t stream) {
  DINOX_REQUIRE(out, DINOX_E_BADARG, "gemv_bf16_multi: null output");
  GemvEma ema = {};
  return gemv_multi_impl(W, ldw, X, nvec, K, D, alphas_host, divisors_dev, bias, beta, out, ema, stream);
}

int dinox_gemv_bf16_multi_ema(const void* W, int64_t ldw, const float* X, int nvec, int64_t K, int64_t D,
                              const float* alphas_host, const float* divisors_dev, const float* bias, float beta,
                              float* const* targets, const float* momenta_host, dinox_stream_t stream) {
  DINOX_REQUIRE(targets && momenta_host && nvec >= 1 && nvec <= 4, DINOX_E_BADARG, "gemv_bf16_multi_ema: bad arguments");
  GemvEma ema = {};
  for (int v = 0; v < nvec; ++v) {
    DINOX_REQUIRE(targets[v], DINOX_E_BADARG, "gemv_bf16_multi_ema: null target");
    ema.target[v] = targets[v];
    ema.m[v] = momenta_host[v];
  }
  return gemv_multi_impl(W, ldw, X, nvec, K, D, alphas_host, divisors_dev, bias, beta, nullptr, ema, stream);
}

static int gemv_multi_impl(const void* W, int64_t ldw, const float* X, int nvec, int64_t K, int64_t D,
                           const float* alphas_host, const float* divisors_dev, const float* bias, float beta, float* out,
                           const GemvEma& ema, dinox_stream_t stream) {
  DINOX_REQUIRE(W && X && alphas_host && K > 0 && D > 0 && D % 8 == 0 && ldw % 8 == 0 && nvec >= 1 && nvec <= 4,
                DINOX_E_BADARG, "gemv_bf16_multi: bad arguments (D, ldw multiples of 8; 1..4 vectors)");
  DINOX_REQUIRE(aligned16(W) && aligned16(X), DINOX_E_ALIGN, "gemv_bf16_multi: misaligned");
  GemvAlphas al;
  for (int v = 0; v < 4; ++v) al.a[v] = v < nvec ? alphas_host[v] : 0.f;
  const unsigned grid = (unsigned)((K + 8 * kGemvRows - 1) / (8 * kGemvRows));
  const __nv_bfloat16* w = (const __nv_bfloat16*)W;
  switch (nvec) {
    case 1: gemv_bf16_multi_kernel<1><<<grid, 256, 0, stream>>>(w, ldw, X, K, (int)D, al, divisors_dev, bias, beta, out, ema); break;
    case 2: gemv_bf16_multi_kernel<2><<<grid, 256, 0, stream>>>(w, ldw, X, K, (int)D, al, divisors_dev, bias, beta, out, ema); break;
    case 3: gemv_bf16_multi_kernel<3><<<grid, 256, 0, stream>>>(w, ldw, X, K, (int)D, al, divisors_dev, bias, beta, out, ema); break;
    default: gemv_bf16_multi_kernel<4><<<grid, 256, 0, stream>>>(w, ldw, X, K, (int)D, al, divisors_dev, bias, beta, out, ema); break;
  }
  return check_launch("gemv_bf16_multi_kernel", stream);
}

int dinox_sum_slabs(const float* src, int slabs, int64_t slab_stride, int64_t n, const float* scale_dev, float scale,
                    float* dst, int accumulate, dinox_stream_t stream) {
  DINOX_REQUIRE(src && dst && slabs >= 1 && n > 0 && n % 4 == 0 && slab_stride % 4 == 0 && aligned16(src) && aligned16(dst),
                DINOX_E_BADARG, "sum_slabs: bad arguments (n, slab_stride multiples of 4; 16-byte aligned)");
  sum_slabs_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, stream>>>(src, slabs, slab_stride, n / 4, scale_dev, scale, dst, accumulate);
  return check_launch("sum_slabs_kernel", stream);
}

int dinox_gather_sum_rows(const float* src, int64_t ld_src, int slabs, int64_t slab_stride, const int64_t* ptr,
                          const int64_t* ent, int64_t rows, int64_t D, const float* scale_dev, float scale, float* dst,
                          int64_t ld_dst, int accumulate, dinox_stream_t stream) {
  DINOX_REQUIRE(src && ptr && ent && dst && rows >= 0 && D > 0 && D % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0 &&
                    slabs >= 1 && slab_stride % 4 == 0,
                DINOX_E_BADARG, "gather_sum_rows: bad arguments");
  if (rows == 0) return DINOX_OK;
  gather_sum_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(src, ld_src, slabs, slab_stride, ptr, ent, rows, (int)D, scale_dev, scale, dst, ld_dst, accumulate);
  return check_launch("gather_sum_kernel", stream);
}

int dinox_normalize_tokens(const void* feats, int dtype, int64_t batch, int64_t tokens_total, int64_t D,
                           int64_t stride_b, int64_t stride_t, int skip, void* xn_bf16, float* inv_norm,
                           dinox_stream_t stream) {
  DINOX_REQUIRE(feats && xn_bf16 && inv_norm && batch > 0 && tokens_total > skip && D > 0 && D % 2 == 0, DINOX_E_BADARG,
                "normalize_tokens: bad arguments");
  const int tokens = (int)(tokens_total - skip);
  const int64_t n_rows = batch * tokens;
  const unsigned grid = (unsigned)((n_rows + 7) / 8);
  if (dtype == DINOX_F32)
    normalize_tokens_kernel<float><<<grid, 256, 0, stream>>>((const float*)feats, stride_b, stride_t, tokens, skip, (int)D, n_rows, (__nv_bfloat16*)xn_bf16, inv_norm);
  else if (dtype == DINOX_BF16)
    normalize_tokens_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)feats, stride_b, stride_t, tokens, skip, (int)D, n_rows, (__nv_bfloat16*)xn_bf16, inv_norm);
  else { set_error("normalize_tokens: dtype must be f32 or bf16"); return DINOX_E_BADARG; }
  return check_launch("normalize_tokens_kernel", stream);
}

int dinox_normalize_tokens_bwd(const void* feats, int dtype, int64_t batch, int64_t tokens_total, int64_t D,
                               int64_t stride_b, int64_t stride_t, int skip, const float* dxn, const float* inv_norm,
                               const float* scale_dev, float scale, float* grad, int64_t gstride_b, int64_t gstride_t,
                               dinox_stream_t stream) {
  DINOX_REQUIRE(feats && dxn && inv_norm && grad && batch > 0 && tokens_total > skip && D > 0, DINOX_E_BADARG,
                "normalize_tokens_bwd: bad arguments");
  const int tokens = (int)(tokens_total - skip);
  const int64_t n_rows = batch * tokens;
  const unsigned grid = (unsigned)((n_rows + 7) / 8);
  if (dtype == DINOX_F32)
    normalize_bwd_kernel<float><<<grid, 256, 0, stream>>>((const float*)feats, stride_b, stride_t, tokens, skip, (int)D, n_rows, dxn, inv_norm, scale_dev, scale, grad, gstride_b, gstride_t);
  else if (dtype == DINOX_BF16)
    normalize_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)feats, stride_b, stride_t, tokens, skip, (int)D, n_rows, dxn, inv_norm, scale_dev, scale, grad, gstride_b, gstride_t);
  else { set_error("normalize_tokens_bwd: dtype must be f32 or bf16"); return DINOX_E_BADARG; }
  return check_launch("normalize_bwd_kernel", stream);
}

/* step glue (scripts/phase5_big_run.py:1749-1772): loss = scale * sum_i w_i * term_i and, for the backward,
 * the per-term upstream gradients g * scale * w_i - two launches instead of a dozen scalar framework ops */
int dinox_scalar_combine(const float* const* terms, const float* weights, int n, float scale, float* out,
                         float* out_unscaled, dinox_stream_t stream) {
  DINOX_REQUIRE(terms && weights && out && n >= 1 && n <= 8, DINOX_E_BADARG, "scalar_combine: 1..8 terms");
  ScalarTerms t;
  for (int i = 0; i < 8; ++i) { t.p[i] = i < n ? terms[i] : nullptr; t.w[i] = i < n ? weights[i] : 0.f; }
  for (int i = 0; i < n; ++i) DINOX_REQUIRE(t.p[i], DINOX_E_BADARG, "scalar_combine: null term");
  scalar_combine_kernel<<<1, 32, 0, stream>>>(t, n, scale, out, out_unscaled);
  return check_launch("scalar_combine_kernel", stream);
}

int dinox_scalar_fanout(const float* upstream, const float* weights, int n, float scale, float* out, dinox_stream_t stream) {
  DINOX_REQUIRE(upstream && weights && out && n >= 1 && n <= 8, DINOX_E_BADARG, "scalar_fanout: 1..8 terms");
  ScalarTerms t;
  for (int i = 0; i < 8; ++i) { t.p[i] = nullptr; t.w[i] = i < n ? weights[i] : 0.f; }
  scalar_fanout_kernel<<<1, 32, 0, stream>>>(upstream, t, n, scale, out);
  return check_launch("scalar_fanout_kernel", stream);
}

int dinox_head_offsets(const float* b2_student, const float* b2_teacher, const float* center, const float* center_patch,
                       float inv_tau_s, float inv_tau_t, float* cs2, float* ct2, float* ct2_patch, int64_t K,
                       dinox_stream_t stream) {
  DINOX_REQUIRE(b2_student && b2_teacher && center && cs2 && ct2 && K > 0 && (!ct2_patch || center_patch), DINOX_E_BADARG,
                "head_offsets: bad arguments");
  head_offsets_kernel<<<(unsigned)((K + 255) / 256), 256, 0, stream>>>(b2_student, b2_teacher, center, center_patch,
                                                                      inv_tau_s * DINOX_LOG2E, inv_tau_t * DINOX_LOG2E, cs2, ct2,
                                                                      ct2_patch, K);
  return check_launch("head_offsets_kernel", stream);
}

int dinox_entry_weights(const float* base, int64_t total, const float* mask_weights, int64_t offset, int64_t n, float scale,
                        float* out, dinox_stream_t stream) {
  DINOX_REQUIRE(base && out && total > 0 && n >= 0 && offset >= 0 && offset + n <= total && (n == 0 || mask_weights),
                DINOX_E_BADARG, "entry_weights: bad arguments");
  entry_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(base, total, mask_weights, offset, n, scale, out);
  return check_launch("entry_weights_kernel", stream);
}

int dinox_fill_f32(float* p, int64_t n, float v, dinox_stream_t stream) {
  DINOX_REQUIRE(p && n >= 0, DINOX_E_BADARG, "fill_f32: bad arguments");
  if (n == 0) return DINOX_OK;
  if (aligned16(p) && n % 4 == 0 && n >= 4096) {
    const int64_t n4 = n / 4;
    fill_f32x4_kernel<<<(unsigned)((n4 + 1023) / 1024), 256, 0, stream>>>(reinterpret_cast<float4*>(p), n4, v);
  } else {
    fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, n, v);
  }
  return check_launch("fill_f32_kernel", stream);
}

}  // extern "C"

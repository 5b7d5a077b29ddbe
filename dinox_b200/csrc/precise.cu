// fp32-faithful contraction mode (the reference WITHOUT --amp runs nn.Linear / torch.bmm in true fp32,
// scripts/phase5_big_run.py:1322 `--amp` is store_true): every dense contraction is evaluated as THREE bf16
// tensor-core GEMMs on hi/lo splits of its fp32 operands,
//     a = a_hi + a_lo,  a_hi = bf16(a),  a_lo = bf16(a - a_hi)         (a_lo carries the next 8 mantissa bits)
//     A.B^T ~= A_hi.B_hi^T + A_hi.B_lo^T + A_lo.B_hi^T                  (fp32 accumulation in TMEM / L2 reduce-add)
// dropping only the A_lo.B_lo term (2^-18 relative) - about 16 mantissa bits per product, i.e. fp32-grade results
// from the bf16 tensor pipe (tf32 would keep 10 bits).  This file holds the element-wise helpers of that mode:
// the operand split and fp32-output variants of GELU / token normalisation / Gram difference.
#include "common.cuh"

namespace dinox {

__device__ __forceinline__ float gelu_pf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_pf(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// hi/lo split of a (rows, cols) matrix with row pitch ld_src into two bf16 matrices with pitch ld_dst
template <typename T>
__global__ void split_bf16_kernel(const T* __restrict__ src, int64_t rows, int64_t cols, int64_t ld_src,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld_dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols, c = i % cols;
  const float v = to_f32<T>(src[r * ld_src + c]);
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[r * ld_dst + c] = h;
  lo[r * ld_dst + c] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void gelu_fwd_f32_kernel(const float* __restrict__ a, int64_t n, float* __restrict__ h) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) h[i] = gelu_pf(a[i]);
}

// da = dh * scale * gelu'(a) in fp32, plus per-slab column sums (fixed order) for db1
constexpr int kSlab = 16;
__global__ void gelu_bwd_f32_kernel(const float* __restrict__ dh, const float* __restrict__ a, int64_t rows, int D,
                                    const float* __restrict__ scale_dev, float* __restrict__ da,
                                    float* __restrict__ colsum_partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  const float sc = scale_dev ? *scale_dev : 1.f;
  const int64_t r0 = (int64_t)blockIdx.y * kSlab, r1 = r0 + kSlab < rows ? r0 + kSlab : rows;
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float d = dh[r * D + c] * sc * gelu_grad_pf(a[r * D + c]);
    da[r * D + c] = d;
    s += d;
  }
  if (colsum_partial) colsum_partial[(int64_t)blockIdx.y * D + c] = s;
}

template <typename T>
__global__ void normalize_tokens_f32_kernel(const T* __restrict__ feats, int64_t stride_b, int64_t stride_t, int tokens,
                                            int skip, int D, int64_t n_rows, float* __restrict__ xn, float* __restrict__ inv_norm) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int64_t b = r / tokens, t = r % tokens;
  const T* x = feats + b * stride_b + (t + skip) * stride_t;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = to_f32<T>(x[c]); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < D; c += 32) xn[r * D + c] = to_f32<T>(x[c]) * inv;
  if (lane == 0) inv_norm[r] = inv;
}

// delta = a - b (fp32, optional), per-block partial sums of delta^2 in fixed order
__global__ void __launch_bounds__(256) sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                     float* __restrict__ delta, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float d = a[i] - b[i];
    if (delta) delta[i] = d;
    s = fmaf(d, d, s);
  }
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) *out = s * scale;
}

}  // namespace dinox

extern "C" {
using namespace dinox;

int dinox_split_bf16(const void* src, int dtype, int64_t rows, int64_t cols, int64_t ld_src, void* hi, void* lo,
                     int64_t ld_dst, dinox_stream_t stream) {
  DINOX_REQUIRE(src && hi && lo && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, DINOX_E_BADARG, "split_bf16: bad arguments");
  const unsigned grid = (unsigned)((rows * cols + 255) / 256);
  if (dtype == DINOX_F32) split_bf16_kernel<float><<<grid, 256, 0, stream>>>((const float*)src, rows, cols, ld_src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_dst);
  else if (dtype == DINOX_BF16) split_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)src, rows, cols, ld_src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_dst);
  else if (dtype == DINOX_F16) split_bf16_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)src, rows, cols, ld_src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_dst);
  else { set_error("split_bf16: dtype must be f32, bf16 or f16"); return DINOX_E_BADARG; }
  return check_launch("split_bf16_kernel", stream);
}

int dinox_gelu_fwd_f32(const float* a, int64_t n, float* h, dinox_stream_t stream) {
  DINOX_REQUIRE(a && h && n > 0, DINOX_E_BADARG, "gelu_fwd_f32: bad arguments");
  gelu_fwd_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, n, h);
  return check_launch("gelu_fwd_f32_kernel", stream);
}

size_t dinox_gelu_bwd_f32_workspace_bytes(int64_t rows, int64_t D) { return (size_t)((rows + kSlab - 1) / kSlab) * D * sizeof(float); }

int dinox_gelu_bwd_f32(const float* dh, const float* a, int64_t rows, int64_t D, const float* scale_dev, float* da,
                       float* colsum_partial, dinox_stream_t stream) {
  DINOX_REQUIRE(dh && a && da && rows > 0 && D > 0, DINOX_E_BADARG, "gelu_bwd_f32: bad arguments");
  dim3 grid((unsigned)((D + 127) / 128), (unsigned)((rows + kSlab - 1) / kSlab));
  gelu_bwd_f32_kernel<<<grid, 128, 0, stream>>>(dh, a, rows, (int)D, scale_dev, da, colsum_partial);
  return check_launch("gelu_bwd_f32_kernel", stream);
}

int dinox_normalize_tokens_f32(const void* feats, int dtype, int64_t batch, int64_t tokens_total, int64_t D, int64_t stride_b,
                               int64_t stride_t, int skip, float* xn, float* inv_norm, dinox_stream_t stream) {
  DINOX_REQUIRE(feats && xn && inv_norm && batch > 0 && tokens_total > skip && D > 0, DINOX_E_BADARG, "normalize_tokens_f32: bad arguments");
  const int tokens = (int)(tokens_total - skip);
  const int64_t n_rows = batch * tokens;
  const unsigned grid = (unsigned)((n_rows + 7) / 8);
  if (dtype == DINOX_F32) normalize_tokens_f32_kernel<float><<<grid, 256, 0, stream>>>((const float*)feats, stride_b, stride_t, tokens, skip, (int)D, n_rows, xn, inv_norm);
  else if (dtype == DINOX_BF16) normalize_tokens_f32_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)feats, stride_b, stride_t, tokens, skip, (int)D, n_rows, xn, inv_norm);
  else { set_error("normalize_tokens_f32: dtype must be f32 or bf16"); return DINOX_E_BADARG; }
  return check_launch("normalize_tokens_f32_kernel", stream);
}

size_t dinox_sqdiff_workspace_bytes(void) { return 1024 * sizeof(float); }

int dinox_sqdiff_f32(const float* a, const float* b, int64_t n, float scale, float* delta, float* loss_out, void* workspace,
                     dinox_stream_t stream) {
  DINOX_REQUIRE(a && b && loss_out && workspace && n > 0, DINOX_E_BADARG, "sqdiff_f32: bad arguments");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  sqdiff_kernel<<<blocks, 256, 0, stream>>>(a, b, n, delta, (float*)workspace);
  int rc = check_launch("sqdiff_kernel", stream);
  if (rc) return rc;
  partial_sum_kernel<<<1, 256, 0, stream>>>((const float*)workspace, blocks, scale, loss_out);
  return check_launch("partial_sum_kernel", stream);
}

}  // extern "C"

// KoLeo (Kozachenko-Leonenko) entropy regulariser on head outputs - scripts/phase5_big_run.py:742-773:
//   x = z / max(||z||, 1e-12);  d_i = min_{j != i} ||x_i - x_j||;  loss = -mean_i log(d_i + eps)
// and its gradient with respect to z (autograd of normalize -> cdist -> min -> log).
//
// Nearest neighbours of unit vectors are the pairs with the largest cosine, and d^2 = 2 - 2 cos
// cancels catastrophically for close pairs.  So the cosine Gram matrix (a tcgen05 split-K GEMM of the
// bf16 rows, launched by the host through dinox_gemm_bf16_splitk) only RANKS candidates; the distances
// of the top candidates are then recomputed exactly as sum_k (x_ik - x_jk)^2 in fp32 from the
// original rows.  Everything is reduced in fixed order (deterministic).
#include "common.cuh"

namespace dinox {
constexpr int kKoleoCand = 8;      // candidates per row whose distance is recomputed exactly (bf16 ranking noise: the true nearest neighbour of a near-duplicate cluster must be among them)
constexpr int kKoleoMaxRows = 1024;

template <typename T>
__global__ void __launch_bounds__(256) koleo_rownorm_kernel(const T* __restrict__ z, int64_t ld, int64_t K,
                                                             float* __restrict__ inv_norm,
                                                             __nv_bfloat16* __restrict__ zb /* (R, ldb) or NULL */, int64_t ldb) {
  __shared__ float red[64];
  const int64_t r = blockIdx.x;
  const T* row = z + r * ld;
  float ss = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += 256) {
    const float v = to_f32<T>(row[k]);
    ss = fmaf(v, v, ss);
    if (zb) zb[r * ldb + k] = __float2bfloat16_rn(v);
  }
  ss = block_sum<256>(ss, red);
  if (threadIdx.x == 0) inv_norm[r] = 1.f / fmaxf(sqrtf(ss), 1e-12f);
}

// per row: the kKoleoCand largest cosines (excluding the row itself) from the approximate Gram matrix
__global__ void koleo_candidates_kernel(const float* __restrict__ gram, int64_t ldg, const float* __restrict__ inv_norm,
                                        int R, int* __restrict__ cand /* (R, kKoleoCand) */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  float best[kKoleoCand];
  int idx[kKoleoCand];
#pragma unroll
  for (int c = 0; c < kKoleoCand; ++c) { best[c] = -INFINITY; idx[c] = -1; }
  const float ni = inv_norm[i];
  for (int j = 0; j < R; ++j) {
    if (j == i) continue;
    float v = gram[(int64_t)i * ldg + j] * ni * inv_norm[j];
    int jj = j;
#pragma unroll
    for (int c = 0; c < kKoleoCand; ++c) {     // insertion into the sorted top list (ties: lower index first)
      if (v > best[c]) { const float tv = best[c]; const int ti = idx[c]; best[c] = v; idx[c] = jj; v = tv; jj = ti; }
    }
  }
#pragma unroll
  for (int c = 0; c < kKoleoCand; ++c) cand[i * kKoleoCand + c] = idx[c];
}

// exact squared distance of (row i, candidate c): sum_k (x_ik - x_jk)^2 with x = z * inv_norm
template <typename T>
__global__ void __launch_bounds__(256) koleo_exact_kernel(const T* __restrict__ z, int64_t ld, int64_t K,
                                                           const float* __restrict__ inv_norm, const int* __restrict__ cand,
                                                           float* __restrict__ d2 /* (R, kKoleoCand) */) {
  __shared__ float red[64];
  const int i = blockIdx.x, c = blockIdx.y;
  const int j = cand[i * kKoleoCand + c];
  if (j < 0) { if (threadIdx.x == 0) d2[i * kKoleoCand + c] = INFINITY; return; }
  const T* zi = z + (int64_t)i * ld;
  const T* zj = z + (int64_t)j * ld;
  const float ni = inv_norm[i], nj = inv_norm[j];
  float s = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += 256) {
    const float d = to_f32<T>(zi[k]) * ni - to_f32<T>(zj[k]) * nj;
    s = fmaf(d, d, s);
  }
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) d2[i * kKoleoCand + c] = s;
}

// nearest neighbour per row among the candidates, loss = -mean log(d + eps)   (single CTA, fixed order)
__global__ void __launch_bounds__(1024) koleo_finalize_kernel(const float* __restrict__ d2, const int* __restrict__ cand, int R,
                                                               float eps, int* __restrict__ nn, float* __restrict__ dist,
                                                               float* __restrict__ loss) {
  __shared__ float red[64];
  float acc = 0.f;
  for (int i = threadIdx.x; i < R; i += 1024) {
    float best = INFINITY;
    int bj = -1;
#pragma unroll
    for (int c = 0; c < kKoleoCand; ++c) {
      const float v = d2[i * kKoleoCand + c];
      const int j = cand[i * kKoleoCand + c];
      if (j >= 0 && (v < best || (v == best && j < bj))) { best = v; bj = j; }
    }
    const float d = bj >= 0 ? sqrtf(best) : 1e9f;    // a single row has no neighbour: the reference's 1e9 diagonal
    nn[i] = bj;
    dist[i] = d;
    acc += -logf(d + eps);
  }
  acc = block_sum<1024>(acc, red);
  if (threadIdx.x == 0) *loss = acc / (float)R;
}

// dz[r, k] = inv_r * (dx_r[k] - x_r[k] * <x_r, dx_r>),
//   dx_r = (c_r + sum_{i: nn_i = r} c_i) x_r - c_r x_{nn_r} - sum_{i: nn_i = r} c_i x_i,   c_i = -up / (R (d_i + eps) d_i)
template <typename T>
__global__ void __launch_bounds__(256) koleo_bwd_kernel(const T* __restrict__ z, int64_t ld, int64_t K,
                                                         const float* __restrict__ inv_norm, const int* __restrict__ nn,
                                                         const float* __restrict__ dist, int R, float eps,
                                                         const float* __restrict__ upstream, T* __restrict__ dz, int64_t ldd) {
  __shared__ int s_src[kKoleoMaxRows];
  __shared__ float s_coef[kKoleoMaxRows];
  __shared__ int s_n;
  __shared__ float s_self, s_dot;
  const int r = blockIdx.x;
  const float up = upstream ? *upstream : 1.f;
  if (threadIdx.x == 0) {
    // terms of dx_r as (source row, coefficient on x_source), in ascending source order (deterministic)
    int n = 0;
    float self = 0.f, dot = 0.f;
    auto coef = [&](int i) { const float d = dist[i]; return (d > 0.f && nn[i] >= 0) ? -up / ((float)R * (d + eps) * d) : 0.f; };
    const float cr = coef(r);
    if (nn[r] >= 0 && cr != 0.f) {
      self += cr;
      s_src[n] = nn[r]; s_coef[n] = -cr; ++n;
      dot += -cr * (1.f - 0.5f * dist[r] * dist[r]);
    }
    for (int i = 0; i < R; ++i) {
      if (i == r || nn[i] != r) continue;
      const float ci = coef(i);
      if (ci == 0.f) continue;
      self += ci;
      s_src[n] = i; s_coef[n] = -ci; ++n;
      dot += -ci * (1.f - 0.5f * dist[i] * dist[i]);
    }
    s_n = n; s_self = self; s_dot = dot + self;   // <x_r, x_r> = 1
  }
  __syncthreads();
  const int n = s_n;
  const float self = s_self, dot = s_dot, nr = inv_norm[r];
  const T* zr = z + (int64_t)r * ld;
  for (int64_t k = (int64_t)blockIdx.y * 256 + threadIdx.x; k < K; k += (int64_t)gridDim.y * 256) {
    const float xr = to_f32<T>(zr[k]) * nr;
    float dx = self * xr;
    for (int t = 0; t < n; ++t) dx = fmaf(s_coef[t], to_f32<T>(z[(int64_t)s_src[t] * ld + k]) * inv_norm[s_src[t]], dx);
    dz[(int64_t)r * ldd + k] = from_f32<T>((dx - xr * dot) * nr);
  }
}

}  // namespace dinox

extern "C" {
using namespace dinox;

int dinox_koleo_candidates() { return kKoleoCand; }

#define KOLEO_DISPATCH(dtype, T, body)                                                \
  do {                                                                                \
    if ((dtype) == DINOX_F32) { using T = float; body; }                              \
    else if ((dtype) == DINOX_BF16) { using T = __nv_bfloat16; body; }                \
    else if ((dtype) == DINOX_F16) { using T = __half; body; }                        \
    else { set_error("koleo: dtype must be f32, bf16 or f16"); return DINOX_E_BADARG; } \
  } while (0)

int dinox_koleo_rownorm(const void* z, int dtype, int64_t rows, int64_t K, int64_t ld, float* inv_norm, void* z_bf16,
                        int64_t ldb, dinox_stream_t stream) {
  DINOX_REQUIRE(z && inv_norm && rows > 0 && rows <= kKoleoMaxRows && K > 0 && ld >= K, DINOX_E_BADARG,
                "koleo_rownorm: bad arguments (1..%d rows)", kKoleoMaxRows);
  int rc = require_sm100();
  if (rc) return rc;
  KOLEO_DISPATCH(dtype, T, (koleo_rownorm_kernel<T><<<(unsigned)rows, 256, 0, stream>>>(
                               (const T*)z, ld, K, inv_norm, (__nv_bfloat16*)z_bf16, ldb)));
  return check_launch("koleo_rownorm_kernel", stream);
}

int dinox_koleo_fwd(const void* z, int dtype, int64_t rows, int64_t K, int64_t ld, const float* inv_norm,
                    const float* gram, int64_t ldg, float eps, int* cand, float* d2, int* nn, float* dist, float* loss,
                    dinox_stream_t stream) {
  DINOX_REQUIRE(z && inv_norm && gram && cand && d2 && nn && dist && loss && rows > 0 && rows <= kKoleoMaxRows && ldg >= rows,
                DINOX_E_BADARG, "koleo_fwd: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  koleo_candidates_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, stream>>>(gram, ldg, inv_norm, (int)rows, cand);
  rc = check_launch("koleo_candidates_kernel", stream);
  if (rc) return rc;
  const dim3 grid((unsigned)rows, kKoleoCand);
  KOLEO_DISPATCH(dtype, T, (koleo_exact_kernel<T><<<grid, 256, 0, stream>>>((const T*)z, ld, K, inv_norm, cand, d2)));
  rc = check_launch("koleo_exact_kernel", stream);
  if (rc) return rc;
  koleo_finalize_kernel<<<1, 1024, 0, stream>>>(d2, cand, (int)rows, eps, nn, dist, loss);
  return check_launch("koleo_finalize_kernel", stream);
}

int dinox_koleo_bwd(const void* z, int dtype, int64_t rows, int64_t K, int64_t ld, const float* inv_norm, const int* nn,
                    const float* dist, float eps, const float* upstream, void* dz, int64_t ldd, dinox_stream_t stream) {
  DINOX_REQUIRE(z && inv_norm && nn && dist && dz && rows > 0 && rows <= kKoleoMaxRows && K > 0 && ldd >= K, DINOX_E_BADARG,
                "koleo_bwd: bad arguments");
  int rc = require_sm100();
  if (rc) return rc;
  const unsigned gy = (unsigned)((K + 2047) / 2048 < 1 ? 1 : (K + 2047) / 2048 > 64 ? 64 : (K + 2047) / 2048);
  const dim3 grid((unsigned)rows, gy);
  KOLEO_DISPATCH(dtype, T, (koleo_bwd_kernel<T><<<grid, 256, 0, stream>>>((const T*)z, ld, K, inv_norm, nn, dist, (int)rows, eps,
                                                                         upstream, (T*)dz, ldd)));
  return check_launch("koleo_bwd_kernel", stream);
}

}  // extern "C"

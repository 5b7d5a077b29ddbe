// SURVEY 8f next #2 (first half): multi-tensor AdamW step + global gradient norm in ONE launch.
// Replaces, at the optimizer step of scripts/phase5_big_run.py:1781-1796,
//   * the per-parameter `p.grad.norm(2).item()` loop (one host sync per tensor, 161 of them), and
//   * torch.optim.AdamW's per-tensor kernels (:1621; decoupled weight decay, bias correction),
// with the arithmetic of torch.optim.AdamW (single-tensor path):
//   p *= 1 - lr*wd;  m += (g - m)(1 - b1);  v = v*b2 + (1 - b2) g*g;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
// HBM-bound: 16 B read + 12 B written per parameter.  The squared gradient norm is reduced per chunk
// and then over chunks in fixed order (deterministic); it stays on the device.
#include "common.cuh"
#include <vector>

namespace dinox {

struct AdamChunk {
  float* p;
  const float* g;
  float* m;
  float* v;
  int n;
  int pad;
};

constexpr int kAdamThreads = 256;
constexpr int kAdamVec = 4;                              // float4 per thread and array in flight
constexpr int kAdamChunk = kAdamThreads * 4 * kAdamVec;   // 4096 elements

struct AdamScalars {
  float decay;       // 1 - lr*wd
  float one_m_b1;    // 1 - beta1
  float b2, one_m_b2;
  float step_size;   // lr / bc1
  float inv_sqrt_bc2;
  float eps;
  float grad_scale;  // gradients are multiplied by this first (1/loss-scale, or 1)
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& s, float& ss) {
  g *= s.grad_scale;
  ss = fmaf(g, g, ss);
  p = __fmul_rn(p, s.decay);
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), s.one_m_b1));
  v = __fadd_rn(__fmul_rn(v, s.b2), __fmul_rn(__fmul_rn(s.one_m_b2, g), g));
  const float denom = __fadd_rn(__fmul_rn(sqrtf(v), s.inv_sqrt_bc2), s.eps);
  p = __fadd_rn(p, __fmul_rn(-s.step_size, __fdiv_rn(m, denom)));
}

__global__ void __launch_bounds__(kAdamThreads)
adamw_multi_kernel(const AdamChunk* __restrict__ chunks, int n_chunks, AdamScalars s, float* __restrict__ sumsq_partial) {
  __shared__ float red[64];
  float ss = 0.f;   // per-thread partial over every chunk this CTA walks (static assignment: deterministic)
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const AdamChunk ch = chunks[c];
    const int nvec = ch.n >> 2;
    float4 p[kAdamVec], g[kAdamVec], m[kAdamVec], v[kAdamVec];
#pragma unroll
    for (int j = 0; j < kAdamVec; ++j) {
      const int i = threadIdx.x + j * kAdamThreads;
      if (i < nvec) {
        g[j] = ldg_stream_f4(reinterpret_cast<const float4*>(ch.g) + i);
        p[j] = reinterpret_cast<const float4*>(ch.p)[i];
        m[j] = reinterpret_cast<const float4*>(ch.m)[i];
        v[j] = reinterpret_cast<const float4*>(ch.v)[i];
      }
    }
#pragma unroll
    for (int j = 0; j < kAdamVec; ++j) {
      const int i = threadIdx.x + j * kAdamThreads;
      if (i < nvec) {
        adam_one(p[j].x, g[j].x, m[j].x, v[j].x, s, ss); adam_one(p[j].y, g[j].y, m[j].y, v[j].y, s, ss);
        adam_one(p[j].z, g[j].z, m[j].z, v[j].z, s, ss); adam_one(p[j].w, g[j].w, m[j].w, v[j].w, s, ss);
        reinterpret_cast<float4*>(ch.p)[i] = p[j];
        reinterpret_cast<float4*>(ch.m)[i] = m[j];
        reinterpret_cast<float4*>(ch.v)[i] = v[j];
      }
    }
    const int tail0 = nvec << 2;
    if (threadIdx.x < ch.n - tail0) {
      const int i = tail0 + threadIdx.x;
      float pp = ch.p[i], mm = ch.m[i], vv = ch.v[i];
      adam_one(pp, ch.g[i], mm, vv, s, ss);
      ch.p[i] = pp; ch.m[i] = mm; ch.v[i] = vv;
    }
  }
  ss = block_sum<kAdamThreads>(ss, red);
  if (threadIdx.x == 0) sumsq_partial[blockIdx.x] = ss;
}

__global__ void __launch_bounds__(1024) adamw_norm_kernel(const float* __restrict__ partial, int n, float* __restrict__ norm_out) {
  __shared__ float red[64];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) a += partial[i];
  a = block_sum<1024>(a, red);
  if (threadIdx.x == 0) *norm_out = sqrtf(a);
}

}  // namespace dinox

struct dinox_adamw_plan {
  dinox::AdamChunk* d_chunks = nullptr;
  float* d_partial = nullptr;
  int n_chunks = 0;
  int64_t numel = 0;
};

extern "C" {
using namespace dinox;

int dinox_adamw_plan_create(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                            const int64_t* numel, int n_tensors, dinox_adamw_plan** out) {
  DINOX_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && out && n_tensors > 0, DINOX_E_BADARG,
                "adamw_plan_create: null argument or n_tensors <= 0");
  int rc = require_sm100();
  if (rc) return rc;
  std::vector<AdamChunk> chunks;
  int64_t total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    DINOX_REQUIRE(numel[i] >= 0, DINOX_E_BADARG, "adamw_plan_create: numel[%d] < 0", i);
    if (numel[i] == 0) continue;
    DINOX_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], DINOX_E_BADARG, "adamw_plan_create: tensor %d is null", i);
    DINOX_REQUIRE(aligned16(params[i]) && aligned16(grads[i]) && aligned16(exp_avg[i]) && aligned16(exp_avg_sq[i]), DINOX_E_ALIGN,
                  "adamw_plan_create: tensor %d is not 16-byte aligned", i);
    total += numel[i];
    for (int64_t off = 0; off < numel[i]; off += kAdamChunk) {
      AdamChunk c;
      c.p = static_cast<float*>(params[i]) + off;
      c.g = static_cast<const float*>(grads[i]) + off;
      c.m = static_cast<float*>(exp_avg[i]) + off;
      c.v = static_cast<float*>(exp_avg_sq[i]) + off;
      c.n = static_cast<int>(numel[i] - off < kAdamChunk ? numel[i] - off : kAdamChunk);
      c.pad = 0;
      chunks.push_back(c);
    }
  }
  dinox_adamw_plan* p = new dinox_adamw_plan();
  p->n_chunks = static_cast<int>(chunks.size());
  p->numel = total;
  if (p->n_chunks > 0) {
    cudaError_t e = cudaMalloc(&p->d_chunks, chunks.size() * sizeof(AdamChunk));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_partial, chunks.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_chunks, chunks.data(), chunks.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("adamw_plan_create: %s", cudaGetErrorString(e));
      if (p->d_chunks) cudaFree(p->d_chunks);
      if (p->d_partial) cudaFree(p->d_partial);
      delete p;
      return DINOX_E_CUDA;
    }
  }
  *out = p;
  return DINOX_OK;
}

int dinox_adamw_plan_destroy(dinox_adamw_plan* plan) {
  if (!plan) return DINOX_OK;
  if (plan->d_chunks) cudaFree(plan->d_chunks);
  if (plan->d_partial) cudaFree(plan->d_partial);
  delete plan;
  return DINOX_OK;
}

int dinox_adamw_step(const dinox_adamw_plan* plan, double lr, double beta1, double beta2, double eps, double weight_decay,
                     double bias_correction1, double bias_correction2, float grad_scale, float* grad_norm_out,
                     dinox_stream_t stream) {
  DINOX_REQUIRE(plan && grad_norm_out, DINOX_E_BADARG, "adamw_step: null plan or output");
  DINOX_REQUIRE(bias_correction1 > 0 && bias_correction2 > 0, DINOX_E_BADARG, "adamw_step: bias corrections must be > 0");
  if (plan->n_chunks == 0) return DINOX_OK;
  AdamScalars s;
  // the scalars are formed like torch.optim.adamw._single_tensor_adamw does (python floats = double)
  s.decay = (float)(1.0 - lr * weight_decay);
  s.one_m_b1 = (float)(1.0 - beta1);
  s.b2 = (float)beta2;
  s.one_m_b2 = (float)(1.0 - beta2);
  s.step_size = (float)(lr / bias_correction1);
  s.inv_sqrt_bc2 = (float)(1.0 / sqrt(bias_correction2));
  s.eps = (float)eps;
  s.grad_scale = grad_scale;
  int grid = num_sms() * 8;
  if (grid > plan->n_chunks) grid = plan->n_chunks;
  adamw_multi_kernel<<<grid, kAdamThreads, 0, stream>>>(plan->d_chunks, plan->n_chunks, s, plan->d_partial);
  int rc = check_launch("adamw_multi_kernel", stream);
  if (rc) return rc;
  adamw_norm_kernel<<<1, 1024, 0, stream>>>(plan->d_partial, grid, grad_norm_out);
  return check_launch("adamw_norm_kernel", stream);
}

}  // extern "C"

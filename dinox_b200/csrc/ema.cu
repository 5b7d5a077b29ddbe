// a8: multi-tensor EMA teacher update, one launch for all parameters.
// Replaces scripts/phase5_big_run.py:1798-1802 (2 launches per tensor, 322-610 launches).
// HBM-bound: 12 B/param (read p_s, read p_t, write p_t); 128-bit loads/stores; chunk table in
// device memory so that 100+ tiny tensors and one 25M-element tensor share one balanced grid.
#include "common.cuh"
#include <vector>

namespace dinox {

struct EmaChunk {
  const float* ps;
  float* pt;
  int n;        // elements in this chunk (multiple of 4 unless tail)
  int pad;
};

constexpr int kEmaThreads = 256;
constexpr int kEmaVecPerThread = 4;                                  // 4 x float4 in flight per thread
constexpr int kEmaChunk = kEmaThreads * kEmaVecPerThread * 4;       // 4096 elements = 16 KB

__global__ void __launch_bounds__(kEmaThreads)
ema_multi_kernel(const EmaChunk* __restrict__ chunks, int n_chunks, float m, float om) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const EmaChunk ch = chunks[c];
    const int nvec = ch.n >> 2;
    const float4* ps4 = reinterpret_cast<const float4*>(ch.ps);
    float4* pt4 = reinterpret_cast<float4*>(ch.pt);
    float4 s[kEmaVecPerThread], t[kEmaVecPerThread];
#pragma unroll
    for (int j = 0; j < kEmaVecPerThread; ++j) {
      const int i = threadIdx.x + j * kEmaThreads;
      if (i < nvec) {
        s[j] = ldg_stream_f4(ps4 + i);
        t[j] = pt4[i];
      }
    }
#pragma unroll
    for (int j = 0; j < kEmaVecPerThread; ++j) {
      const int i = threadIdx.x + j * kEmaThreads;
      if (i < nvec) {
        float4 r;
        r.x = __fmaf_rn(om, s[j].x, __fmul_rn(t[j].x, m));
        r.y = __fmaf_rn(om, s[j].y, __fmul_rn(t[j].y, m));
        r.z = __fmaf_rn(om, s[j].z, __fmul_rn(t[j].z, m));
        r.w = __fmaf_rn(om, s[j].w, __fmul_rn(t[j].w, m));
        pt4[i] = r;
      }
    }
    // scalar tail (numel not a multiple of 4)
    const int tail0 = nvec << 2;
    if (threadIdx.x < ch.n - tail0) {
      const int i = tail0 + threadIdx.x;
      ch.pt[i] = __fmaf_rn(om, ch.ps[i], __fmul_rn(ch.pt[i], m));
    }
  }
}

}  // namespace dinox

struct dinox_ema_plan {
  dinox::EmaChunk* d_chunks = nullptr;
  int n_chunks = 0;
  int64_t numel = 0;
  int device = 0;
};

extern "C" {

int dinox_ema_plan_create(const void* const* student, void* const* teacher, const int64_t* numel,
                          int n_tensors, dinox_ema_plan** out) {
  using namespace dinox;
  DINOX_REQUIRE(student && teacher && numel && out && n_tensors > 0, DINOX_E_BADARG,
                "ema_plan_create: null argument or n_tensors <= 0");
  int rc = require_sm100();
  if (rc != DINOX_OK) return rc;
  std::vector<EmaChunk> chunks;
  int64_t total = 0;
  for (int i = 0; i < n_tensors; ++i) {
    DINOX_REQUIRE(numel[i] >= 0, DINOX_E_BADARG, "ema_plan_create: numel[%d] < 0", i);
    if (numel[i] == 0) continue;
    DINOX_REQUIRE(student[i] && teacher[i], DINOX_E_BADARG, "ema_plan_create: tensor %d is null", i);
    DINOX_REQUIRE(aligned16(student[i]) && aligned16(teacher[i]), DINOX_E_ALIGN,
                  "ema_plan_create: tensor %d is not 16-byte aligned", i);
    DINOX_REQUIRE(student[i] != teacher[i], DINOX_E_BADARG, "ema_plan_create: tensor %d aliases itself", i);
    total += numel[i];
    for (int64_t off = 0; off < numel[i]; off += kEmaChunk) {
      EmaChunk c;
      c.ps = static_cast<const float*>(student[i]) + off;
      c.pt = static_cast<float*>(teacher[i]) + off;
      c.n = static_cast<int>(numel[i] - off < kEmaChunk ? numel[i] - off : kEmaChunk);
      c.pad = 0;
      chunks.push_back(c);
    }
  }
  dinox_ema_plan* p = new dinox_ema_plan();
  p->n_chunks = static_cast<int>(chunks.size());
  p->numel = total;
  cudaGetDevice(&p->device);
  if (p->n_chunks > 0) {
    cudaError_t e = cudaMalloc(&p->d_chunks, chunks.size() * sizeof(EmaChunk));
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_chunks, chunks.data(), chunks.size() * sizeof(EmaChunk), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("ema_plan_create: %s", cudaGetErrorString(e));
      if (p->d_chunks) cudaFree(p->d_chunks);
      delete p;
      return DINOX_E_CUDA;
    }
  }
  *out = p;
  return DINOX_OK;
}

int dinox_ema_plan_destroy(dinox_ema_plan* plan) {
  if (!plan) return DINOX_OK;
  if (plan->d_chunks) cudaFree(plan->d_chunks);
  delete plan;
  return DINOX_OK;
}

int64_t dinox_ema_plan_numel(const dinox_ema_plan* plan) { return plan ? plan->numel : -1; }

int dinox_ema_apply(const dinox_ema_plan* plan, float m, float one_minus_m, dinox_stream_t stream) {
  using namespace dinox;
  DINOX_REQUIRE(plan, DINOX_E_BADARG, "ema_apply: null plan");
  if (plan->n_chunks == 0) return DINOX_OK;
  // 148 SMs x 8 resident CTAs of 256 threads; grid-stride over chunks
  int grid = num_sms() * 8;
  if (grid > plan->n_chunks) grid = plan->n_chunks;
  ema_multi_kernel<<<grid, kEmaThreads, 0, stream>>>(plan->d_chunks, plan->n_chunks, m, one_minus_m);
  return check_launch("ema_multi_kernel", stream);
}

}  // extern "C"

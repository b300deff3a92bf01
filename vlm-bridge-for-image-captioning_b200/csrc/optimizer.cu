// Fused unscale + global gradient norm + clip + AdamW over the bridge's flat parameter / gradient
// arenas. Replaces, in the reference training step
// (core_training_loop.py:84-104): GradScaler.unscale_ (one read+write pass over the gradients), the
// per-parameter `grad.norm(2).item()` loop (52 host syncs), clip_grad_norm_ (a norm pass and a scaling
// pass) and torch.optim.AdamW.step (training_setup.py:248-254) -- about 4.4 GB of HBM traffic and 50+
// host round trips -- by two launches: a deterministic sum of squares, and one pass that reads
// p, g, m, v once, writes p, m, v once and also emits the bf16 operand copy of the updated weights
// (so the next forward needs no separate re-cast). HBM-bound: 28 B per element + 2 B per weight.
#include <math.h>

#include "common.cuh"
#include "launch.h"

namespace b200b {

constexpr int kNormThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (warp == 0) {
    s = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    s = warp_sum(s);
  }
  return s;  // valid in warp 0
}

// out[0] = sum g^2 (fixed summation order: per-thread strided, block tree, then the partials in block
// order by the last block to finish), out[1] = 1 if that sum is not finite. `state` = {ticket counter}.
// The first n4_bf16 groups of four are read from g16 (bf16: the averaged weight-gradient arena of the
// data-parallel exchange) when it is given, everything else from g (fp32, same element offsets).
__global__ void __launch_bounds__(kNormThreads) grad_sqnorm_kernel(const float4* __restrict__ g, long long n4,
                                                                   const uint2* __restrict__ g16, long long n4_bf16,
                                                                   float* __restrict__ partials,
                                                                   unsigned int* __restrict__ ticket,
                                                                   float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ bool last;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v;
    if (i < n4_bf16) {
      const uint2 h = __ldg(g16 + i);
      v = make_float4(bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y));
    } else {
      v = __ldg(g + i);
    }
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(partials + i);
  __syncthreads();
  t = block_sum(t, red);
  if (threadIdx.x == 0) {
    out[0] = t;
    out[1] = isfinite(t) ? 0.f : 1.f;
    *ticket = 0;  // ready for the next launch
  }
}

struct AdamParams {
  float4* p;
  const float4* g;
  const uint2* g16;        // bf16 gradients of the first n4_bf16 groups (data-parallel bf16 arena), or null
  float4* m;
  float4* v;
  uint2* w16;              // bf16 mirror of the first n4_bf16 float4 groups (may be null)
  long long n4, n4_bf16;
  const float* sqnorm;     // [2] from grad_sqnorm_kernel (may be null: no clipping, no finite check)
  const float* grad_scale; // device scalar the gradients were multiplied by (GradScaler), or null
  const float* found_inf;  // device scalar != 0 -> skip the step (GradScaler), or null
  float max_norm, lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2_sqrt;
  const float* step_dev;   // {steps applied so far, steps skipped}: when given, the bias corrections come from it
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamParams& a, float gmul) {
  g *= gmul;
  p *= (1.0f - a.lr * a.weight_decay);              // decoupled weight decay (AdamW)
  m = m + (1.0f - a.beta1) * (g - m);               // exp_avg.lerp_(grad, 1 - beta1)
  v = a.beta2 * v + (1.0f - a.beta2) * g * g;       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / a.bias_corr2_sqrt + a.eps;
  p = p - (a.lr / a.bias_corr1) * (m / denom);
  return p;
}

__device__ __forceinline__ bool adam_skipped(const AdamParams& a) {
  if (a.found_inf != nullptr && __ldg(a.found_inf) != 0.f) return true;
  return a.sqnorm != nullptr && __ldg(a.sqnorm + 1) != 0.f;   // non-finite gradients
}

// after the update: count the step as applied or skipped (torch's fused / capturable AdamW does not advance
// `step` on a step GradScaler skips; the next step's bias corrections depend on it)
__global__ void adamw_count_kernel(const AdamParams a, float* step_dev) {
  if (adam_skipped(a)) step_dev[1] += 1.0f;
  else step_dev[0] += 1.0f;
}

__global__ void __launch_bounds__(256) adamw_fused_kernel(AdamParams a) {
  if (a.step_dev != nullptr) {
    const double t = (double)__ldg(a.step_dev) + 1.0;
    a.bias_corr1 = (float)(1.0 - pow((double)a.beta1, t));
    a.bias_corr2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, t));
  }
  float gmul = 1.0f;
  if (a.grad_scale != nullptr) gmul = 1.0f / __ldg(a.grad_scale);
  if (a.found_inf != nullptr && __ldg(a.found_inf) != 0.f) return;
  if (a.sqnorm != nullptr) {
    if (__ldg(a.sqnorm + 1) != 0.f) return;         // non-finite gradients: skip the step
    if (a.max_norm > 0.f) {
      const float total = sqrtf(__ldg(a.sqnorm)) * gmul;
      const float coef = a.max_norm / (total + 1e-6f);   // torch.nn.utils.clip_grad_norm_
      gmul *= fminf(coef, 1.0f);
    }
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = a.p[i], m = a.m[i], v = a.v[i];
    float4 g;
    if (a.g16 != nullptr && i < a.n4_bf16) {
      const uint2 h = __ldg(a.g16 + i);
      g = make_float4(bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y));
    } else {
      g = __ldg(a.g + i);
    }
    adam_one(p.x, g.x, m.x, v.x, a, gmul);
    adam_one(p.y, g.y, m.y, v.y, a, gmul);
    adam_one(p.z, g.z, m.z, v.z, a, gmul);
    adam_one(p.w, g.w, m.w, v.w, a, gmul);
    a.p[i] = p;
    a.m[i] = m;
    a.v[i] = v;
    if (i < a.n4_bf16) a.w16[i] = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
  }
}

}  // namespace b200b

using namespace b200b;

extern "C" size_t b200b_grad_sqnorm_workspace_bytes(void) { return (size_t)(148 * 8 + 64) * sizeof(float); }

extern "C" int b200b_grad_sqnorm(const float* grad, int64_t n, const void* grad_bf16, int64_t n_bf16, void* workspace,
                                 size_t workspace_bytes, float* out2, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!grad || !workspace || !out2 || n <= 0 || (n % 4) || (reinterpret_cast<uintptr_t>(grad) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15)) {
    set_last_error("grad_sqnorm: need 16-byte aligned pointers and n a positive multiple of 4");
    return B200B_ERR_ARG;
  }
  if (grad_bf16 == nullptr) n_bf16 = 0;
  if (n_bf16 < 0 || n_bf16 > n || (n_bf16 % 4) || (reinterpret_cast<uintptr_t>(grad_bf16) & 7)) {
    set_last_error("grad_sqnorm: the bf16 part needs 0 <= n_bf16 <= n, n_bf16 %% 4 == 0 and an 8-byte aligned pointer");
    return B200B_ERR_ARG;
  }
  if (workspace_bytes < b200b_grad_sqnorm_workspace_bytes()) {
    set_last_error("grad_sqnorm: workspace too small");
    return B200B_ERR_WORKSPACE;
  }
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != B200B_OK) return rc;
  const long long n4 = n / 4;
  long long blocks = (n4 + kNormThreads - 1) / kNormThreads;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  // workspace: [0] ticket counter (zero before the first use: the caller zero-fills once), then partials
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  float* partials = reinterpret_cast<float*>(workspace) + 16;
  grad_sqnorm_kernel<<<(int)blocks, kNormThreads, 0, stream>>>(reinterpret_cast<const float4*>(grad), n4,
                                                              reinterpret_cast<const uint2*>(grad_bf16), n_bf16 / 4,
                                                              partials, ticket, out2);
  return check_launch("grad_sqnorm", stream);
}

extern "C" int b200b_adamw_fused(float* param, const float* grad, const void* grad_bf16, float* exp_avg, float* exp_avg_sq,
                                 void* weights_bf16,
                                 int64_t n, int64_t n_bf16, const float* sqnorm2, float max_grad_norm,
                                 const float* grad_scale, const float* found_inf, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int64_t step, float* step_dev, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || (n % 4) || n_bf16 < 0 || n_bf16 > n || (n_bf16 % 4) ||
      (step < 1 && step_dev == nullptr)) {
    set_last_error("adamw_fused: null pointer, n / n_bf16 not multiples of 4, or step < 1 without step_dev");
    return B200B_ERR_ARG;
  }
  const uintptr_t al = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
  if ((al & 15) || (reinterpret_cast<uintptr_t>(grad_bf16) & 7) ||
      (n_bf16 > 0 && (weights_bf16 == nullptr || (reinterpret_cast<uintptr_t>(weights_bf16) & 7)))) {
    set_last_error("adamw_fused: arenas must be 16-byte aligned (bf16 mirror 8-byte)");
    return B200B_ERR_ALIGN;
  }
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != B200B_OK) return rc;
  AdamParams a;
  a.p = reinterpret_cast<float4*>(param);
  a.g = reinterpret_cast<const float4*>(grad);
  a.g16 = reinterpret_cast<const uint2*>(grad_bf16);
  a.m = reinterpret_cast<float4*>(exp_avg);
  a.v = reinterpret_cast<float4*>(exp_avg_sq);
  a.w16 = reinterpret_cast<uint2*>(weights_bf16);
  a.n4 = n / 4;
  a.n4_bf16 = n_bf16 / 4;
  a.sqnorm = sqnorm2;
  a.grad_scale = grad_scale;
  a.found_inf = found_inf;
  a.max_norm = max_grad_norm;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.step_dev = step_dev;
  const double t = step < 1 ? 1.0 : (double)step;
  a.bias_corr1 = (float)(1.0 - pow((double)beta1, t));
  a.bias_corr2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
  long long blocks = (a.n4 + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  adamw_fused_kernel<<<(int)blocks, 256, 0, stream>>>(a);
  rc = check_launch("adamw_fused", stream);
  if (rc != B200B_OK || step_dev == nullptr) return rc;
  adamw_count_kernel<<<1, 1, 0, stream>>>(a, step_dev);
  return check_launch("adamw_count", stream);
}

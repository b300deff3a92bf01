// fp32 inference path of the bridge: every operand, product and sum in IEEE fp32 on the CUDA cores.
//
// Why it exists: the reference decodes with no autocast (full_model.py:221-261 runs the bridge in fp32),
// and a greedy loop amplifies any logit perturbation above the top-2 margin into a different caption.
// The bf16-operand tensor-core kernels reproduce the reference's *training* numerics (autocast); this
// path reproduces its *decode* numerics, so that greedy token ids equal the fp32 reference's on every
// step. It is the precision option of the decode path, not a fallback:
// it runs on the GPU only, through the same C ABI, and nothing selects it implicitly.
//
// Kernels: a register-tiled FFMA GEMM (y = x W^T + b with nn.Linear's [out, in] weight layout, optional
// exact-erf GELU or residual add in the epilogue; K is accumulated in index order, so results do not
// depend on the grid), a row LayerNorm, and an attention kernel that keeps one query row's scores in
// shared memory (softmax(q K^T / sqrt(d)) V, no mask). Reference call sites: bridge_module.py:98-100,118,
// 122-139,196-198,216,230-237,292-295,316-333.
#include <string.h>

#include "common.cuh"
#include "launch.h"

namespace b200b {
namespace {

constexpr int kTM = 64, kTN = 64, kTK = 16, kPitch = kTM + 4;
enum { kEpiBias = 0, kEpiBiasGelu = 1, kEpiBiasResid = 2 };

// C[M, N] = A[M, K] W[N, K]^T + bias (+ GELU | + resid). K % 16 == 0, lda / ldw % 4 == 0.
template <int EPI>
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, long long lda,
                                                       const float* __restrict__ W, long long ldw,
                                                       const float* __restrict__ bias, const float* __restrict__ resid,
                                                       long long ldr, float* __restrict__ C, long long ldc, int M, int N,
                                                       int K) {
  __shared__ __align__(16) float As[kTK][kPitch];
  __shared__ __align__(16) float Ws[kTK][kPitch];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const bool a_ok = m0 + lr < M, w_ok = n0 + lr < N;
  const float* ap = A + (long long)(a_ok ? m0 + lr : 0) * lda + lk;
  const float* wp = W + (long long)(w_ok ? n0 + lr : 0) * ldw + lk;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ra = a_ok ? *reinterpret_cast<const float4*>(ap) : zero4;
  float4 rw = w_ok ? *reinterpret_cast<const float4*>(wp) : zero4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kTK) {
    As[lk + 0][lr] = ra.x; As[lk + 1][lr] = ra.y; As[lk + 2][lr] = ra.z; As[lk + 3][lr] = ra.w;
    Ws[lk + 0][lr] = rw.x; Ws[lk + 1][lr] = rw.y; Ws[lk + 2][lr] = rw.z; Ws[lk + 3][lr] = rw.w;
    __syncthreads();
    if (k0 + kTK < K) {
      ra = a_ok ? *reinterpret_cast<const float4*>(ap + k0 + kTK) : zero4;
      rw = w_ok ? *reinterpret_cast<const float4*>(wp + k0 + kTK) : zero4;
    }
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= N) return;  // N % 4 == 0 is checked by the host
  float4 bz = bias ? *reinterpret_cast<const float4*>(bias + n) : zero4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) break;
    float4 v = make_float4(acc[i][0] + bz.x, acc[i][1] + bz.y, acc[i][2] + bz.z, acc[i][3] + bz.w);
    if (EPI == kEpiBiasGelu) {
      v.x = 0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f));
      v.y = 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f));
      v.z = 0.5f * v.z * (1.0f + erff(v.z * 0.70710678118654752f));
      v.w = 0.5f * v.w * (1.0f + erff(v.w * 0.70710678118654752f));
    }
    if (EPI == kEpiBiasResid) {
      const float4 r = *reinterpret_cast<const float4*>(resid + (long long)m * ldr + n);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    *reinterpret_cast<float4*>(C + (long long)m * ldc + n) = v;
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += scratch[i];
  return t;
}

__device__ __forceinline__ float block_max_256(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[w] = v;
  __syncthreads();
  float t = scratch[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) t = fmaxf(t, scratch[i]);
  return t;
}

// y = (x - mean) / sqrt(var + eps) * gamma + beta, biased variance from the centred values (two passes)
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            int dim, float eps) {
  __shared__ float scratch[8];
  const float* xr = x + (long long)blockIdx.x * dim;
  float* yr = y + (long long)blockIdx.x * dim;
  float s = 0.f;
  for (int c = threadIdx.x; c < dim; c += 256) s += xr[c];
  const float mean = block_sum_256(s, scratch) / (float)dim;
  float q = 0.f;
  for (int c = threadIdx.x; c < dim; c += 256) {
    const float d = xr[c] - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum_256(q, scratch) / (float)dim;
  const float rstd = 1.0f / sqrtf(var + eps);
  for (int c = threadIdx.x; c < dim; c += 256) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

// one CTA per (query row, head, image): scores of all keys in shared memory, exact softmax, then P V
__global__ void __launch_bounds__(256) attention_f32_kernel(const float* __restrict__ q, long long ldq,
                                                            const float* __restrict__ k, long long ldk,
                                                            const float* __restrict__ v, long long ldv,
                                                            float* __restrict__ o, long long ldo, int Lq, int Lk, int hd,
                                                            float scale) {
  extern __shared__ float sm[];
  float* qs = sm;          // [hd]
  float* ps = sm + hd;     // [Lk]
  __shared__ float scratch[8];
  const int qi = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* qr = q + ((long long)b * Lq + qi) * ldq + (long long)h * hd;
  for (int c = threadIdx.x; c < hd; c += 256) qs[c] = qr[c];
  __syncthreads();
  const float* kb = k + (long long)b * Lk * ldk + (long long)h * hd;
  for (int j = warp; j < Lk; j += 8) {
    const float* kr = kb + (long long)j * ldk;
    float s = 0.f;
    for (int c = lane; c < hd; c += 32) s = fmaf(qs[c], kr[c], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) ps[j] = s * scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < Lk; j += 256) mx = fmaxf(mx, ps[j]);
  mx = block_max_256(mx, scratch);
  float sum = 0.f;
  for (int j = threadIdx.x; j < Lk; j += 256) {
    const float e = expf(ps[j] - mx);
    ps[j] = e;
    sum += e;
  }
  sum = block_sum_256(sum, scratch);   // also orders the ps[] writes before the reads below
  const float inv = 1.0f / sum;
  const float* vb = v + (long long)b * Lk * ldv + (long long)h * hd;
  float* orow = o + ((long long)b * Lq + qi) * ldo + (long long)h * hd;
  for (int c = threadIdx.x; c < hd; c += 256) {
    float acc = 0.f;
    for (int j = 0; j < Lk; ++j) acc = fmaf(ps[j], vb[(long long)j * ldv + c], acc);
    orow[c] = acc * inv;
  }
}

int sgemm(int epi, const float* a, long long lda, const float* w, long long ldw, const float* bias, const float* resid,
          long long ldr, float* c, long long ldc, int m, int n, int k, cudaStream_t st) {
  if (m <= 0 || n <= 0 || k <= 0 || (k % kTK) || (n % 4) || (lda % 4) || (ldw % 4) || (ldc % 4) || (ldr % 4)) {
    set_last_error("sgemm_f32: needs k %% 16 == 0 and n / leading dimensions %% 4 == 0 (m=%d n=%d k=%d)", m, n, k);
    return B200B_ERR_SHAPE;
  }
  const dim3 grid((n + kTN - 1) / kTN, (m + kTM - 1) / kTM);
  switch (epi) {
    case kEpiBias: sgemm_nt_kernel<kEpiBias><<<grid, 256, 0, st>>>(a, lda, w, ldw, bias, resid, ldr, c, ldc, m, n, k); break;
    case kEpiBiasGelu: sgemm_nt_kernel<kEpiBiasGelu><<<grid, 256, 0, st>>>(a, lda, w, ldw, bias, resid, ldr, c, ldc, m, n, k); break;
    default: sgemm_nt_kernel<kEpiBiasResid><<<grid, 256, 0, st>>>(a, lda, w, ldw, bias, resid, ldr, c, ldc, m, n, k); break;
  }
  return check_launch("sgemm_f32", st);
}

int layernorm(const float* x, const float* g, const float* b, float* y, int rows, int dim, cudaStream_t st) {
  layernorm_f32_kernel<<<rows, 256, 0, st>>>(x, g, b, y, dim, 1e-5f);
  return check_launch("layernorm_f32", st);
}

int attention(const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv, float* o,
              long long ldo, int B, int H, int Lq, int Lk, int hd, cudaStream_t st) {
  const size_t smem = (size_t)(hd + Lk) * sizeof(float);
  if (smem > 96 * 1024) {
    set_last_error("attention_f32: %d keys x head dim %d do not fit the score buffer", Lk, hd);
    return B200B_ERR_SHAPE;
  }
  static bool attr_set = false;   // idempotent; a race only repeats the call
  if (!attr_set) {
    cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_set = true;
  }
  attention_f32_kernel<<<dim3(Lq, H, B), 256, smem, st>>>(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk, hd,
                                                          1.0f / sqrtf((float)hd));
  return check_launch("attention_f32", st);
}

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct WsF32 {
  float *xn, *qkv, *o, *x1, *x2, *h;
  size_t bytes;
};
WsF32 carve(size_t T, size_t D, size_t F, void* base) {
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t n) {
    float* r = reinterpret_cast<float*>(p + off);
    off += al256(n * sizeof(float));
    return r;
  };
  WsF32 w;
  w.xn = take(T * D); w.qkv = take(3 * T * D); w.o = take(T * D); w.x1 = take(T * D); w.x2 = take(T * D); w.h = take(T * F);
  w.bytes = off;
  return w;
}

int check_dims(const b200b_bridge_dims* d, const char* what) {
  if (d == nullptr || d->batch <= 0 || d->len_text <= 0 || d->len_vision <= 0 || d->dim <= 0 || d->dim_vision <= 0 ||
      d->dim_ffn <= 0 || d->heads_cross <= 0 || d->heads_self <= 0 || d->num_blocks <= 0) {
    set_last_error("%s: null dims or a non-positive dimension", what);
    return B200B_ERR_SHAPE;
  }
  if ((d->dim % 16) || (d->dim_vision % 16) || (d->dim_ffn % 16) || (d->dim % d->heads_cross) || (d->dim % d->heads_self)) {
    set_last_error("%s: dims must be multiples of 16 and dim divisible by the head counts", what);
    return B200B_ERR_SHAPE;
  }
  return B200B_OK;
}

}  // namespace
}  // namespace b200b

using namespace b200b;

#define B200B_TRY(expr)              \
  do {                               \
    int rc_ = (expr);                \
    if (rc_ != B200B_OK) return rc_; \
  } while (0)

extern "C" size_t b200b_bridge_f32_workspace_bytes(const b200b_bridge_dims* dims) {
  if (check_dims(dims, "f32_workspace_bytes") != B200B_OK) return 0;
  return carve((size_t)dims->batch * dims->len_text, dims->dim, dims->dim_ffn, nullptr).bytes;
}

extern "C" int b200b_bridge_kv_project_f32(const b200b_bridge_dims* dims, const float* vision, const float* wkv_all,
                                           const float* bkv_all, float* kv, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200B_TRY(check_dims(dims, "kv_project_f32"));
  if (!vision || !wkv_all || !bkv_all || !kv) {
    set_last_error("kv_project_f32: null argument");
    return B200B_ERR_ARG;
  }
  int sms = 0;
  B200B_TRY(device_sm_count(&sms));   // also rejects non-sm_100 devices
  const int n = 2 * dims->dim * dims->num_blocks;
  return sgemm(kEpiBias, vision, dims->dim_vision, wkv_all, dims->dim_vision, bkv_all, nullptr, 0, kv, n,
               dims->batch * dims->len_vision, n, dims->dim_vision, st);
}

extern "C" int b200b_bridge_block_forward_f32(const b200b_bridge_dims* dims, int i, const b200b_block_weights* w,
                                              const float* x_in, const float* kv, float* x_out, void* workspace,
                                              size_t workspace_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  B200B_TRY(check_dims(dims, "block_forward_f32"));
  if (!w || !x_in || !kv || !x_out || !workspace || i < 0 || i >= dims->num_blocks) {
    set_last_error("block_forward_f32: null argument or bad block index");
    return B200B_ERR_ARG;
  }
  int sms = 0;
  B200B_TRY(device_sm_count(&sms));
  const int B = dims->batch, L = dims->len_text, Nv = dims->len_vision, D = dims->dim, F = dims->dim_ffn;
  const int T = B * L;
  WsF32 ws = carve((size_t)T, D, F, workspace);
  if (workspace_bytes < ws.bytes) {
    set_last_error("block_forward_f32: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes);
    return B200B_ERR_WORKSPACE;
  }
  const bool do_cross = !(dims->flags & B200B_BRIDGE_PART_REST), do_rest = !(dims->flags & B200B_BRIDGE_PART_CROSS);
  if (!do_cross && !do_rest) {
    set_last_error("block_forward_f32: PART_CROSS and PART_REST are exclusive");
    return B200B_ERR_ARG;
  }
  const float* wq_c = reinterpret_cast<const float*>(w->wq_c);
  const float* wo_c = reinterpret_cast<const float*>(w->wo_c);
  const float* wqkv_s = reinterpret_cast<const float*>(w->wqkv_s);
  const float* wo_s = reinterpret_cast<const float*>(w->wo_s);
  const float* w1 = reinterpret_cast<const float*>(w->w1);
  const float* w2 = reinterpret_cast<const float*>(w->w2);
  const long long ldkv = 2LL * D * dims->num_blocks;
  const float* kblk = kv + (size_t)2 * D * i;
  float* x1w = do_rest ? ws.x1 : x_out;
  const float* x1 = do_cross ? x1w : x_in;
  if (do_cross) {   // bridge_module.py:316-323
    B200B_TRY(layernorm(x_in, w->ln_c_g, w->ln_c_b, ws.xn, T, D, st));
    B200B_TRY(sgemm(kEpiBias, ws.xn, D, wq_c, D, w->bq_c, nullptr, 0, ws.qkv, D, T, D, D, st));
    B200B_TRY(attention(ws.qkv, D, kblk, ldkv, kblk + D, ldkv, ws.o, D, B, dims->heads_cross, L, Nv,
                        D / dims->heads_cross, st));
    B200B_TRY(sgemm(kEpiBiasResid, ws.o, D, wo_c, D, w->bo_c, x_in, D, x1w, D, T, D, D, st));
  }
  if (!do_rest) return B200B_OK;
  // self-attention, non-causal and unmasked (:326-328)
  B200B_TRY(layernorm(x1, w->ln_s_g, w->ln_s_b, ws.xn, T, D, st));
  B200B_TRY(sgemm(kEpiBias, ws.xn, D, wqkv_s, D, w->bqkv_s, nullptr, 0, ws.qkv, 3 * D, T, 3 * D, D, st));
  B200B_TRY(attention(ws.qkv, 3 * D, ws.qkv + D, 3 * D, ws.qkv + 2 * D, 3 * D, ws.o, D, B, dims->heads_self, L, L,
                      D / dims->heads_self, st));
  B200B_TRY(sgemm(kEpiBiasResid, ws.o, D, wo_s, D, w->bo_s, x1, D, ws.x2, D, T, D, D, st));
  // FFN (:331-333)
  B200B_TRY(layernorm(ws.x2, w->ln_f_g, w->ln_f_b, ws.xn, T, D, st));
  B200B_TRY(sgemm(kEpiBiasGelu, ws.xn, D, w1, D, w->b1, nullptr, 0, ws.h, F, T, F, D, st));
  B200B_TRY(sgemm(kEpiBiasResid, ws.h, F, w2, F, w->b2, ws.x2, D, x_out, D, T, D, F, st));
  return B200B_OK;
}

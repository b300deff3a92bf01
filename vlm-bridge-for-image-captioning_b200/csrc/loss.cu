// Fused cross-entropy over the language model's vocabulary logits.
// Replaces, in the reference training step (core_training_loop.py:51-55,68-69):
//   labels = input_ids shifted left by one, last position = -100
//   loss = nn.CrossEntropyLoss(ignore_index=-100)(logits.view(-1, V), labels.view(-1))
// which in PyTorch up-casts the [B*L, V] logits to fp32 under autocast, materialises their
// log-softmax (1.05 GB at B=8, L=128, V=256000) and walks the matrix five to six times over forward
// and backward. Here: one read of the logits in the forward (online max / sum per row -> log-sum-exp
// and the row's loss), one read + one write in the backward (softmax recomputed from the saved
// log-sum-exp, one-hot subtracted, scaled by grad_loss / count, stored in the logits' dtype). Only
// [rows] floats are saved between the two. HBM-bound: forward V*e bytes per row, backward 2*V*e
// (e = 4 for fp32 logits, 2 for bf16).
#include <math.h>

#include "common.cuh"
#include "launch.h"

namespace b200b {

constexpr int kCeFwdThreads = 512;
constexpr int kCeBwdThreads = 256;
constexpr int kCeUnroll = 4;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// label of `row`: labels[row], or -- with shift_len = L > 0 (`labels` is then the [B, L] input_ids) --
// input_ids[row + 1] for every position but the last of a sequence, which is ignored
// (core_training_loop.py:52-54)
__device__ __forceinline__ long long ce_label(const long long* __restrict__ labels, long long row, long long shift_len,
                                              long long ignore_index) {
  if (shift_len > 0) {
    if (row % shift_len == shift_len - 1) return ignore_index;
    return __ldg(labels + row + 1);
  }
  return __ldg(labels + row);
}

// VEC consecutive logits starting at element `off` of `base`, as floats
template <bool BF16, int VEC>
__device__ __forceinline__ void ce_load(const void* __restrict__ base, long long off, float (&x)[VEC]) {
  if constexpr (VEC == 8) {
    if constexpr (BF16) {
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
      x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
      x[4] = bf16_lo(v.z); x[5] = bf16_hi(v.z); x[6] = bf16_lo(v.w); x[7] = bf16_hi(v.w);
    } else {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
      const float4 a = p[0], b = p[1];
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
  } else {
    if constexpr (BF16)
      x[0] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off]);
    else
      x[0] = reinterpret_cast<const float*>(base)[off];
  }
}

template <bool BF16, int VEC>
__device__ __forceinline__ void ce_store(void* __restrict__ base, long long off, const float (&g)[VEC]) {
  if constexpr (VEC == 8) {
    if constexpr (BF16) {
      const uint4 v = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
      __stcs(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off), v);
    } else {
      float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
      __stcs(p, make_float4(g[0], g[1], g[2], g[3]));
      __stcs(p + 1, make_float4(g[4], g[5], g[6], g[7]));
    }
  } else {
    if constexpr (BF16)
      reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(g[0]);
    else
      reinterpret_cast<float*>(base)[off] = g[0];
  }
}

// One CTA per row: lse[row] = log sum_j exp(logits[row, j]) (natural log),
// loss_rows[row] = lse - logits[row, label] (0 for an ignored row, NaN for a label outside [0, vocab)).
template <bool BF16, int VEC>
__global__ void __launch_bounds__(kCeFwdThreads) ce_fwd_kernel(const void* __restrict__ logits, long long ld,
                                                                const long long* __restrict__ labels, long long vocab,
                                                                long long ignore_index, long long shift_len,
                                                                float* __restrict__ lse, float* __restrict__ loss_rows) {
  __shared__ float red_m[kCeFwdThreads / 32], red_s[kCeFwdThreads / 32];
  const long long row = blockIdx.x;
  const long long base = row * ld;
  const long long nvec = vocab / VEC;
  // running maximum m and sum s of 2^(x*log2e - m) over this thread's elements
  float m = -INFINITY, s = 0.f;
  for (long long v0 = threadIdx.x; v0 < nvec; v0 += (long long)kCeFwdThreads * kCeUnroll) {
    float x[kCeUnroll][VEC];
#pragma unroll
    for (int u = 0; u < kCeUnroll; ++u) {
      const long long v = v0 + (long long)u * kCeFwdThreads;
      if (v < nvec) {
        ce_load<BF16, VEC>(logits, base + v * VEC, x[u]);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) x[u][e] = -INFINITY;
      }
    }
    float mv = -INFINITY;
#pragma unroll
    for (int u = 0; u < kCeUnroll; ++u)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        x[u][e] *= kLog2e;
        mv = fmaxf(mv, x[u][e]);
      }
    if (mv > m) {
      s *= ex2_approx(m - mv);   // m = -inf on the first group: 2^-inf = 0 and s is still 0
      m = mv;
    }
    if (m > -INFINITY) {
#pragma unroll
      for (int u = 0; u < kCeUnroll; ++u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) s += ex2_approx(x[u][e] - m);
    }
  }
  // warp, then block combination of (m, s)
  float wm = warp_max(m);
  s = (m > -INFINITY) ? s * ex2_approx(m - wm) : 0.f;
  s = warp_sum(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red_m[warp] = wm;
    red_s[warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float m2 = lane < kCeFwdThreads / 32 ? red_m[lane] : -INFINITY;
    float s2 = lane < kCeFwdThreads / 32 ? red_s[lane] : 0.f;
    const float bm = warp_max(m2);
    s2 = (m2 > -INFINITY) ? s2 * ex2_approx(m2 - bm) : 0.f;
    s2 = warp_sum(s2);
    if (lane == 0) {
      const float l = (bm + log2f(s2)) * kLn2;
      lse[row] = l;
      const long long lab = ce_label(labels, row, shift_len, ignore_index);
      float loss = 0.f;
      if (lab != ignore_index) {
        if (lab >= 0 && lab < vocab) {
          float t[1];
          ce_load<BF16, 1>(logits, base + lab, t);
          loss = l - t[0];
        } else {
          loss = NAN;   // PyTorch device-asserts on such a label; here the loss says so
        }
      }
      loss_rows[row] = loss;
    }
  }
}

// out2[0] = mean of loss_rows over the rows whose label is not ignore_index, out2[1] = their count.
// Fixed summation order (per-thread strided, then a block tree): deterministic.
__global__ void __launch_bounds__(1024) ce_finalize_kernel(const float* __restrict__ loss_rows,
                                                           const long long* __restrict__ labels, long long rows,
                                                           long long ignore_index, long long shift_len,
                                                           float* __restrict__ out2) {
  __shared__ float red_a[32], red_c[32];
  float a = 0.f, c = 0.f;
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    if (ce_label(labels, r, shift_len, ignore_index) != ignore_index) {
      a += loss_rows[r];
      c += 1.f;
    }
  }
  a = warp_sum(a);
  c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red_a[warp] = a;
    red_c[warp] = c;
  }
  __syncthreads();
  if (warp == 0) {
    a = lane < (int)(blockDim.x >> 5) ? red_a[lane] : 0.f;
    c = lane < (int)(blockDim.x >> 5) ? red_c[lane] : 0.f;
    a = warp_sum(a);
    c = warp_sum(c);
    if (lane == 0) {
      out2[0] = a / c;   // no valid row: 0 / 0 = NaN, as nn.CrossEntropyLoss(reduction="mean")
      out2[1] = c;
    }
  }
}

// dlogits[row, j] = (softmax(logits[row])[j] - [j == label]) * grad_loss / count; 0 for ignored rows.
// Each CTA handles a contiguous run of vectors of one row.
template <bool BF16, int VEC>
__global__ void __launch_bounds__(kCeBwdThreads) ce_bwd_kernel(const void* __restrict__ logits, long long ld,
                                                                const long long* __restrict__ labels, long long vocab,
                                                                long long ignore_index, long long shift_len,
                                                                const float* __restrict__ lse,
                                                                const float* __restrict__ out2,
                                                                const float* __restrict__ grad_loss,
                                                                void* __restrict__ dlogits, long long ldd,
                                                                int blocks_per_row) {
  const long long row = blockIdx.x / blocks_per_row;
  const int chunk = blockIdx.x - (int)(row * blocks_per_row);
  const long long nvec = vocab / VEC;
  const long long lab = ce_label(labels, row, shift_len, ignore_index);
  const bool valid = lab != ignore_index;
  const float count = __ldg(out2 + 1);
  const float coef = valid ? __ldg(grad_loss) / count : 0.f;
  const float l2 = __ldg(lse + row) * kLog2e;
  const long long v_begin = (long long)chunk * kCeBwdThreads * kCeUnroll;
#pragma unroll
  for (int u = 0; u < kCeUnroll; ++u) {
    const long long v = v_begin + (long long)u * kCeBwdThreads + threadIdx.x;
    if (v >= nvec) break;
    float g[VEC];
    if (valid) {
      float x[VEC];
      ce_load<BF16, VEC>(logits, row * ld + v * VEC, x);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float p = ex2_approx(fmaf(x[e], kLog2e, -l2));
        g[e] = (p - ((v * VEC + e) == lab ? 1.f : 0.f)) * coef;
      }
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) g[e] = 0.f;
    }
    ce_store<BF16, VEC>(dlogits, row * ldd + v * VEC, g);
  }
}

static int ce_check(const char* what, const void* logits, int dtype, long long ld, const void* labels, long long rows,
                    long long vocab, long long shift_len) {
  if (!logits || !labels || rows <= 0 || vocab <= 0 || ld < vocab || (dtype != 0 && dtype != 1)) {
    set_last_error("%s: null pointer, rows / vocab <= 0, ld < vocab or dtype not 0 (f32) / 1 (bf16)", what);
    return B200B_ERR_ARG;
  }
  if (shift_len < 0 || (shift_len > 0 && rows % shift_len != 0)) {
    set_last_error("%s: shift_len must be 0 or divide rows", what);
    return B200B_ERR_SHAPE;
  }
  if (rows > 0x7fffffffLL / 1024) {
    set_last_error("%s: too many rows", what);
    return B200B_ERR_SHAPE;
  }
  return B200B_OK;
}

// 16-byte vectors are usable when every row starts 16-byte aligned and holds a whole number of them
static bool ce_vectorizable(const void* p, int dtype, long long ld, long long vocab) {
  const size_t esz = dtype == 1 ? 2 : 4;
  return vocab % 8 == 0 && (ld * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

}  // namespace b200b

using namespace b200b;

extern "C" int b200b_cross_entropy_fwd(const void* logits, int dtype, int64_t ld, const int64_t* labels, int64_t rows,
                                       int64_t vocab, int64_t ignore_index, int64_t shift_len, float* lse,
                                       float* loss_rows, float* out2, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = ce_check("cross_entropy_fwd", logits, dtype, ld, labels, rows, vocab, shift_len);
  if (rc != B200B_OK) return rc;
  if (!lse || !loss_rows || !out2) {
    set_last_error("cross_entropy_fwd: null output");
    return B200B_ERR_ARG;
  }
  int sms = 0;
  rc = device_sm_count(&sms);
  if (rc != B200B_OK) return rc;
  const long long* lab = reinterpret_cast<const long long*>(labels);
  const bool vec = ce_vectorizable(logits, dtype, ld, vocab);
  const dim3 grid((unsigned)rows), block(kCeFwdThreads);
  if (dtype == 1) {
    if (vec) ce_fwd_kernel<true, 8><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, loss_rows);
    else ce_fwd_kernel<true, 1><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, loss_rows);
  } else {
    if (vec) ce_fwd_kernel<false, 8><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, loss_rows);
    else ce_fwd_kernel<false, 1><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, loss_rows);
  }
  rc = check_launch("cross_entropy_fwd", stream);
  if (rc != B200B_OK) return rc;
  ce_finalize_kernel<<<1, 1024, 0, stream>>>(loss_rows, lab, rows, ignore_index, shift_len, out2);
  return check_launch("cross_entropy_finalize", stream);
}

extern "C" int b200b_cross_entropy_bwd(const void* logits, int dtype, int64_t ld, const int64_t* labels, int64_t rows,
                                       int64_t vocab, int64_t ignore_index, int64_t shift_len, const float* lse,
                                       const float* out2, const float* grad_loss, void* dlogits, int64_t ldd,
                                       void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = ce_check("cross_entropy_bwd", logits, dtype, ld, labels, rows, vocab, shift_len);
  if (rc != B200B_OK) return rc;
  if (!lse || !out2 || !grad_loss || !dlogits || ldd < vocab) {
    set_last_error("cross_entropy_bwd: null pointer or ldd < vocab");
    return B200B_ERR_ARG;
  }
  int sms = 0;
  rc = device_sm_count(&sms);
  if (rc != B200B_OK) return rc;
  const long long* lab = reinterpret_cast<const long long*>(labels);
  const bool vec = ce_vectorizable(logits, dtype, ld, vocab) && ce_vectorizable(dlogits, dtype, ldd, vocab);
  const long long nvec = vec ? vocab / 8 : vocab;
  const long long per_block = (long long)kCeBwdThreads * kCeUnroll;
  const long long bpr = (nvec + per_block - 1) / per_block;
  if (bpr * rows > 0x7fffffffLL) {
    set_last_error("cross_entropy_bwd: grid too large");
    return B200B_ERR_SHAPE;
  }
  const dim3 grid((unsigned)(bpr * rows)), block(kCeBwdThreads);
  if (dtype == 1) {
    if (vec) ce_bwd_kernel<true, 8><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, out2, grad_loss, dlogits, ldd, (int)bpr);
    else ce_bwd_kernel<true, 1><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, out2, grad_loss, dlogits, ldd, (int)bpr);
  } else {
    if (vec) ce_bwd_kernel<false, 8><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, out2, grad_loss, dlogits, ldd, (int)bpr);
    else ce_bwd_kernel<false, 1><<<grid, block, 0, stream>>>(logits, ld, lab, vocab, ignore_index, shift_len, lse, out2, grad_loss, dlogits, ldd, (int)bpr);
  }
  return check_launch("cross_entropy_bwd", stream);
}

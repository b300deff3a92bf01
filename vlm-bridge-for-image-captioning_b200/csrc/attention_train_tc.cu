// Training attention on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory).
//
// Replaces F.scaled_dot_product_attention in the training forward (bridge_module.py:132-139: cross-attention,
// 8 heads of 288 over 257 / 1370 vision tokens; :230-237: non-causal, unmasked self-attention, 18 heads of 128),
// including the head split / merge views (:103-115, :201-213). The mma.sync kernels of attention.cu (which top
// out near 144 TF/s on B200's legacy tensor pipe) stay as the path for query blocks this kernel does not cover
// and as an A/B switch (B200B_ATTN_TC=0).
//
// One CTA per (128 query rows, head, image):
//   warp 5   TMA producer: the Q block once, then K and V tiles of 64 keys through two 2-deep rings, every tile
//            fetched as HD/32 boxes of [rows][32 columns] with the 64-byte swizzle (tensor maps over the
//            projection outputs, so the per-head column slice is read in place; rows past the sequence end
//            arrive as zeros)
//   warp 4   one elected thread issues the MMAs: S = Q K^T (M = 128, N = 64, K = HD; both operands K-major from
//            shared memory) into one of two S buffers in tensor memory, and O += P V (A = P read from TENSOR
//            memory, B = the V tile as an MN-major shared-memory operand -- V is stored [key][d], d contiguous;
//            N = HD is issued as 160 + 128 columns for HD = 288)
//   warps 0-3  one thread per query row (M = 128: TMEM lane = row): tcgen05.ld of the S row, online softmax in the
//            exp2 domain with a lazily updated reference maximum (O in tensor memory is only rescaled when a row
//            maximum grows by more than 2^8), dropout on the probabilities (Philox mask, same element indexing as
//            the backward kernels), P packed to bf16 and stored back into tensor memory as the A operand of the
//            P V product (it never touches shared memory); at the end O / l as bf16 to global memory and the
//            row's log-sum-exp for the backward.
// Operand layouts and descriptor forms were measured with tests/gpu_checks/probe_umma_layouts.cu
// (profiles/r02_probe_umma_layouts.jsonl): 64-byte-swizzle K-major tiles for any K that is a multiple of 16,
// MN-major tiles for K extents of 64 / 128 rows, A from tensor memory.
#include <stdlib.h>
#include <string.h>

#include "../../include/b200_bridge.h"
#include "common.cuh"
#include "launch.h"

namespace b200b {

struct FwdTcParams {
  __nv_bfloat16* o; long long ldo;
  float* lse2;                       // [B, H, Lq], log2 domain
  int B, H, Lq, Lk, Lkp;
  float scale_log2;
  DropoutCfg drop;
  uint32_t drop_stream;
};

int attn_tc_mask();

constexpr int kFBM = 128;            // query rows per CTA
constexpr int kFBN = 64;             // keys per tile
constexpr int kFThreads = 192;

template <int HD>
struct FwdTcCfg {
  static constexpr int kChunks = HD / 32;                 // 32-column (64-byte) chunks per row
  static constexpr int kQBytes = kFBM * HD * 2;
  static constexpr int kTileBytes = kFBN * HD * 2;        // one K or V tile
  static constexpr int kN0 = HD > 256 ? 160 : HD;         // P V issued as kN0 (+ kN1) output columns
  static constexpr int kN1 = HD - kN0;
  static constexpr int kSCol = 0;                         // TMEM: S buffers [0, 128)
  static constexpr int kPCol = 2 * kFBN;                  //       P buffers [128, 192) (bf16 pairs: 32 columns each)
  static constexpr int kOCol = kPCol + kFBN;              //       O [192, 192 + HD)
  static constexpr uint32_t kTmemCols = (kOCol + HD) <= 256 ? 256 : 512;
  static constexpr size_t kSmemBytes = (size_t)kQBytes + 4 * (size_t)kTileBytes + 1024;
  static_assert(HD % 32 == 0 && kN0 % 32 == 0 && kN1 % 32 == 0 && kN0 <= 256 && kOCol + HD <= 512, "head dim");
};

template <int HD>
__global__ void __launch_bounds__(kFThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const FwdTcParams p) {
  using Cfg = FwdTcCfg<HD>;
  extern __shared__ uint8_t smem_ftc_raw[];
  // q_full: Q landed; k_full / v_full: a tile landed; k_free / v_free: the MMAs that read it completed;
  // s_full: S(t) produced; p_full: P(t) stored (S(t) read, O rescaled if needed); pv_done: P V(t) completed
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_free[2], v_full[2], v_free[2], s_full[2], p_full[2], pv_done[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kFBM, h = blockIdx.y, b = blockIdx.z;
  const uint32_t s0 = (smem_u32(smem_ftc_raw) + 1023u) & ~1023u;
  const uint32_t q_tile = s0;
  const uint32_t k_tile0 = s0 + Cfg::kQBytes;
  const uint32_t v_tile0 = k_tile0 + 2 * Cfg::kTileBytes;
  const int nt = (p.Lk + kFBN - 1) / kFBN;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_free[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_free[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 5) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, Cfg::kQBytes);
#pragma unroll
      for (int c = 0; c < Cfg::kChunks; ++c)
        tma_load_4d(q_tile + (uint32_t)c * kFBM * 64, &tm_q, smem_u32(&q_full), c * 32, h, q0, b);
      for (int t = 0; t < nt; ++t) {
        const int st = t & 1;
        if (t >= 2) mbar_wait(&k_free[st], (uint32_t)(((t >> 1) - 1) & 1));
        mbar_arrive_expect_tx(&k_full[st], Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
          tma_load_4d(k_tile0 + (uint32_t)st * Cfg::kTileBytes + (uint32_t)c * kFBN * 64, &tm_k, smem_u32(&k_full[st]),
                      c * 32, h, t * kFBN, b);
        if (t >= 2) mbar_wait(&v_free[st], (uint32_t)(((t >> 1) - 1) & 1));
        mbar_arrive_expect_tx(&v_full[st], Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
          tma_load_4d(v_tile0 + (uint32_t)st * Cfg::kTileBytes + (uint32_t)c * kFBN * 64, &tm_v, smem_u32(&v_full[st]),
                      c * 32, h, t * kFBN, b);
      }
    }
  } else if (warp == 4) {
    // ------------------------------------ MMA issue ------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kFBM, kFBN, false, false);
      constexpr uint32_t idesc_o0 = umma_idesc_bf16(kFBM, Cfg::kN0, false, true);
      constexpr uint32_t idesc_o1 = umma_idesc_bf16(kFBM, Cfg::kN1 > 0 ? Cfg::kN1 : 16, false, true);
      auto issue_s = [&](int t) {
        const int st = t & 1;
        mbar_wait(&k_full[st], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        const uint32_t kt = k_tile0 + (uint32_t)st * Cfg::kTileBytes;
        const uint32_t d = tmem_base + (uint32_t)(Cfg::kSCol + st * kFBN);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(d, desc_sw64_kmajor(q_tile, kFBM, ks), desc_sw64_kmajor(kt, kFBN, ks), idesc_s, ks > 0);
        umma_commit(&s_full[st]);
        umma_commit(&k_free[st]);
      };
      mbar_wait(&q_full, 0);
      issue_s(0);
      if (nt > 1) issue_s(1);
      for (int t = 0; t < nt; ++t) {
        const int st = t & 1;
        mbar_wait(&p_full[st], (uint32_t)((t >> 1) & 1));
        mbar_wait(&v_full[st], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        const uint32_t vt = v_tile0 + (uint32_t)st * Cfg::kTileBytes;
        const uint32_t pa = tmem_base + (uint32_t)(Cfg::kPCol + st * (kFBN / 2));
        const uint32_t od = tmem_base + (uint32_t)Cfg::kOCol;
#pragma unroll
        for (int kc = 0; kc < kFBN / 16; ++kc) {
          umma_bf16_tmem_a(od, pa + (uint32_t)(kc * 8), desc_sw64_mnmajor(vt, kFBN, 0, kc), idesc_o0, (t | kc) != 0);
          if constexpr (Cfg::kN1 > 0)
            umma_bf16_tmem_a(od + (uint32_t)Cfg::kN0, pa + (uint32_t)(kc * 8),
                             desc_sw64_mnmajor(vt, kFBN, Cfg::kN0 / 32, kc), idesc_o1, (t | kc) != 0);
        }
        umma_commit(&pv_done[st]);
        umma_commit(&v_free[st]);
        // S(t + 2) overwrites the S buffer whose row the softmax warps finished reading before p_full(t)
        if (t + 2 < nt) issue_s(t + 2);
      }
    }
  } else {
    // ------------------------------------ softmax: one thread per query row ------------------------------------
    const DropoutCfg drop = dropout_resolve(p.drop);
    const int row = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const size_t bh = (size_t)b * p.H + h;
    const uint64_t ridx = (bh * p.Lq + (uint64_t)(q0 + row)) * (uint64_t)p.Lkp;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < nt; ++t) {
      const int st = t & 1;
      mbar_wait(&s_full[st], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      float sv[kFBN];
      float tmax = -INFINITY;
      {
        uint32_t v[kFBN / 16][16];
#pragma unroll
        for (int g = 0; g < kFBN / 16; ++g)
          tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kSCol + st * kFBN + g * 16), v[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < kFBN / 16; ++g)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = (t * kFBN + g * 16 + j < p.Lk) ? __uint_as_float(v[g][j]) * p.scale_log2 : -INFINITY;
            sv[g * 16 + j] = x;
            tmax = fmaxf(tmax, x);
          }
      }
      float alpha = 1.0f;
      if (tmax > m_run + 8.0f) {
        alpha = exp2f(m_run - tmax);   // 0 on the first tile
        m_run = tmax;
      }
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < kFBN; ++j) {
        sv[j] = exp2f(sv[j] - m_run);
        rs += sv[j];
      }
      l_run = l_run * alpha + rs;
      if (drop.thr != 0) {
#pragma unroll
        for (int j8 = 0; j8 < kFBN / 8; ++j8) {
          const uint4 bits = dropout_bits8(drop, p.drop_stream, (ridx + (uint64_t)(t * kFBN + j8 * 8)) >> 3);
#pragma unroll
          for (int e = 0; e < 8; ++e) sv[j8 * 8 + e] = dropout_keep(bits, e, drop.thr) ? sv[j8 * 8 + e] * drop.scale : 0.f;
        }
      }
      // P buffer st was last read by P V(t - 2)
      if (t >= 2) mbar_wait(&pv_done[st], (uint32_t)(((t >> 1) - 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < kFBN / 32; ++g) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = pack_bf16(sv[g * 32 + 2 * j], sv[g * 32 + 2 * j + 1]);
        tmem_st_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kPCol + st * (kFBN / 2) + g * 16), w);
      }
      if (t >= 1) {
        const unsigned moved = __ballot_sync(0xffffffffu, alpha != 1.0f);
        if (moved) {   // rare: a row maximum of this warp jumped; rescale its 32 rows of O once P V(t - 1) is in
          mbar_wait(&pv_done[(t - 1) & 1], (uint32_t)(((t - 1) >> 1) & 1));
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < HD; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kOCol + c0), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
            tmem_st_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kOCol + c0), v);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    }
    mbar_wait(&pv_done[(nt - 1) & 1], (uint32_t)(((nt - 1) >> 1) & 1));
    tc_fence_after();
    const bool valid = q0 + row < p.Lq;
    const float inv = 1.0f / l_run;
    if (valid) p.lse2[bh * p.Lq + q0 + row] = m_run + log2f(l_run);
    __nv_bfloat16* dst = p.o + ((size_t)b * p.Lq + (valid ? q0 + row : 0)) * p.ldo + (size_t)h * HD;
#pragma unroll 1
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t v0[16], v1[16];
      tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kOCol + c0), v0);
      tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kOCol + c0 + 16), v1);
      tmem_ld_wait();
      if (valid) {
        uint4 u[4];
        uint32_t* w = reinterpret_cast<uint32_t*>(u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          w[j] = pack_bf16(__uint_as_float(v0[2 * j]) * inv, __uint_as_float(v0[2 * j + 1]) * inv);
          w[8 + j] = pack_bf16(__uint_as_float(v1[2 * j]) * inv, __uint_as_float(v1[2 * j + 1]) * inv);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + c0 + j * 8) = u[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int HD>
static int launch_fwd_tc(const b200b_attn_args* a, cudaStream_t stream) {
  using Cfg = FwdTcCfg<HD>;
  static bool attr_done = false;   // idempotent; a race only repeats the call
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_last_error("attention_fwd (tcgen05): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_heads_sw64(&tq, a->q, a->ldq, HD, a->heads, a->len_q, a->batch, kFBM);
  if (rc != B200B_OK) return rc;
  rc = make_tmap_heads_sw64(&tk, a->k, a->ldk, HD, a->heads, a->len_k, a->batch, kFBN);
  if (rc != B200B_OK) return rc;
  rc = make_tmap_heads_sw64(&tv, a->v, a->ldv, HD, a->heads, a->len_k, a->batch, kFBN);
  if (rc != B200B_OK) return rc;
  FwdTcParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(a->o);
  p.ldo = a->ldo;
  p.lse2 = a->lse;
  p.B = a->batch; p.H = a->heads; p.Lq = a->len_q; p.Lk = a->len_k; p.Lkp = (a->len_k + 7) & ~7;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  p.drop_stream = a->dropout_stream;
  p.drop = make_dropout_cfg(a->dropout_p, a->seed, &p.drop_stream);
  dim3 grid((a->len_q + kFBM - 1) / kFBM, a->heads, a->batch);
  launch_pdl(kPdlAttn, attn_fwd_tc_kernel<HD>, grid, dim3(kFThreads), Cfg::kSmemBytes, stream, tq, tk, tv, p);
  return check_launch("attn_fwd_tc", stream);
}

// =====================================================================================================
// backward, query-major: dQ = scale * sum_t dS_t K_t           (autograd of bridge_module.py:132-139, 230-237)
// =====================================================================================================
// One CTA per (128 query rows, head, image). Q and dO stay resident (K-major A operands of S = Q K^T and
// dP = dO V^T); K and V tiles of 64 keys are streamed through ONE buffer each (two 128 x HD tiles + two 64 x HD
// tiles is all of the 227 KB at HD = 288): V is refilled as soon as dP(t) has been issued, K once dQ(t) -- which
// reads the same K tile again as an MN-major B operand -- has completed. The softmax warps rebuild P from the
// saved log-sum-exp, form dS = P * (dP - delta) with the forward's dropout mask and store it as bf16 into tensor
// memory, from where it is the A operand of dQ += dS K. Nothing is spilled to global memory (the mma.sync
// version wrote P and dS as [B, H, Lq, Lk] bf16 scratch for the key-major pass; here that pass recomputes them).
struct BwdTcParams {
  __nv_bfloat16* dq; long long lddq;
  __nv_bfloat16* dk; long long lddk;
  __nv_bfloat16* dv; long long lddv;
  const float* lse2;                  // [B, H, Lq]
  const float* delta;                 // [B, H, Lq]
  int B, H, Lq, Lk, Lkp;
  float scale, scale_log2;
  DropoutCfg drop;
  uint32_t drop_stream;
};

template <int HD>
struct DqTcCfg {
  static constexpr int kChunks = HD / 32;
  static constexpr int kQBytes = kFBM * HD * 2;
  static constexpr int kTileBytes = kFBN * HD * 2;
  static constexpr int kN0 = HD > 256 ? 160 : HD;
  static constexpr int kN1 = HD - kN0;
  static constexpr int kSCol = 0;                     // TMEM: S [0, 64), dP [64, 128)
  static constexpr int kDpCol = kFBN;
  static constexpr int kDsCol = 2 * kFBN;             //       dS buffers [128, 192) (bf16 pairs: 32 columns each)
  static constexpr int kDqCol = 3 * kFBN;             //       dQ [192, 192 + HD)
  static constexpr uint32_t kTmemCols = (kDqCol + HD) <= 256 ? 256 : 512;
  static constexpr size_t kSmemBytes = 2 * (size_t)kQBytes + 2 * (size_t)kTileBytes + 1024;
};

template <int HD>
__global__ void __launch_bounds__(kFThreads, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                      const BwdTcParams p) {
  using Cfg = DqTcCfg<HD>;
  extern __shared__ uint8_t smem_ftc_raw[];
  // per key tile t (phase parity t & 1 unless indexed): k_full / v_full: tile landed; v_free: dP(t) issued and done;
  // k_free: dQ(t) done; s_full: S(t) and dP(t) done; ds_full: dS(t) stored; dq_done[t & 1]: dQ(t) done
  __shared__ __align__(8) uint64_t qdo_full, k_full, k_free, v_full, v_free, s_full, ds_full, dq_done[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kFBM, h = blockIdx.y, b = blockIdx.z;
  const uint32_t s0 = (smem_u32(smem_ftc_raw) + 1023u) & ~1023u;
  const uint32_t q_tile = s0, do_tile = s0 + Cfg::kQBytes;
  const uint32_t k_tile = do_tile + Cfg::kQBytes, v_tile = k_tile + Cfg::kTileBytes;
  const int nt = (p.Lk + kFBN - 1) / kFBN;

  if (threadIdx.x == 0) {
    mbar_init(&qdo_full, 1);
    mbar_init(&k_full, 1);
    mbar_init(&k_free, 1);
    mbar_init(&v_full, 1);
    mbar_init(&v_free, 1);
    mbar_init(&s_full, 1);
    mbar_init(&ds_full, 4);
    mbar_init(&dq_done[0], 1);
    mbar_init(&dq_done[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 5) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&qdo_full, 2 * Cfg::kQBytes);
#pragma unroll
      for (int c = 0; c < Cfg::kChunks; ++c) {
        tma_load_4d(q_tile + (uint32_t)c * kFBM * 64, &tm_q, smem_u32(&qdo_full), c * 32, h, q0, b);
        tma_load_4d(do_tile + (uint32_t)c * kFBM * 64, &tm_do, smem_u32(&qdo_full), c * 32, h, q0, b);
      }
      for (int t = 0; t < nt; ++t) {
        if (t >= 1) mbar_wait(&k_free, (uint32_t)((t - 1) & 1));
        mbar_arrive_expect_tx(&k_full, Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
          tma_load_4d(k_tile + (uint32_t)c * kFBN * 64, &tm_k, smem_u32(&k_full), c * 32, h, t * kFBN, b);
        if (t >= 1) mbar_wait(&v_free, (uint32_t)((t - 1) & 1));
        mbar_arrive_expect_tx(&v_full, Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
          tma_load_4d(v_tile + (uint32_t)c * kFBN * 64, &tm_v, smem_u32(&v_full), c * 32, h, t * kFBN, b);
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kFBM, kFBN, false, false);
      constexpr uint32_t idesc_q0 = umma_idesc_bf16(kFBM, Cfg::kN0, false, true);
      constexpr uint32_t idesc_q1 = umma_idesc_bf16(kFBM, Cfg::kN1 > 0 ? Cfg::kN1 : 16, false, true);
      mbar_wait(&qdo_full, 0);
      for (int t = 0; t < nt; ++t) {
        const uint32_t ph = (uint32_t)(t & 1);
        // S(t) and dP(t): their TMEM buffers were read by the softmax warps before ds_full(t - 1), waited on below
        mbar_wait(&k_full, ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tmem_base + Cfg::kSCol, desc_sw64_kmajor(q_tile, kFBM, ks), desc_sw64_kmajor(k_tile, kFBN, ks), idesc_s,
                    ks > 0);
        mbar_wait(&v_full, ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tmem_base + Cfg::kDpCol, desc_sw64_kmajor(do_tile, kFBM, ks), desc_sw64_kmajor(v_tile, kFBN, ks),
                    idesc_s, ks > 0);
        umma_commit(&s_full);
        umma_commit(&v_free);
        mbar_wait(&ds_full, ph);
        tc_fence_after();
        const uint32_t da = tmem_base + (uint32_t)(Cfg::kDsCol + (t & 1) * (kFBN / 2));
        const uint32_t dd = tmem_base + (uint32_t)Cfg::kDqCol;
#pragma unroll
        for (int kc = 0; kc < kFBN / 16; ++kc) {
          umma_bf16_tmem_a(dd, da + (uint32_t)(kc * 8), desc_sw64_mnmajor(k_tile, kFBN, 0, kc), idesc_q0, (t | kc) != 0);
          if constexpr (Cfg::kN1 > 0)
            umma_bf16_tmem_a(dd + (uint32_t)Cfg::kN0, da + (uint32_t)(kc * 8),
                             desc_sw64_mnmajor(k_tile, kFBN, Cfg::kN0 / 32, kc), idesc_q1, (t | kc) != 0);
        }
        umma_commit(&dq_done[t & 1]);
        umma_commit(&k_free);
      }
    }
  } else {
    const DropoutCfg drop = dropout_resolve(p.drop);
    const int row = warp * 32 + lane;
    const bool valid = q0 + row < p.Lq;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const size_t bh = (size_t)b * p.H + h;
    const uint64_t ridx = (bh * p.Lq + (uint64_t)(q0 + row)) * (uint64_t)p.Lkp;
    const float lse = valid ? p.lse2[bh * p.Lq + q0 + row] : 0.f;
    const float del = valid ? p.delta[bh * p.Lq + q0 + row] : 0.f;
    for (int t = 0; t < nt; ++t) {
      mbar_wait(&s_full, (uint32_t)(t & 1));
      tc_fence_after();
      // dS buffer t & 1 was last read by dQ(t - 2)
      if (t >= 2) mbar_wait(&dq_done[t & 1], (uint32_t)(((t >> 1) - 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < kFBN / 32; ++half) {
        uint32_t sv[2][16], dv[2][16];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kSCol + half * 32 + g * 16), sv[g]);
          tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kDpCol + half * 32 + g * 16), dv[g]);
        }
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float ds[16];
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            const int col0 = t * kFBN + half * 32 + g * 16 + j8 * 8;
            uint4 bits = make_uint4(0, 0, 0, 0);
            if (drop.thr != 0) bits = dropout_bits8(drop, p.drop_stream, (ridx + (uint64_t)col0) >> 3);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = j8 * 8 + e;
              const float pr = (col0 + e < p.Lk) ? exp2f(__uint_as_float(sv[g][j]) * p.scale_log2 - lse) : 0.f;
              float dpr = __uint_as_float(dv[g][j]);
              if (drop.thr != 0) dpr = dropout_keep(bits, e, drop.thr) ? dpr * drop.scale : 0.f;
              ds[j] = pr * (dpr - del);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) w[g * 8 + j] = pack_bf16(ds[2 * j], ds[2 * j + 1]);
        }
        tmem_st_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kDsCol + (t & 1) * (kFBN / 2) + half * 16), w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full);
    }
    mbar_wait(&dq_done[(nt - 1) & 1], (uint32_t)(((nt - 1) >> 1) & 1));
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + ((size_t)b * p.Lq + (valid ? q0 + row : 0)) * p.lddq + (size_t)h * HD;
#pragma unroll 1
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t v0[16], v1[16];
      tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kDqCol + c0), v0);
      tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kDqCol + c0 + 16), v1);
      tmem_ld_wait();
      if (valid) {
        uint4 u[4];
        uint32_t* w = reinterpret_cast<uint32_t*>(u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          w[j] = pack_bf16(__uint_as_float(v0[2 * j]) * p.scale, __uint_as_float(v0[2 * j + 1]) * p.scale);
          w[8 + j] = pack_bf16(__uint_as_float(v1[2 * j]) * p.scale, __uint_as_float(v1[2 * j + 1]) * p.scale);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + c0 + j * 8) = u[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// =====================================================================================================
// backward, key-major: dV = Pd^T dO, dK = scale * dS^T Q, for query blocks of up to 128 rows (the training shapes)
// =====================================================================================================
// One CTA per (key block(s) of 128 keys, head, image) -- one block while the grid fits two waves of the SMs, several
// consecutive blocks with Q / dO kept resident for long key sequences -- in a TRANSPOSED formulation:
// S^T = K_j Q^T and dP^T = V_j dO^T put a key on every TMEM lane, so P^T and dS^T come out in exactly the layout the tensor-memory A operand of dV = P^T dO and
// dK = dS^T Q needs (lane = key = output row, columns = queries = contraction index); dO and Q are then read a second
// time from the SAME shared-memory tiles as MN-major B operands. Three 128 x HD tiles are resident (K_j, later
// overwritten by V_j; Q; dO) -- 221 KB at HD = 288. The bf16 P^T / dS^T overwrite the fp32 S^T / dP^T columns they
// were computed from (slice by slice), which leaves room for one 128 x HD fp32 output: dV is produced and drained,
// then dK.
// A thread owns a key and walks over queries, but the dropout mask is indexed [query][8 consecutive keys per Philox
// call] (the forward's natural order), so the 32 lanes of a warp -- 4 key groups -- share the work: each lane
// evaluates Philox for 4 of the 32 (query, key group) pairs of a slice and the bits are exchanged with shuffles.
template <int HD>
struct DkvTcCfg {
  static constexpr int kChunks = HD / 32;
  static constexpr int kTileBytes = kFBM * HD * 2;     // 128 rows
  static constexpr int kN0 = HD > 256 ? 160 : HD;
  static constexpr int kN1 = HD - kN0;
  static constexpr int kStCol = 0;                     // TMEM: S^T [0, 128) -> P^T bf16 in [0, 64)
  static constexpr int kDptCol = 128;                  //       dP^T [128, 256) -> dS^T bf16 in [128, 192)
  static constexpr int kOutCol = 192;                  //       dV, then dK: [192, 192 + HD)
  static constexpr uint32_t kTmemCols = 512;
  static constexpr size_t kSmemBytes = 3 * (size_t)kTileBytes + 1024;
  static_assert(kOutCol + HD <= 512, "head dim");
};

template <int HD>
__global__ void __launch_bounds__(kFThreads, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                       const BwdTcParams p, const int blocks_per_cta) {
  using Cfg = DkvTcCfg<HD>;
  extern __shared__ uint8_t smem_ftc_raw[];
  // q_full / do_full: the query block's Q / dO landed (once per CTA). Per key block `it` of this CTA (phase it & 1):
  // k_full / v_full: K_j / V_j landed in the A buffer; k_free / v_free: S^T / dP^T have consumed it;
  // st_full / dpt_full: S^T / dP^T produced; pds_full: P^T and dS^T stored; dv_done / dk_done: dV / dK produced;
  // dv_drained / dk_drained: the output columns have been read out
  __shared__ __align__(8) uint64_t q_full, do_full, k_full, v_full, k_free, v_free, st_full, dpt_full, pds_full, dv_done,
      dv_drained, dk_done, dk_drained;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float lse_s[kFBM], del_s[kFBM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nkb = (p.Lk + kFBM - 1) / kFBM;
  const int jb0 = blockIdx.x * blocks_per_cta;
  const int nit = min(blocks_per_cta, nkb - jb0);         // key blocks of this CTA (>= 1 by construction of the grid)
  const uint32_t s0 = (smem_u32(smem_ftc_raw) + 1023u) & ~1023u;
  const uint32_t a_tile = s0, q_tile = s0 + Cfg::kTileBytes, do_tile = q_tile + Cfg::kTileBytes;
  const size_t bh = (size_t)b * p.H + h;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    mbar_init(&do_full, 1);
    mbar_init(&k_full, 1);
    mbar_init(&v_full, 1);
    mbar_init(&k_free, 1);
    mbar_init(&v_free, 1);
    mbar_init(&st_full, 1);
    mbar_init(&dpt_full, 1);
    mbar_init(&pds_full, 4);
    mbar_init(&dv_done, 1);
    mbar_init(&dv_drained, 4);
    mbar_init(&dk_done, 1);
    mbar_init(&dk_drained, 4);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < kFBM) {
    const bool ok = (int)threadIdx.x < p.Lq;
    lse_s[threadIdx.x] = ok ? p.lse2[bh * p.Lq + threadIdx.x] : 0.f;
    del_s[threadIdx.x] = ok ? p.delta[bh * p.Lq + threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 5) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, Cfg::kTileBytes);
#pragma unroll
      for (int c = 0; c < Cfg::kChunks; ++c)
        tma_load_4d(q_tile + (uint32_t)c * kFBM * 64, &tm_q, smem_u32(&q_full), c * 32, h, 0, b);
      for (int it = 0; it < nit; ++it) {
        const int j0 = (jb0 + it) * kFBM;
        const uint32_t prev = (uint32_t)((it - 1) & 1);
        if (it >= 1) mbar_wait(&v_free, prev);       // dP^T of the previous key block has consumed V in the A buffer
        mbar_arrive_expect_tx(&k_full, Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
          tma_load_4d(a_tile + (uint32_t)c * kFBM * 64, &tm_k, smem_u32(&k_full), c * 32, h, j0, b);
        if (it == 0) {
          mbar_arrive_expect_tx(&do_full, Cfg::kTileBytes);
#pragma unroll
          for (int c = 0; c < Cfg::kChunks; ++c)
            tma_load_4d(do_tile + (uint32_t)c * kFBM * 64, &tm_do, smem_u32(&do_full), c * 32, h, 0, b);
        }
        mbar_wait(&k_free, (uint32_t)(it & 1));      // S^T has consumed K_j: V_j takes its place
        mbar_arrive_expect_tx(&v_full, Cfg::kTileBytes);
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
          tma_load_4d(a_tile + (uint32_t)c * kFBM * 64, &tm_v, smem_u32(&v_full), c * 32, h, j0, b);
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kFBM, kFBM, false, false);
      constexpr uint32_t idesc_o0 = umma_idesc_bf16(kFBM, Cfg::kN0, false, true);
      constexpr uint32_t idesc_o1 = umma_idesc_bf16(kFBM, Cfg::kN1 > 0 ? Cfg::kN1 : 16, false, true);
      const uint32_t od = tmem_base + (uint32_t)Cfg::kOutCol;
      mbar_wait(&q_full, 0);
      for (int it = 0; it < nit; ++it) {
        const uint32_t ph = (uint32_t)(it & 1);
        // S^T overwrites the columns P^T(it - 1) lived in: dV(it - 1) has read them (dv_done was waited for by the
        // epilogue before dv_drained, which this thread waited for before issuing dK(it - 1))
        mbar_wait(&k_full, ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tmem_base + Cfg::kStCol, desc_sw64_kmajor(a_tile, kFBM, ks), desc_sw64_kmajor(q_tile, kFBM, ks), idesc_s,
                    ks > 0);
        umma_commit(&st_full);
        umma_commit(&k_free);
        if (it == 0) mbar_wait(&do_full, 0);
        mbar_wait(&v_full, ph);
        tc_fence_after();
        // dP^T overwrites the columns dS^T(it - 1) lived in; dK(it - 1), issued earlier by this thread, completes first
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          umma_bf16(tmem_base + Cfg::kDptCol, desc_sw64_kmajor(a_tile, kFBM, ks), desc_sw64_kmajor(do_tile, kFBM, ks),
                    idesc_s, ks > 0);
        umma_commit(&dpt_full);
        umma_commit(&v_free);
        mbar_wait(&pds_full, ph);
        if (it >= 1) mbar_wait(&dk_drained, (uint32_t)((it - 1) & 1));   // the output columns still hold dK(it - 1)
        tc_fence_after();
#pragma unroll
        for (int kc = 0; kc < kFBM / 16; ++kc) {      // dV = Pd^T dO
          umma_bf16_tmem_a(od, tmem_base + (uint32_t)(Cfg::kStCol + kc * 8), desc_sw64_mnmajor(do_tile, kFBM, 0, kc),
                           idesc_o0, kc != 0);
          if constexpr (Cfg::kN1 > 0)
            umma_bf16_tmem_a(od + (uint32_t)Cfg::kN0, tmem_base + (uint32_t)(Cfg::kStCol + kc * 8),
                             desc_sw64_mnmajor(do_tile, kFBM, Cfg::kN0 / 32, kc), idesc_o1, kc != 0);
        }
        umma_commit(&dv_done);
        mbar_wait(&dv_drained, ph);
        tc_fence_after();
#pragma unroll
        for (int kc = 0; kc < kFBM / 16; ++kc) {      // dK = dS^T Q
          umma_bf16_tmem_a(od, tmem_base + (uint32_t)(Cfg::kDptCol + kc * 8), desc_sw64_mnmajor(q_tile, kFBM, 0, kc),
                           idesc_o0, kc != 0);
          if constexpr (Cfg::kN1 > 0)
            umma_bf16_tmem_a(od + (uint32_t)Cfg::kN0, tmem_base + (uint32_t)(Cfg::kDptCol + kc * 8),
                             desc_sw64_mnmajor(q_tile, kFBM, Cfg::kN0 / 32, kc), idesc_o1, kc != 0);
        }
        umma_commit(&dk_done);
      }
    }
  } else {
    const DropoutCfg drop = dropout_resolve(p.drop);
    const int row = warp * 32 + lane;          // key j0 + row lives on TMEM lane `row`
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int it = 0; it < nit; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int j0 = (jb0 + it) * kFBM;
      const int key = j0 + row;
      const bool key_ok = key < p.Lk;
      const uint64_t kgrp = (uint64_t)((j0 + warp * 32) >> 3) + (uint64_t)(lane >> 3);   // this lane's Philox key group
      mbar_wait(&st_full, ph);
      mbar_wait(&dpt_full, ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kFBM / 32; ++c) {      // 32 queries per slice
        uint32_t sv[2][16], dv[2][16];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kStCol + c * 32 + g * 16), sv[g]);
          tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kDptCol + c * 32 + g * 16), dv[g]);
        }
        tmem_ld_wait();
        // dropout bits of the slice: lane L evaluates (query c*32 + r*8 + (L & 7), key group L >> 3) for r = 0..3
        uint4 bits[4];
        if (drop.thr != 0) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const uint64_t qi = (uint64_t)(c * 32 + r * 8 + (lane & 7));
            bits[r] = dropout_bits8(drop, p.drop_stream, ((bh * p.Lq + qi) * (uint64_t)p.Lkp >> 3) + kgrp);
          }
        }
        uint32_t wp[16], wd[16];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float pv[16], dsv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int qi = c * 32 + g * 16 + j;       // query index inside the block
            float pr = (key_ok && qi < p.Lq) ? exp2f(__uint_as_float(sv[g][j]) * p.scale_log2 - lse_s[qi]) : 0.f;
            float dpr = __uint_as_float(dv[g][j]);
            float prd = pr;
            if (drop.thr != 0) {
              const int r = (g * 16 + j) >> 3, m = (g * 16 + j) & 7;
              const int src = (lane & 24) | m;
              uint4 bb;
              bb.x = __shfl_sync(0xffffffffu, bits[r].x, src);
              bb.y = __shfl_sync(0xffffffffu, bits[r].y, src);
              bb.z = __shfl_sync(0xffffffffu, bits[r].z, src);
              bb.w = __shfl_sync(0xffffffffu, bits[r].w, src);
              const bool keep = dropout_keep(bb, lane & 7, drop.thr);
              prd = keep ? pr * drop.scale : 0.f;
              dpr = keep ? dpr * drop.scale : 0.f;
            }
            pv[j] = prd;
            dsv[j] = pr * (dpr - del_s[qi]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            wp[g * 8 + j] = pack_bf16(pv[2 * j], pv[2 * j + 1]);
            wd[g * 8 + j] = pack_bf16(dsv[2 * j], dsv[2 * j + 1]);
          }
        }
        // in place: the bf16 slice [16c, 16c + 16) lies inside the fp32 columns [0, 32c + 32) already consumed
        tmem_st_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kStCol + c * 16), wp);
        tmem_st_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kDptCol + c * 16), wd);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full);
      auto drain = [&](__nv_bfloat16* base, long long ld, float mul) {
        __nv_bfloat16* dst = base + ((size_t)b * p.Lk + (key_ok ? key : 0)) * ld + (size_t)h * HD;
#pragma unroll 1
        for (int c0 = 0; c0 < HD; c0 += 32) {
          uint32_t v0[16], v1[16];
          tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kOutCol + c0), v0);
          tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(Cfg::kOutCol + c0 + 16), v1);
          tmem_ld_wait();
          if (key_ok) {
            uint4 u[4];
            uint32_t* w = reinterpret_cast<uint32_t*>(u);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              w[j] = pack_bf16(__uint_as_float(v0[2 * j]) * mul, __uint_as_float(v0[2 * j + 1]) * mul);
              w[8 + j] = pack_bf16(__uint_as_float(v1[2 * j]) * mul, __uint_as_float(v1[2 * j + 1]) * mul);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + c0 + j * 8) = u[j];
          }
        }
      };
      mbar_wait(&dv_done, ph);
      tc_fence_after();
      drain(p.dv, p.lddv, 1.0f);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dv_drained);
      mbar_wait(&dk_done, ph);
      tc_fence_after();
      drain(p.dk, p.lddk, p.scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dk_drained);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int HD>
static int launch_bwd_tc(const b200b_attn_args* a, const float* delta, cudaStream_t stream) {
  static bool attr_done = false;   // idempotent; a race only repeats the calls
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)DqTcCfg<HD>::kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)DkvTcCfg<HD>::kSmemBytes);
    if (e != cudaSuccess) {
      set_last_error("attention_bwd (tcgen05): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  BwdTcParams p;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.lddq = a->lddq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.lddk = a->lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.lddv = a->lddv;
  p.lse2 = a->lse;
  p.delta = delta;
  p.B = a->batch; p.H = a->heads; p.Lq = a->len_q; p.Lk = a->len_k; p.Lkp = (a->len_k + 7) & ~7;
  p.scale = 1.0f / sqrtf((float)HD);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.drop_stream = a->dropout_stream;
  p.drop = make_dropout_cfg(a->dropout_p, a->seed, &p.drop_stream);
  CUtensorMap tq, td, tk, tv;
  int rc;
  // query-major pass: Q / dO blocks of 128 rows, K / V tiles of 64 keys
  if ((rc = make_tmap_heads_sw64(&tq, a->q, a->ldq, HD, a->heads, a->len_q, a->batch, kFBM)) != B200B_OK) return rc;
  if ((rc = make_tmap_heads_sw64(&td, a->d_o, a->lddo, HD, a->heads, a->len_q, a->batch, kFBM)) != B200B_OK) return rc;
  if ((rc = make_tmap_heads_sw64(&tk, a->k, a->ldk, HD, a->heads, a->len_k, a->batch, kFBN)) != B200B_OK) return rc;
  if ((rc = make_tmap_heads_sw64(&tv, a->v, a->ldv, HD, a->heads, a->len_k, a->batch, kFBN)) != B200B_OK) return rc;
  dim3 gq((a->len_q + kFBM - 1) / kFBM, a->heads, a->batch);
  launch_pdl(kPdlAttn, attn_bwd_dq_tc_kernel<HD>, gq, dim3(kFThreads), DqTcCfg<HD>::kSmemBytes, stream, tq, td, tk, tv, p);
  if ((rc = check_launch("attn_bwd_dq_tc", stream)) != B200B_OK) return rc;
  // key-major pass: K / V blocks of 128 keys
  if ((rc = make_tmap_heads_sw64(&tk, a->k, a->ldk, HD, a->heads, a->len_k, a->batch, kFBM)) != B200B_OK) return rc;
  if ((rc = make_tmap_heads_sw64(&tv, a->v, a->ldv, HD, a->heads, a->len_k, a->batch, kFBM)) != B200B_OK) return rc;
  // Key blocks per CTA: one while the grid fits two waves of the SMs (a CTA re-loads the query block's Q and dO, 2 tiles
  // of the 3 it holds, for every key block), more for long key sequences (1370 vision tokens = 11 key blocks per head)
  // so that Q / dO are fetched once per CTA and the blocks of a head are spread evenly over its CTAs.
  int num_sms = 0;
  if ((rc = device_sm_count(&num_sms)) != B200B_OK) return rc;
  const int nkb = (a->len_k + kFBM - 1) / kFBM;
  const long long items = (long long)nkb * a->heads * a->batch;
  int bpc = (int)((items + 2LL * num_sms - 1) / (2LL * num_sms));
  if (bpc < 1) bpc = 1;
  const int ctas_per_head = (nkb + bpc - 1) / bpc;
  bpc = (nkb + ctas_per_head - 1) / ctas_per_head;
  dim3 gk((nkb + bpc - 1) / bpc, a->heads, a->batch);
  launch_pdl(kPdlAttn, attn_bwd_dkv_tc_kernel<HD>, gk, dim3(kFThreads), DkvTcCfg<HD>::kSmemBytes, stream, tq, td, tk, tv, p,
             bpc);
  return check_launch("attn_bwd_dkv_tc", stream);
}

// Backward through the tcgen05 kernels when the shape qualifies (query blocks of at most 128 rows: the key-major pass
// keeps one query block resident); `delta` = rowsum(dO * O) has been computed by the caller.
int attention_bwd_tc(const b200b_attn_args* a, const float* delta, cudaStream_t stream, int* taken) {
  *taken = 0;
  if (!(attn_tc_mask() & 2) || a->len_q > kFBM) return B200B_OK;
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
                       reinterpret_cast<uintptr_t>(a->d_o) | reinterpret_cast<uintptr_t>(a->dq) |
                       reinterpret_cast<uintptr_t>(a->dk) | reinterpret_cast<uintptr_t>(a->dv);
  if ((al & 15) || (a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->lddo % 8) || (a->lddq % 8) || (a->lddk % 8) ||
      (a->lddv % 8))
    return B200B_OK;
  int rc;
  switch (a->head_dim) {
    case 64: rc = launch_bwd_tc<64>(a, delta, stream); break;
    case 128: rc = launch_bwd_tc<128>(a, delta, stream); break;
    case 288: rc = launch_bwd_tc<288>(a, delta, stream); break;
    default: return B200B_OK;
  }
  *taken = 1;
  return rc;
}

// bit 0 = tcgen05 forward, bit 1 = tcgen05 backward where they apply (default 3); the environment variable
// B200B_ATTN_TC sets the initial value, b200b_attention_set_tc() changes it (A/B measurements, tests)
static int g_attn_tc = -1;
int attn_tc_mask() {
  if (g_attn_tc < 0) {
    const char* e = getenv("B200B_ATTN_TC");
    g_attn_tc = e ? atoi(e) : 3;
  }
  return g_attn_tc;
}

// Forward through the tcgen05 kernel when the shape qualifies; returns 1 if it was not taken (caller falls back).
int attention_fwd_tc(const b200b_attn_args* a, cudaStream_t stream, int* taken) {
  *taken = 0;
  if (!(attn_tc_mask() & 1)) return B200B_OK;
  // tensor maps need 16-byte aligned bases and row pitches; the head slices must tile into 32-column boxes
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
                       reinterpret_cast<uintptr_t>(a->o);
  if ((al & 15) || (a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 8)) return B200B_OK;
  int rc;
  switch (a->head_dim) {
    case 64: rc = launch_fwd_tc<64>(a, stream); break;
    case 128: rc = launch_fwd_tc<128>(a, stream); break;
    case 288: rc = launch_fwd_tc<288>(a, stream); break;
    default: return B200B_OK;
  }
  *taken = 1;
  return rc;
}

}  // namespace b200b

extern "C" int b200b_attention_set_tc(int mask) {
  const int prev = b200b::attn_tc_mask();
  if (mask >= 0) b200b::g_attn_tc = mask;
  return prev;
}

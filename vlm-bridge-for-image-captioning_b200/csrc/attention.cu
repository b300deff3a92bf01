// Fused multi-head attention for the bridge (forward + backward), bf16 operands / fp32 softmax.
//
// Replaces F.scaled_dot_product_attention at bridge_module.py:132-139 (cross-attention, 8 heads of
// 288 over 257 / 1370 vision tokens) and :230-237 (non-causal, unmasked self-attention, 18 heads of
// 128), including the head split/merge views (:103-115, :201-213): Q/K/V are read in place from the
// projection outputs ([tokens, heads*d] with arbitrary row pitch) and O is written head-merged.
//
// Data movement: K/V (and Q/dO) rows are staged into shared memory with TMA bulk copies
// (cp.async.bulk, one row = one head slice, completion on an mbarrier), double buffered so the
// copy of key tile t+1 overlaps the math of tile t. Rows are padded by 16 bytes in shared memory,
// which makes every ldmatrix access bank-conflict free for d in {64, 128, 288}.
// Math: mma.sync m16n8k16 (bf16 -> fp32); online softmax in the exp2 domain with the row max / row
// sum reduced across the 4 lanes that share a row via warp shuffles; O and dQ accumulators
// (16 rows x d per warp) live in registers.
//
// Backward is three kernels: row dots delta = sum(dO*O); a query-major pass that recomputes
// P from the saved log-sum-exp, forms dS, accumulates dQ and spills P(dropped) and dS as bf16 to a
// scratch [B,H,Lq,Lkp]; and a key-major pass dV = P^T dO, dK = dS^T Q that reads that scratch.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "launch.h"

namespace b200b {

// attention_train_tc.cu: the tcgen05 forward; *taken = 1 when it launched (or failed), 0 when the shape is not its
int attention_fwd_tc(const b200b_attn_args* a, cudaStream_t stream, int* taken);
// the tcgen05 backward (dQ pass + dK / dV pass) given delta = rowsum(dO * O)
int attention_bwd_tc(const b200b_attn_args* a, const float* delta, cudaStream_t stream, int* taken);

struct AttnParams {
  const __nv_bfloat16* q; long long ldq;
  const __nv_bfloat16* k; long long ldk;
  const __nv_bfloat16* v; long long ldv;
  __nv_bfloat16* o; long long ldo;        // fwd: output; bwd: forward output (read)
  const __nv_bfloat16* d_o; long long lddo;
  __nv_bfloat16* dq; long long lddq;
  __nv_bfloat16* dk; long long lddk;
  __nv_bfloat16* dv; long long lddv;
  float* lse2;                             // [B,H,Lq] log2-domain log-sum-exp of the scaled scores
  float* delta;                            // [B,H,Lq]
  __nv_bfloat16* p_scr;                    // [B,H,Lq,Lkp] dropped probabilities
  __nv_bfloat16* ds_scr;                   // [B,H,Lq,Lkp] dS (unscaled)
  int B, H, Lq, Lk, Lkp;
  long long kv_bstride, kv_hstride;        // packed K/V cache (decode): element strides per sample / head
  float scale, scale_log2;
  DropoutCfg drop;
  uint32_t drop_stream;
};

template <int HD>
struct AttnCfg {
  static constexpr int kBM = 64;                      // query rows per CTA: 4 warps x 16
  static constexpr int kBN = (HD > 128) ? 32 : 64;    // keys per tile
  static constexpr int kLd = HD + 8;                  // padded smem row, elements
  static constexpr int kRowBytes = HD * 2;
  static constexpr int kQT = 32;                      // query rows per step of the key-major pass
  static constexpr int kScrLd = 40;                   // padded scratch-tile row (32 keys + 8)
  static constexpr size_t kFwdSmem = (size_t)(kBM + 4 * kBN) * kLd * 2;
  static constexpr size_t kBwdQSmem = (size_t)(2 * kBM + 4 * kBN) * kLd * 2;
  static constexpr size_t kBwdKVSmem = (size_t)2 * (2 * kQT * kScrLd + 2 * kQT * kLd) * 2;
};

__device__ __forceinline__ void zero_row(__nv_bfloat16* row, int bytes, int lane, int nlanes) {
  for (int c = lane * 16; c < bytes; c += nlanes * 16)
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(row) + c) = make_uint4(0, 0, 0, 0);
}

// Stage one K/V tile (keys [j0, j0+kBN) of sample b, head h) into `ks`/`vs`; executed by warp 0.
template <int HD>
__device__ __forceinline__ void issue_kv_tile(const AttnParams& p, int b, int h, int j0, __nv_bfloat16* ks,
                                              __nv_bfloat16* vs, uint64_t* bar, int lane) {
  using Cfg = AttnCfg<HD>;
  const int nk = min(Cfg::kBN, p.Lk - j0);
  // rows past the end of the sequence must read as zeros (0 * garbage could be NaN)
  for (int r = nk; r < Cfg::kBN; ++r) {
    zero_row(ks + (size_t)r * Cfg::kLd, Cfg::kRowBytes, lane, 32);
    zero_row(vs + (size_t)r * Cfg::kLd, Cfg::kRowBytes, lane, 32);
  }
  if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(2 * nk * Cfg::kRowBytes));
  __syncwarp();
  for (int r = lane; r < nk; r += 32) {
    const size_t grow = (size_t)b * p.Lk + j0 + r;
    bulk_load_1d(ks + (size_t)r * Cfg::kLd, p.k + grow * p.ldk + (size_t)h * HD, Cfg::kRowBytes, bar);
    bulk_load_1d(vs + (size_t)r * Cfg::kLd, p.v + grow * p.ldv + (size_t)h * HD, Cfg::kRowBytes, bar);
  }
}

// S[16 x BN] (+)= A[16 x HD] * Bt[BN x HD]^T with A rows at `a_rows` and Bt rows at `b_rows` (both
// padded row-major in shared memory). Used for Q K^T and dO V^T.
template <int HD, int BN>
__device__ __forceinline__ void mma_rows_x_rows(float (&s)[BN / 8][4], const __nv_bfloat16* a_rows,
                                                const __nv_bfloat16* b_rows, int lane) {
  constexpr int kLd = HD + 8;
  const uint32_t a_base = smem_u32(a_rows + (size_t)(lane & 15) * kLd + (lane >> 4) * 8);
  const uint32_t b_base = smem_u32(b_rows + (size_t)((lane & 7) + ((lane >> 4) << 3)) * kLd + ((lane >> 3) & 1) * 8);
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    uint32_t a[4];
    ldmatrix_x4(a, a_base + ks * 32);
#pragma unroll
    for (int np = 0; np < BN / 16; ++np) {
      uint32_t bb[4];
      ldmatrix_x4(bb, b_base + (uint32_t)(np * 16 * kLd * 2) + ks * 32);
      const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
      mma_m16n8k16(s[2 * np], a, b0);
      mma_m16n8k16(s[2 * np + 1], a, b1);
    }
  }
}

// acc[16 x HD] += P[16 x BN] (bf16 A fragments built from the C-fragment layout) * Bk[BN x HD]
// with Bk rows (the contraction index) in padded row-major shared memory. Used for P V and dS K.
template <int HD, int BN>
__device__ __forceinline__ void mma_frag_x_cols(float (&acc)[HD / 8][4], const float (&pf)[BN / 8][4],
                                                const __nv_bfloat16* b_rows, int lane) {
  constexpr int kLd = HD + 8;
  const uint32_t b_base = smem_u32(b_rows + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * kLd + (lane >> 4) * 8);
#pragma unroll
  for (int kk = 0; kk < BN / 16; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(pf[2 * kk][0], pf[2 * kk][1]);
    a[1] = pack_bf16(pf[2 * kk][2], pf[2 * kk][3]);
    a[2] = pack_bf16(pf[2 * kk + 1][0], pf[2 * kk + 1][1]);
    a[3] = pack_bf16(pf[2 * kk + 1][2], pf[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < HD / 16; ++dp) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, b_base + (uint32_t)(kk * 16 * kLd * 2) + dp * 32);
      const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
      mma_m16n8k16(acc[2 * dp], a, b0);
      mma_m16n8k16(acc[2 * dp + 1], a, b1);
    }
  }
}

// Write this warp's 16 x HD accumulator (scaled) as bf16 through its own staging rows, then copy
// the rows out with 16-byte coalesced stores.
template <int HD>
__device__ __forceinline__ void store_rows_bf16(const float (&acc)[HD / 8][4], float s0, float s1,
                                                __nv_bfloat16* stage_rows /*16 rows, padded*/,
                                                __nv_bfloat16* gdst /*row 0 of this warp*/, long long ld,
                                                int rows_valid, int lane) {
  constexpr int kLd = HD + 8;
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    *reinterpret_cast<uint32_t*>(stage_rows + (size_t)g * kLd + n * 8 + 2 * t) = pack_bf16(acc[n][0] * s0, acc[n][1] * s0);
    *reinterpret_cast<uint32_t*>(stage_rows + (size_t)(g + 8) * kLd + n * 8 + 2 * t) =
        pack_bf16(acc[n][2] * s1, acc[n][3] * s1);
  }
  __syncwarp();
  constexpr int kChunks = HD / 8;  // 16-byte chunks per row
  for (int idx = lane; idx < 16 * kChunks; idx += 32) {
    const int r = idx / kChunks, c = idx % kChunks;
    if (r < rows_valid)
      *reinterpret_cast<uint4*>(gdst + (size_t)r * ld + c * 8) =
          *reinterpret_cast<const uint4*>(stage_rows + (size_t)r * kLd + c * 8);
  }
}

// ================================================================================================
// forward
// ================================================================================================
template <int HD>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
  pdl_prologue();
  const DropoutCfg drop = dropout_resolve(p.drop);
  using Cfg = AttnCfg<HD>;
  constexpr int BN = Cfg::kBN, kLd = Cfg::kLd;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Ks = Qs + (size_t)Cfg::kBM * kLd;   // [2][BN][kLd]
  __nv_bfloat16* Vs = Ks + (size_t)2 * BN * kLd;     // [2][BN][kLd]
  __shared__ __align__(8) uint64_t qbar, full_bar[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * Cfg::kBM;
  const int nq = min(Cfg::kBM, p.Lq - q0);
  const int nt = (p.Lk + BN - 1) / BN;

  if (threadIdx.x == 0) {
    mbar_init(&qbar, 1);
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) mbar_arrive_expect_tx(&qbar, (uint32_t)(nq * Cfg::kRowBytes));
    __syncwarp();
    for (int r = lane; r < nq; r += 32)
      bulk_load_1d(Qs + (size_t)r * kLd, p.q + ((size_t)b * p.Lq + q0 + r) * p.ldq + (size_t)h * HD, Cfg::kRowBytes,
                   &qbar);
    issue_kv_tile<HD>(p, b, h, 0, Ks, Vs, &full_bar[0], lane);
  }
  __syncthreads();  // zero-filled rows visible to all warps

  float o_acc[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o_acc[n][0] = o_acc[n][1] = o_acc[n][2] = o_acc[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int g = lane >> 2, t4 = lane & 3;
  const size_t bh = (size_t)b * p.H + h;

  mbar_wait(&qbar, 0);
  for (int t = 0; t < nt; ++t) {
    const int stage = t & 1;
    if (warp == 0 && t + 1 < nt)
      issue_kv_tile<HD>(p, b, h, (t + 1) * BN, Ks + (size_t)(stage ^ 1) * BN * kLd, Vs + (size_t)(stage ^ 1) * BN * kLd,
                        &full_bar[stage ^ 1], lane);
    mbar_wait(&full_bar[stage], (uint32_t)((t >> 1) & 1));

    float s[BN / 8][4];
#pragma unroll
    for (int n = 0; n < BN / 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    mma_rows_x_rows<HD, BN>(s, Qs + (size_t)warp * 16 * kLd, Ks + (size_t)stage * BN * kLd, lane);

    const int j0 = t * BN;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < BN / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = j0 + n * 8 + 2 * t4 + (e & 1);
        s[n][e] = (col < p.Lk) ? s[n][e] * p.scale_log2 : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
      mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
    }
    // the 4 lanes of a quad share rows g and g+8
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
    const float al0 = exp2f(m_run[0] - mn0), al1 = exp2f(m_run[1] - mn1);
    m_run[0] = mn0;
    m_run[1] = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int n = 0; n < BN / 8; ++n) {
      s[n][0] = exp2f(s[n][0] - mn0);
      s[n][1] = exp2f(s[n][1] - mn0);
      s[n][2] = exp2f(s[n][2] - mn1);
      s[n][3] = exp2f(s[n][3] - mn1);
      rs0 += s[n][0] + s[n][1];
      rs1 += s[n][2] + s[n][3];
    }
    l_run[0] = l_run[0] * al0 + rs0;
    l_run[1] = l_run[1] * al1 + rs1;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      o_acc[n][0] *= al0; o_acc[n][1] *= al0;
      o_acc[n][2] *= al1; o_acc[n][3] *= al1;
    }
    if (drop.thr != 0) {
      const uint64_t r0 = (bh * p.Lq + (uint64_t)(q0 + warp * 16 + g)) * (uint64_t)p.Lkp;
      const uint64_t r1 = r0 + (uint64_t)8 * p.Lkp;
#pragma unroll
      for (int n = 0; n < BN / 8; ++n) {
        const uint4 b0 = dropout_bits8(drop, p.drop_stream, (r0 + j0 + n * 8) >> 3);
        const uint4 b1 = dropout_bits8(drop, p.drop_stream, (r1 + j0 + n * 8) >> 3);
        s[n][0] = dropout_keep(b0, 2 * t4, drop.thr) ? s[n][0] * drop.scale : 0.f;
        s[n][1] = dropout_keep(b0, 2 * t4 + 1, drop.thr) ? s[n][1] * drop.scale : 0.f;
        s[n][2] = dropout_keep(b1, 2 * t4, drop.thr) ? s[n][2] * drop.scale : 0.f;
        s[n][3] = dropout_keep(b1, 2 * t4 + 1, drop.thr) ? s[n][3] * drop.scale : 0.f;
      }
    }
    mma_frag_x_cols<HD, BN>(o_acc, s, Vs + (size_t)stage * BN * kLd, lane);
    __syncthreads();  // everyone is done with this stage before it is refilled
  }

  l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 1);
  l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 2);
  l_run[1] += __shfl_xor_sync(0xffffffffu, l_run[1], 1);
  l_run[1] += __shfl_xor_sync(0xffffffffu, l_run[1], 2);
  const int row0 = q0 + warp * 16 + g;
  if (t4 == 0) {
    if (row0 < p.Lq) p.lse2[bh * p.Lq + row0] = m_run[0] + log2f(l_run[0]);
    if (row0 + 8 < p.Lq) p.lse2[bh * p.Lq + row0 + 8] = m_run[1] + log2f(l_run[1]);
  }
  const int rows_valid = max(0, min(16, nq - warp * 16));
  store_rows_bf16<HD>(o_acc, 1.0f / l_run[0], 1.0f / l_run[1], Qs + (size_t)warp * 16 * kLd,
                      p.o + ((size_t)b * p.Lq + q0 + warp * 16) * p.ldo + (size_t)h * HD, p.ldo, rows_valid, lane);
}

// ================================================================================================
// decode forward: a short query block (<= 64 rows per CTA) against a long K/V, no dropout.
// This is the caption-decode shape (full_model.py:241-261 calls the bridge on a prefix of 1..64
// tokens against the 257 cached vision keys of each image): the work is reading K/V once, so the
// kernel is organised around keeping HBM requests in flight and enough warps resident to hide the
// latency of the (legacy-path) mma.sync chains.
//   * K/V are streamed in 16-key tiles through a 3/4-stage ring of TMA bulk copies with full/empty
//     mbarriers; a stage is refilled by the first warp of the group that consumed it one iteration
//     earlier. With the packed cache layout (b200b_kv_cache_pack) a tile is two 9 KB copies.
//   * a PAIR of warps owns 16 query rows: each computes Q K^T over half of the head dim, the two
//     partial score tiles are exchanged through shared memory (one named barrier per tile), both
//     run the same online softmax, and each accumulates P V for half of the output columns. That
//     halves the accumulator registers, so 16 warps (2 CTAs of 8) are resident per SM.
//   * with <= 32 query rows the spare pairs take every other key tile (flash-decoding inside the
//     CTA); the partial (max, sum, O) are merged through the ring's shared memory at the end.
//   * padding rows are duplicates of the last valid row (or zeros), so shared memory only ever
//     holds finite data; padded keys are masked to -inf.
// ================================================================================================
template <int HD>
struct DecodeCfg {
  static constexpr int kTile = 16;   // keys per stage
  static constexpr int kLd = HD + 8;
  static constexpr int kRowBytes = HD * 2;
  static constexpr int kStageElems = 2 * kTile * kLd;  // K rows, then V rows
  static constexpr int kXchgBytes = 4 * 2 * 2 * 8 * 32 * 4;  // [pair][tile parity][half][8 floats][lane]
  static int stages(int row_groups) { return row_groups >= 3 ? 3 : 4; }
  static size_t smem_bytes(int row_groups) {
    return ((size_t)row_groups * 16 * kLd + (size_t)stages(row_groups) * kStageElems) * 2 + kXchgBytes;
  }
};

__device__ __forceinline__ void pair_barrier(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

template <int HD, bool PACKED>
__global__ void __launch_bounds__(256, 2) attn_decode_kernel(const AttnParams p) {
  pdl_prologue();
  using Cfg = DecodeCfg<HD>;
  constexpr int kLd = Cfg::kLd, TK = Cfg::kTile;
  constexpr int KH = HD / 32;    // 16-wide k-steps of Q K^T per warp (half of the head dim)
  constexpr int NH = HD / 16;    // 8-column accumulator tiles per warp (half of the output columns)
  constexpr int kMaxStages = 4;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, half = warp & 1;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
  const int nq = min(64, p.Lq - q0);
  const int nrg = (nq + 15) >> 4;       // 16-row query groups: 1..4
  const int KS = (nrg <= 2) ? 2 : 1;    // key splits
  const int NS = (nrg >= 3) ? 3 : 4;    // ring stages (a multiple of KS)
  const bool active = pair < nrg * KS;
  const int rg = pair % nrg, ks = pair / nrg;
  const int nt = (p.Lk + TK - 1) / TK;

  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* ring = Qs + (size_t)nrg * 16 * kLd;
  float4* xchg = reinterpret_cast<float4*>(ring + (size_t)NS * Cfg::kStageElems) + (size_t)pair * (2 * 2 * 2 * 32);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (uint32_t)(2 * nrg));
    }
    fence_barrier_init();
  }
  if (PACKED && nt <= NS && (p.Lk % TK) != 0) {
    // the last tile is partial and its stage is never filled by an earlier (full) tile: the rows the
    // copy does not write must hold finite data (their probabilities are exactly 0)
    __nv_bfloat16* st = ring + (size_t)((nt - 1) % NS) * Cfg::kStageElems;
    const int r0 = p.Lk % TK, nz = (TK - r0) * kLd * 2 / 16;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) {
      reinterpret_cast<uint4*>(st + (size_t)r0 * kLd)[i] = make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(st + (size_t)(TK + r0) * kLd)[i] = make_uint4(0, 0, 0, 0);
    }
  }
  __syncthreads();

  // one warp stages key tile t
  auto issue_tile = [&](int t) {
    const int st = t % NS;
    if constexpr (PACKED) {
      // the cache holds, per (sample, head), Lk padded K rows then Lk padded V rows: a tile is two copies
      if (lane == 0) {
        const int nk = min(TK, p.Lk - t * TK);
        const uint32_t bytes = (uint32_t)(nk * kLd * 2);
        __nv_bfloat16* dst = ring + (size_t)st * Cfg::kStageElems;
        const __nv_bfloat16* src = p.k + (size_t)b * p.kv_bstride + (size_t)h * p.kv_hstride + (size_t)t * TK * kLd;
        mbar_arrive_expect_tx(&full_bar[st], 2 * bytes);
        bulk_load_1d(dst, src, bytes, &full_bar[st]);
        bulk_load_1d(dst + (size_t)TK * kLd, src + (size_t)p.Lk * kLd, bytes, &full_bar[st]);
      }
    } else {
      // row-major projections: lanes 0-15 copy the tile's K rows, lanes 16-31 its V rows
      __nv_bfloat16* dst = ring + (size_t)st * Cfg::kStageElems + (size_t)lane * kLd;
      const size_t grow = (size_t)b * p.Lk + min(t * TK + (lane & 15), p.Lk - 1);
      const __nv_bfloat16* src = (lane < 16) ? p.k + grow * p.ldk : p.v + grow * p.ldv;
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[st], (uint32_t)(2 * TK * Cfg::kRowBytes));
      __syncwarp();
      bulk_load_1d(dst, src + (size_t)h * HD, Cfg::kRowBytes, &full_bar[st]);
    }
  };
  const bool refiller = active && rg == 0 && half == 0;

  if (refiller)
    for (int t = ks; t < nt && t < NS; t += KS) issue_tile(t);
  {
    // Q block: plain vector loads by all threads (the K/V copies above are already in flight)
    constexpr int kChunks = HD / 8;
    for (int idx = threadIdx.x; idx < nrg * 16 * kChunks; idx += blockDim.x) {
      const int r = idx / kChunks, c = idx % kChunks;
      *reinterpret_cast<uint4*>(Qs + (size_t)r * kLd + c * 8) = __ldg(reinterpret_cast<const uint4*>(
          p.q + ((size_t)b * p.Lq + q0 + min(r, nq - 1)) * p.ldq + (size_t)h * HD + c * 8));
    }
  }
  __syncthreads();

  float o_acc[NH][4];
#pragma unroll
  for (int n = 0; n < NH; ++n) o_acc[n][0] = o_acc[n][1] = o_acc[n][2] = o_acc[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int g = lane >> 2, t4 = lane & 3;

  if (active) {
    // Q K^T: this warp's half of the head dim; P V: this warp's half of the output columns
    const uint32_t a_base = smem_u32(Qs + (size_t)(rg * 16 + (lane & 15)) * kLd + (lane >> 4) * 8 + half * (HD / 2));
    const uint32_t b_off =
        (uint32_t)((((lane & 7) + ((lane >> 4) << 3)) * kLd + ((lane >> 3) & 1) * 8 + half * (HD / 2)) * 2);
    const uint32_t v_off =
        (uint32_t)(((TK + (lane & 7) + ((lane >> 3) & 1) * 8) * kLd + (lane >> 4) * 8 + half * (HD / 2)) * 2);
    int it = 0;
    for (int t = ks; t < nt; t += KS, ++it) {
      const int st = t % NS;
      if (refiller && t >= KS && t - KS + NS < nt) {  // refill the stage this group released one iteration ago
        const int tp = t - KS;
        mbar_wait(&empty_bar[tp % NS], (uint32_t)((tp / NS) & 1));
        issue_tile(tp + NS);
      }
      mbar_wait(&full_bar[st], (uint32_t)((t / NS) & 1));
      const uint32_t st_base = smem_u32(ring + (size_t)st * Cfg::kStageElems);
#ifdef B200B_DIAG
      if (p.Lkp < 0) {  // diagnostics build only (B200B_DECODE_DEBUG=1): copy-only
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
        continue;
      }
#endif

      // partial S = Q[:, half] K[:, half]^T; two accumulator sets shorten the dependent-MMA chains
      float s[2][4], s2[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[n][e] = s2[n][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KH; ++kk) {
        uint32_t a[4], bb[4];
        ldmatrix_x4(a, a_base + kk * 32);
        ldmatrix_x4(bb, st_base + b_off + kk * 32);
        const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
        if (kk & 1) {
          mma_m16n8k16(s2[0], a, b0);
          mma_m16n8k16(s2[1], a, b1);
        } else {
          mma_m16n8k16(s[0], a, b0);
          mma_m16n8k16(s[1], a, b1);
        }
      }
      // exchange with the other warp of the pair (a + b == b + a exactly: both end with the same S)
      float4* mine = xchg + ((it & 1) * 2 + half) * 64;
      const float4* other = xchg + ((it & 1) * 2 + (half ^ 1)) * 64;
      mine[lane] = make_float4(s[0][0] + s2[0][0], s[0][1] + s2[0][1], s[0][2] + s2[0][2], s[0][3] + s2[0][3]);
      mine[32 + lane] = make_float4(s[1][0] + s2[1][0], s[1][1] + s2[1][1], s[1][2] + s2[1][2], s[1][3] + s2[1][3]);
      pair_barrier(pair);
      {
        const float4 o0 = other[lane], o1 = other[32 + lane];
        const float4 m0 = mine[lane], m1 = mine[32 + lane];
        s[0][0] = m0.x + o0.x; s[0][1] = m0.y + o0.y; s[0][2] = m0.z + o0.z; s[0][3] = m0.w + o0.w;
        s[1][0] = m1.x + o1.x; s[1][1] = m1.y + o1.y; s[1][2] = m1.z + o1.z; s[1][3] = m1.w + o1.w;
      }
      const int j0 = t * TK;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = j0 + n * 8 + 2 * t4 + (e & 1);
          s[n][e] = (col < p.Lk) ? s[n][e] * p.scale_log2 : -INFINITY;
        }
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);   // finite: every tile has a valid key
      const float al0 = exp2f(m_run[0] - mn0), al1 = exp2f(m_run[1] - mn1);
      m_run[0] = mn0;
      m_run[1] = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        s[n][0] = exp2f(s[n][0] - mn0);
        s[n][1] = exp2f(s[n][1] - mn0);
        s[n][2] = exp2f(s[n][2] - mn1);
        s[n][3] = exp2f(s[n][3] - mn1);
        rs0 += s[n][0] + s[n][1];
        rs1 += s[n][2] + s[n][3];
      }
      l_run[0] = l_run[0] * al0 + rs0;
      l_run[1] = l_run[1] * al1 + rs1;
      if (al0 != 1.0f || al1 != 1.0f) {   // the running maximum settles after a few tiles
#pragma unroll
        for (int n = 0; n < NH; ++n) {
          o_acc[n][0] *= al0; o_acc[n][1] *= al0;
          o_acc[n][2] *= al1; o_acc[n][3] *= al1;
        }
      }
      {
        uint32_t a[4];
        a[0] = pack_bf16(s[0][0], s[0][1]);
        a[1] = pack_bf16(s[0][2], s[0][3]);
        a[2] = pack_bf16(s[1][0], s[1][1]);
        a[3] = pack_bf16(s[1][2], s[1][3]);
#pragma unroll
        for (int dp = 0; dp < NH / 2; ++dp) {
          uint32_t bb[4];
          ldmatrix_x4_trans(bb, st_base + v_off + dp * 32);
          const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
          mma_m16n8k16(o_acc[2 * dp], a, b0);
          mma_m16n8k16(o_acc[2 * dp + 1], a, b1);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
  }

  // merge the key splits through the (now idle) ring
  __syncthreads();
  constexpr int kSlotFloats = (NH * 4 + 4) * 32;
  float* scratch = reinterpret_cast<float*>(ring) + (size_t)(rg * 2 + half) * kSlotFloats;
  if (active && KS == 2 && ks == 1) {
#pragma unroll
    for (int n = 0; n < NH; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) scratch[(n * 4 + e) * 32 + lane] = o_acc[n][e];
    scratch[(NH * 4 + 0) * 32 + lane] = m_run[0];
    scratch[(NH * 4 + 1) * 32 + lane] = m_run[1];
    scratch[(NH * 4 + 2) * 32 + lane] = l_run[0];
    scratch[(NH * 4 + 3) * 32 + lane] = l_run[1];
  }
  __syncthreads();
  if (!active || ks != 0) return;
  if (KS == 2) {
    const float mb0 = scratch[(NH * 4 + 0) * 32 + lane], mb1 = scratch[(NH * 4 + 1) * 32 + lane];
    const float lb0 = scratch[(NH * 4 + 2) * 32 + lane], lb1 = scratch[(NH * 4 + 3) * 32 + lane];
    const float mn0 = fmaxf(m_run[0], mb0), mn1 = fmaxf(m_run[1], mb1);
    const float fa0 = exp2f(m_run[0] - mn0), fa1 = exp2f(m_run[1] - mn1);
    const float fb0 = exp2f(mb0 - mn0), fb1 = exp2f(mb1 - mn1);   // 0 when the other split saw no key
    m_run[0] = mn0;
    m_run[1] = mn1;
    l_run[0] = l_run[0] * fa0 + lb0 * fb0;
    l_run[1] = l_run[1] * fa1 + lb1 * fb1;
#pragma unroll
    for (int n = 0; n < NH; ++n) {
      o_acc[n][0] = o_acc[n][0] * fa0 + scratch[(n * 4 + 0) * 32 + lane] * fb0;
      o_acc[n][1] = o_acc[n][1] * fa0 + scratch[(n * 4 + 1) * 32 + lane] * fb0;
      o_acc[n][2] = o_acc[n][2] * fa1 + scratch[(n * 4 + 2) * 32 + lane] * fb1;
      o_acc[n][3] = o_acc[n][3] * fa1 + scratch[(n * 4 + 3) * 32 + lane] * fb1;
    }
  }
  l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 1);
  l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 2);
  l_run[1] += __shfl_xor_sync(0xffffffffu, l_run[1], 1);
  l_run[1] += __shfl_xor_sync(0xffffffffu, l_run[1], 2);
  const int row0 = q0 + rg * 16 + g;
  if (half == 0 && t4 == 0 && p.lse2 != nullptr) {
    const size_t bh = (size_t)b * p.H + h;
    if (row0 < p.Lq) p.lse2[bh * p.Lq + row0] = m_run[0] + log2f(l_run[0]);
    if (row0 + 8 < p.Lq) p.lse2[bh * p.Lq + row0 + 8] = m_run[1] + log2f(l_run[1]);
  }
  // this warp's 16 x HD/2 block: bf16 through its columns of the (idle) Q rows, then 16-byte row stores
  {
    const float s0 = 1.0f / l_run[0], s1 = 1.0f / l_run[1];
    __nv_bfloat16* stage = Qs + (size_t)rg * 16 * kLd + half * (HD / 2);
#pragma unroll
    for (int n = 0; n < NH; ++n) {
      *reinterpret_cast<uint32_t*>(stage + (size_t)g * kLd + n * 8 + 2 * t4) = pack_bf16(o_acc[n][0] * s0, o_acc[n][1] * s0);
      *reinterpret_cast<uint32_t*>(stage + (size_t)(g + 8) * kLd + n * 8 + 2 * t4) =
          pack_bf16(o_acc[n][2] * s1, o_acc[n][3] * s1);
    }
    __syncwarp();
    const int rows_valid = max(0, min(16, nq - rg * 16));
    constexpr int kChunks = HD / 16;  // 16-byte chunks per half row
    __nv_bfloat16* gdst = p.o + ((size_t)b * p.Lq + q0 + rg * 16) * p.ldo + (size_t)h * HD + half * (HD / 2);
    for (int idx = lane; idx < 16 * kChunks; idx += 32) {
      const int r = idx / kChunks, c = idx % kChunks;
      if (r < rows_valid)
        *reinterpret_cast<uint4*>(gdst + (size_t)r * p.ldo + c * 8) = *reinterpret_cast<const uint4*>(stage + (size_t)r * kLd + c * 8);
    }
  }
}

// ================================================================================================
// backward: delta[b,h,i] = sum_c dO[i, h, c] * O[i, h, c]; one warp per (token, head)
// ================================================================================================
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo,
                                  const __nv_bfloat16* __restrict__ d_o, long long lddo, float* __restrict__ delta,
                                  int B, int H, int Lq, int HD) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = blockIdx.x;  // b * Lq + i
  if (warp >= H) return;
  const uint32_t* po = reinterpret_cast<const uint32_t*>(o + row * ldo + (size_t)warp * HD);
  const uint32_t* pd = reinterpret_cast<const uint32_t*>(d_o + row * lddo + (size_t)warp * HD);
  float s = 0.f;
  for (int c = lane; c < HD / 2; c += 32) {
    const uint32_t a = po[c], d = pd[c];
    s += bf16_lo(a) * bf16_lo(d) + bf16_hi(a) * bf16_hi(d);
  }
  s = warp_sum(s);
  if (lane == 0) {
    const long long b = row / Lq, i = row % Lq;
    delta[((size_t)b * H + warp) * Lq + i] = s;
  }
}

// ================================================================================================
// backward, query-major: dQ, plus bf16 spills of dropped P and of dS for the key-major pass
// ================================================================================================
template <int HD>
__global__ void __launch_bounds__(128) attn_bwd_q_kernel(const AttnParams p) {
  pdl_prologue();
  const DropoutCfg drop = dropout_resolve(p.drop);
  using Cfg = AttnCfg<HD>;
  constexpr int BN = Cfg::kBN, kLd = Cfg::kLd;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* dOs = Qs + (size_t)Cfg::kBM * kLd;
  __nv_bfloat16* Ks = dOs + (size_t)Cfg::kBM * kLd;
  __nv_bfloat16* Vs = Ks + (size_t)2 * BN * kLd;
  __shared__ __align__(8) uint64_t qbar, full_bar[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * Cfg::kBM;
  const int nq = min(Cfg::kBM, p.Lq - q0);
  const int nt = (p.Lk + BN - 1) / BN;

  if (threadIdx.x == 0) {
    mbar_init(&qbar, 1);
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) mbar_arrive_expect_tx(&qbar, (uint32_t)(2 * nq * Cfg::kRowBytes));
    __syncwarp();
    for (int r = lane; r < nq; r += 32) {
      const size_t grow = (size_t)b * p.Lq + q0 + r;
      bulk_load_1d(Qs + (size_t)r * kLd, p.q + grow * p.ldq + (size_t)h * HD, Cfg::kRowBytes, &qbar);
      bulk_load_1d(dOs + (size_t)r * kLd, p.d_o + grow * p.lddo + (size_t)h * HD, Cfg::kRowBytes, &qbar);
    }
    issue_kv_tile<HD>(p, b, h, 0, Ks, Vs, &full_bar[0], lane);
  }
  __syncthreads();

  float dq_acc[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) dq_acc[n][0] = dq_acc[n][1] = dq_acc[n][2] = dq_acc[n][3] = 0.f;
  const int g = lane >> 2, t4 = lane & 3;
  const size_t bh = (size_t)b * p.H + h;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  const bool ok0 = row0 < p.Lq, ok1 = row1 < p.Lq;
  const float lse0 = ok0 ? p.lse2[bh * p.Lq + row0] : 0.f, lse1 = ok1 ? p.lse2[bh * p.Lq + row1] : 0.f;
  const float del0 = ok0 ? p.delta[bh * p.Lq + row0] : 0.f, del1 = ok1 ? p.delta[bh * p.Lq + row1] : 0.f;
  const uint64_t ridx0 = (bh * p.Lq + (uint64_t)row0) * (uint64_t)p.Lkp;
  const uint64_t ridx1 = ridx0 + (uint64_t)8 * p.Lkp;

  mbar_wait(&qbar, 0);
  for (int t = 0; t < nt; ++t) {
    const int stage = t & 1;
    if (warp == 0 && t + 1 < nt)
      issue_kv_tile<HD>(p, b, h, (t + 1) * BN, Ks + (size_t)(stage ^ 1) * BN * kLd, Vs + (size_t)(stage ^ 1) * BN * kLd,
                        &full_bar[stage ^ 1], lane);
    mbar_wait(&full_bar[stage], (uint32_t)((t >> 1) & 1));

    float s[BN / 8][4], dp[BN / 8][4];
#pragma unroll
    for (int n = 0; n < BN / 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
    mma_rows_x_rows<HD, BN>(s, Qs + (size_t)warp * 16 * kLd, Ks + (size_t)stage * BN * kLd, lane);
    mma_rows_x_rows<HD, BN>(dp, dOs + (size_t)warp * 16 * kLd, Vs + (size_t)stage * BN * kLd, lane);

    const int j0 = t * BN;
#pragma unroll
    for (int n = 0; n < BN / 8; ++n) {
      uint4 b0 = make_uint4(0, 0, 0, 0), b1 = make_uint4(0, 0, 0, 0);
      if (drop.thr != 0) {
        b0 = dropout_bits8(drop, p.drop_stream, (ridx0 + j0 + n * 8) >> 3);
        b1 = dropout_bits8(drop, p.drop_stream, (ridx1 + j0 + n * 8) >> 3);
      }
      float pd[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = j0 + n * 8 + 2 * t4 + (e & 1);
        const float lse = (e < 2) ? lse0 : lse1, del = (e < 2) ? del0 : del1;
        float pr = (col < p.Lk) ? exp2f(s[n][e] * p.scale_log2 - lse) : 0.f;
        float dpr = dp[n][e];
        float prd = pr;
        if (drop.thr != 0) {
          const bool keep = dropout_keep((e < 2) ? b0 : b1, 2 * t4 + (e & 1), drop.thr);
          prd = keep ? pr * drop.scale : 0.f;
          dpr = keep ? dpr * drop.scale : 0.f;
        }
        pd[e] = prd;
        s[n][e] = pr * (dpr - del);  // dS (without the softmax scale)
      }
      const int colp = j0 + n * 8 + 2 * t4;
      if (colp < p.Lkp) {
        if (ok0) {
          *reinterpret_cast<uint32_t*>(p.p_scr + ridx0 + colp) = pack_bf16(pd[0], pd[1]);
          *reinterpret_cast<uint32_t*>(p.ds_scr + ridx0 + colp) = pack_bf16(s[n][0], s[n][1]);
        }
        if (ok1) {
          *reinterpret_cast<uint32_t*>(p.p_scr + ridx1 + colp) = pack_bf16(pd[2], pd[3]);
          *reinterpret_cast<uint32_t*>(p.ds_scr + ridx1 + colp) = pack_bf16(s[n][2], s[n][3]);
        }
      }
    }
    mma_frag_x_cols<HD, BN>(dq_acc, s, Ks + (size_t)stage * BN * kLd, lane);
    __syncthreads();
  }
  const int rows_valid = max(0, min(16, nq - warp * 16));
  store_rows_bf16<HD>(dq_acc, p.scale, p.scale, Qs + (size_t)warp * 16 * kLd,
                      p.dq + ((size_t)b * p.Lq + q0 + warp * 16) * p.lddq + (size_t)h * HD, p.lddq, rows_valid, lane);
}

// ================================================================================================
// backward, key-major: dV = Pd^T dO, dK = scale * dS^T Q. CTA = 32 keys; warps 0,1 -> dV of the
// two 16-key groups, warps 2,3 -> dK. The contraction runs over query rows in steps of 32.
// ================================================================================================
template <int HD>
__global__ void __launch_bounds__(128) attn_bwd_kv_kernel(const AttnParams p) {
  pdl_prologue();
  using Cfg = AttnCfg<HD>;
  constexpr int QT = Cfg::kQT, kLd = Cfg::kLd, SL = Cfg::kScrLd;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  // per stage: Ps[QT][SL], dSs[QT][SL], dOs[QT][kLd], Qs[QT][kLd]
  constexpr size_t kStageElems = (size_t)2 * QT * SL + (size_t)2 * QT * kLd;
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __shared__ __align__(8) uint64_t full_bar[2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * 32;
  const int nkeys = min(32, p.Lkp - j0);  // multiple of 8
  const int nt = (p.Lq + QT - 1) / QT;
  const size_t bh = (size_t)b * p.H + h;

  if (threadIdx.x == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int t) {  // warp 0
    const int stage = t & 1;
    __nv_bfloat16* Ps = base + (size_t)stage * kStageElems;
    __nv_bfloat16* dSs = Ps + (size_t)QT * SL;
    __nv_bfloat16* dOs = dSs + (size_t)QT * SL;
    __nv_bfloat16* Qs = dOs + (size_t)QT * kLd;
    const int i0 = t * QT;
    const int nr = min(QT, p.Lq - i0);
    for (int r = nr; r < QT; ++r) {  // rows past the sequence end are contraction terms: must be 0
      zero_row(Ps + (size_t)r * SL, 64, lane, 32);
      zero_row(dSs + (size_t)r * SL, 64, lane, 32);
      zero_row(dOs + (size_t)r * kLd, Cfg::kRowBytes, lane, 32);
      zero_row(Qs + (size_t)r * kLd, Cfg::kRowBytes, lane, 32);
    }
    if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(nr * (2 * nkeys * 2 + 2 * Cfg::kRowBytes)));
    __syncwarp();
    if (lane < nr) {
      const int r = lane;
      const size_t srow = (bh * p.Lq + (size_t)(i0 + r)) * (size_t)p.Lkp + j0;
      const size_t grow = (size_t)b * p.Lq + i0 + r;
      bulk_load_1d(Ps + (size_t)r * SL, p.p_scr + srow, (uint32_t)(nkeys * 2), &full_bar[stage]);
      bulk_load_1d(dSs + (size_t)r * SL, p.ds_scr + srow, (uint32_t)(nkeys * 2), &full_bar[stage]);
      bulk_load_1d(dOs + (size_t)r * kLd, p.d_o + grow * p.lddo + (size_t)h * HD, Cfg::kRowBytes, &full_bar[stage]);
      bulk_load_1d(Qs + (size_t)r * kLd, p.q + grow * p.ldq + (size_t)h * HD, Cfg::kRowBytes, &full_bar[stage]);
    }
  };

  if (warp == 0) issue(0);
  __syncthreads();

  float acc[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  const int kg = warp & 1, which = warp >> 1;  // which: 0 = dV, 1 = dK

  for (int t = 0; t < nt; ++t) {
    const int stage = t & 1;
    if (warp == 0 && t + 1 < nt) issue(t + 1);
    mbar_wait(&full_bar[stage], (uint32_t)((t >> 1) & 1));
    const __nv_bfloat16* Ps = base + (size_t)stage * kStageElems;
    const __nv_bfloat16* dSs = Ps + (size_t)QT * SL;
    const __nv_bfloat16* dOs = dSs + (size_t)QT * SL;
    const __nv_bfloat16* Qs = dOs + (size_t)QT * kLd;
    const __nv_bfloat16* A = which ? dSs : Ps;    // [q rows][keys]  -> A^T via ldmatrix.trans
    const __nv_bfloat16* Bm = which ? Qs : dOs;   // [q rows][HD]
    const uint32_t a_base = smem_u32(A + (size_t)((lane & 7) + (lane >> 4) * 8) * SL + kg * 16 + ((lane >> 3) & 1) * 8);
    const uint32_t b_base = smem_u32(Bm + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * kLd + (lane >> 4) * 8);
#pragma unroll
    for (int kk = 0; kk < QT / 16; ++kk) {
      uint32_t a[4];
      ldmatrix_x4_trans(a, a_base + (uint32_t)(kk * 16 * SL * 2));
#pragma unroll
      for (int dpi = 0; dpi < HD / 16; ++dpi) {
        uint32_t bb[4];
        ldmatrix_x4_trans(bb, b_base + (uint32_t)(kk * 16 * kLd * 2) + dpi * 32);
        const uint32_t b0[2] = {bb[0], bb[1]}, b1[2] = {bb[2], bb[3]};
        mma_m16n8k16(acc[2 * dpi], a, b0);
        mma_m16n8k16(acc[2 * dpi + 1], a, b1);
      }
    }
    __syncthreads();
  }

  const int g = lane >> 2, t4 = lane & 3;
  const float sc = which ? p.scale : 1.0f;
  __nv_bfloat16* dst = which ? p.dk : p.dv;
  const long long ldd = which ? p.lddk : p.lddv;
  const int key0 = j0 + kg * 16 + g, key1 = key0 + 8;
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    const int c = n * 8 + 2 * t4;
    if (key0 < p.Lk)
      *reinterpret_cast<uint32_t*>(dst + ((size_t)b * p.Lk + key0) * ldd + (size_t)h * HD + c) =
          pack_bf16(acc[n][0] * sc, acc[n][1] * sc);
    if (key1 < p.Lk)
      *reinterpret_cast<uint32_t*>(dst + ((size_t)b * p.Lk + key1) * ldd + (size_t)h * HD + c) =
          pack_bf16(acc[n][2] * sc, acc[n][3] * sc);
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
template <typename K>
static int set_smem(K kern, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_last_error("%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return B200B_OK;
}

template <int HD>
static int launch_fwd(const AttnParams& p, cudaStream_t stream) {
  using Cfg = AttnCfg<HD>;
  int rc = set_smem(attn_fwd_kernel<HD>, Cfg::kFwdSmem, "attn_fwd");
  if (rc) return rc;
  dim3 grid((p.Lq + Cfg::kBM - 1) / Cfg::kBM, p.H, p.B);
  launch_pdl(kPdlAttn, attn_fwd_kernel<HD>, grid, dim3(128), Cfg::kFwdSmem, stream, p);
  return check_launch("attn_fwd", stream);
}

// Short query blocks without dropout (caption decode, and any inference call on <= 64 tokens) take
// the K/V-streaming kernel.
static bool use_decode_kernel(const AttnParams& p) { return p.drop.thr == 0 && p.Lq <= 64; }

template <int HD, bool PACKED>
static int launch_decode(const AttnParams& p, cudaStream_t stream) {
  using Cfg = DecodeCfg<HD>;
  static bool attr_done = false;  // the maximum is the same for every launch; a race only repeats the call
  if (!attr_done) {
    int rc = set_smem(attn_decode_kernel<HD, PACKED>, Cfg::smem_bytes(2), "attn_decode");
    if (rc) return rc;
    attr_done = true;
  }
  const int nrg = (min(64, p.Lq) + 15) / 16;
  dim3 grid((p.Lq + 63) / 64, p.H, p.B);
  AttnParams pp = p;
#ifdef B200B_DIAG
  static const int dbg = [] { const char* e = getenv("B200B_DECODE_DEBUG"); return e ? atoi(e) : 0; }();
  if (dbg & 1) pp.Lkp = -1;
#endif
  launch_pdl(kPdlAttn, attn_decode_kernel<HD, PACKED>, grid, dim3(256), Cfg::smem_bytes(nrg), stream, pp);
  return check_launch(PACKED ? "attn_decode_packed" : "attn_decode", stream);
}

template <int HD>
static int launch_bwd(const AttnParams& p, cudaStream_t stream, const b200b_attn_args* a) {
  using Cfg = AttnCfg<HD>;
  int rc = set_smem(attn_bwd_q_kernel<HD>, Cfg::kBwdQSmem, "attn_bwd_q");
  if (rc) return rc;
  rc = set_smem(attn_bwd_kv_kernel<HD>, Cfg::kBwdKVSmem, "attn_bwd_kv");
  if (rc) return rc;
  const int delta_threads = 32 * p.H;
  launch_pdl(kPdlAttn, attn_delta_kernel, dim3(p.B * p.Lq), dim3(delta_threads), 0, stream, p.o, p.ldo, p.d_o, p.lddo, p.delta, p.B, p.H, p.Lq,
             HD);
  rc = check_launch("attn_delta", stream);
  if (rc) return rc;
  {
    int taken = 0;
    rc = attention_bwd_tc(a, p.delta, stream, &taken);
    if (taken) return rc;
  }
  dim3 gq((p.Lq + Cfg::kBM - 1) / Cfg::kBM, p.H, p.B);
  launch_pdl(kPdlAttn, attn_bwd_q_kernel<HD>, gq, dim3(128), Cfg::kBwdQSmem, stream, p);
  rc = check_launch("attn_bwd_q", stream);
  if (rc) return rc;
  dim3 gk((p.Lk + 31) / 32, p.H, p.B);
  launch_pdl(kPdlAttn, attn_bwd_kv_kernel<HD>, gk, dim3(128), Cfg::kBwdKVSmem, stream, p);
  return check_launch("attn_bwd_kv", stream);
}

static bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int validate_common(const b200b_attn_args* a, const char* what) {
  if (a == nullptr || !a->q || !a->k || !a->v || !a->o || !a->lse) {
    set_last_error("%s: null argument", what);
    return B200B_ERR_ARG;
  }
  if (a->batch <= 0 || a->heads <= 0 || a->heads > 32 || a->len_q <= 0 || a->len_k <= 0) {
    set_last_error("%s: bad sizes (batch=%d heads=%d len_q=%d len_k=%d)", what, a->batch, a->heads, a->len_q, a->len_k);
    return B200B_ERR_SHAPE;
  }
  if (a->head_dim != 64 && a->head_dim != 128 && a->head_dim != 288) {
    set_last_error("%s: head_dim %d not built (64, 128, 288)", what, a->head_dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16p(a->q) || !al16p(a->k) || !al16p(a->v) || !al16p(a->o) || (a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) ||
      (a->ldo % 8)) {
    set_last_error("%s: q/k/v/o must be 16-byte aligned with row pitch multiple of 8 elements", what);
    return B200B_ERR_ALIGN;
  }
  if (!(a->dropout_p >= 0.0f && a->dropout_p < 1.0f)) {
    set_last_error("%s: dropout_p must be in [0,1)", what);
    return B200B_ERR_ARG;
  }
  return B200B_OK;
}

static AttnParams make_params(const b200b_attn_args* a) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q); p.ldq = a->ldq;
  p.k = reinterpret_cast<const __nv_bfloat16*>(a->k); p.ldk = a->ldk;
  p.v = reinterpret_cast<const __nv_bfloat16*>(a->v); p.ldv = a->ldv;
  p.o = reinterpret_cast<__nv_bfloat16*>(a->o); p.ldo = a->ldo;
  p.lse2 = a->lse;
  p.B = a->batch; p.H = a->heads; p.Lq = a->len_q; p.Lk = a->len_k;
  p.Lkp = (a->len_k + 7) & ~7;
  p.scale = 1.0f / sqrtf((float)a->head_dim);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.drop_stream = a->dropout_stream;
  p.drop = make_dropout_cfg(a->dropout_p, a->seed, &p.drop_stream);
  return p;
}

}  // namespace b200b

using namespace b200b;

extern "C" int b200b_attention_fwd(const b200b_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_common(a, "attention_fwd");
  if (rc) return rc;
  AttnParams p = make_params(a);
  if (use_decode_kernel(p)) {
    switch (a->head_dim) {
      case 64: return launch_decode<64, false>(p, stream);
      case 128: return launch_decode<128, false>(p, stream);
      default: return launch_decode<288, false>(p, stream);
    }
  }
  {
    int taken = 0;
    rc = attention_fwd_tc(a, stream, &taken);
    if (taken) return rc;
  }
  switch (a->head_dim) {
    case 64: return launch_fwd<64>(p, stream);
    case 128: return launch_fwd<128>(p, stream);
    default: return launch_fwd<288>(p, stream);
  }
}

extern "C" size_t b200b_attention_bwd_workspace_bytes(int batch, int heads, int len_q, int len_k) {
  const size_t lkp = ((size_t)len_k + 7) & ~(size_t)7;
  const size_t rows = (size_t)batch * heads * len_q;
  // delta (fp32, padded to 16 B) + two bf16 scratch matrices
  return ((rows * 4 + 255) & ~(size_t)255) + 2 * ((rows * lkp * 2 + 255) & ~(size_t)255);
}

extern "C" int b200b_attention_bwd(const b200b_attn_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = validate_common(a, "attention_bwd");
  if (rc) return rc;
  if (!a->d_o || !a->dq || !a->dk || !a->dv || !a->workspace) {
    set_last_error("attention_bwd: null gradient / workspace pointer");
    return B200B_ERR_ARG;
  }
  if (!al16p(a->d_o) || !al16p(a->dq) || !al16p(a->dk) || !al16p(a->dv) || !al16p(a->workspace) || (a->lddo % 8) ||
      (a->lddq % 8) || (a->lddk % 8) || (a->lddv % 8)) {
    set_last_error("attention_bwd: gradients must be 16-byte aligned with row pitch multiple of 8 elements");
    return B200B_ERR_ALIGN;
  }
  const size_t need = b200b_attention_bwd_workspace_bytes(a->batch, a->heads, a->len_q, a->len_k);
  if (a->workspace_bytes < need) {
    set_last_error("attention_bwd: workspace too small (%zu < %zu)", (size_t)a->workspace_bytes, need);
    return B200B_ERR_WORKSPACE;
  }
  AttnParams p = make_params(a);
  p.d_o = reinterpret_cast<const __nv_bfloat16*>(a->d_o); p.lddo = a->lddo;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.lddq = a->lddq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.lddk = a->lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.lddv = a->lddv;
  const size_t rows = (size_t)a->batch * a->heads * a->len_q;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  p.delta = reinterpret_cast<float*>(ws);
  const size_t off1 = (rows * 4 + 255) & ~(size_t)255;
  const size_t scr = (rows * (size_t)p.Lkp * 2 + 255) & ~(size_t)255;
  p.p_scr = reinterpret_cast<__nv_bfloat16*>(ws + off1);
  p.ds_scr = reinterpret_cast<__nv_bfloat16*>(ws + off1 + scr);
  switch (a->head_dim) {
    case 64: return launch_bwd<64>(p, stream, a);
    case 128: return launch_bwd<128>(p, stream, a);
    default: return launch_bwd<288>(p, stream, a);
  }
}

// ------------------------------------------------------------------------------------------------
// packed K/V cache for decode
// ------------------------------------------------------------------------------------------------
namespace b200b {

// kv [B*Lk, nb*2*H*HD] (block i: K columns then V columns) -> packed [B][nb][H][2][Lk][HD+8]: per
// (sample, block, head) the K rows and then the V rows, each padded to the shared-memory row pitch of
// the decode kernel, so that a 16-key tile is ONE contiguous copy that lands bank-conflict free.
__global__ void kv_cache_pack_kernel(const __nv_bfloat16* __restrict__ kv, long long ldkv,
                                     __nv_bfloat16* __restrict__ packed, int B, int Lk, int H, int HD, int nb) {
  pdl_prologue();
  const int chunks = (HD + 8) / 8;  // 16-byte chunks per padded row
  const long long total = (long long)B * nb * H * 2 * Lk * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % chunks);
    long long r = i / chunks;
    const int j = (int)(r % Lk); r /= Lk;
    const int which = (int)(r % 2); r /= 2;
    const int h = (int)(r % H); r /= H;
    const int blk = (int)(r % nb);
    const int b = (int)(r / nb);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c * 8 < HD)
      v = __ldg(reinterpret_cast<const uint4*>(kv + ((long long)b * Lk + j) * ldkv + (long long)(2 * blk + which) * H * HD +
                                               (long long)h * HD + c * 8));
    reinterpret_cast<uint4*>(packed)[i] = v;
  }
}

}  // namespace b200b

extern "C" size_t b200b_kv_cache_packed_bytes(int batch, int len_k, int heads, int head_dim, int num_blocks) {
  return (size_t)batch * num_blocks * heads * 2 * len_k * (head_dim + 8) * 2;
}

extern "C" int b200b_kv_cache_pack(const void* kv, int64_t ldkv, void* packed, int batch, int len_k, int heads,
                                   int head_dim, int num_blocks, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!kv || !packed || batch <= 0 || len_k <= 0 || heads <= 0 || num_blocks <= 0 || head_dim <= 0 || (head_dim % 8) ||
      (ldkv % 8) || !al16p(kv) || !al16p(packed)) {
    set_last_error("kv_cache_pack: bad argument (need 16-byte aligned pointers, head_dim and ldkv multiples of 8)");
    return B200B_ERR_ARG;
  }
  const long long total = (long long)batch * num_blocks * heads * 2 * len_k * ((head_dim + 8) / 8);
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < 148 * 16 ? (total + threads - 1) / threads : 148 * 16);
  launch_pdl(kPdlAttn, kv_cache_pack_kernel, dim3(blocks), dim3(threads), 0, stream, reinterpret_cast<const __nv_bfloat16*>(kv),
             (long long)ldkv, reinterpret_cast<__nv_bfloat16*>(packed), batch, len_k, heads, head_dim, num_blocks);
  return check_launch("kv_cache_pack", stream);
}

extern "C" int b200b_attention_decode_packed(const void* q, int64_t ldq, const void* kv_packed, int block_index,
                                             int num_blocks, void* o, int64_t ldo, float* lse, int batch, int heads,
                                             int len_q, int len_k, int head_dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !kv_packed || !o || batch <= 0 || heads <= 0 || len_q <= 0 || len_q > 64 || len_k <= 0 ||
      block_index < 0 || block_index >= num_blocks) {
    set_last_error("attention_decode_packed: bad argument (need 1 <= len_q <= 64, 0 <= block_index < num_blocks)");
    return B200B_ERR_ARG;
  }
  if (head_dim != 64 && head_dim != 128 && head_dim != 288) {
    set_last_error("attention_decode_packed: head_dim %d not built (64, 128, 288)", head_dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16p(q) || !al16p(kv_packed) || !al16p(o) || (ldq % 8) || (ldo % 8)) {
    set_last_error("attention_decode_packed: q/kv/o must be 16-byte aligned with row pitch multiple of 8 elements");
    return B200B_ERR_ALIGN;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const long long per_head = 2LL * len_k * (head_dim + 8);
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.ldq = ldq;
  p.k = reinterpret_cast<const __nv_bfloat16*>(kv_packed) + (long long)block_index * heads * per_head;
  p.kv_hstride = per_head;
  p.kv_bstride = (long long)num_blocks * heads * per_head;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = ldo;
  p.lse2 = lse;
  p.B = batch; p.H = heads; p.Lq = len_q; p.Lk = len_k; p.Lkp = (len_k + 7) & ~7;
  p.scale = 1.0f / sqrtf((float)head_dim);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.drop = make_dropout_cfg(0.f, 0, nullptr);
  switch (head_dim) {
    case 64: return launch_decode<64, true>(p, stream);
    case 128: return launch_decode<128, true>(p, stream);
    default: return launch_decode<288, true>(p, stream);
  }
}

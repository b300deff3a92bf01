// Decode cross-attention on the tcgen05 tensor cores (M = 64 query rows per CTA).
//
// Why: the mma.sync decode kernel (attention.cu) is bound by the legacy HMMA pipe (~144 TF/s on B200)
// as soon as a step has more than 16 query rows; tcgen05.mma has ~16x that rate, which makes the
// decode cross-attention (full_model.py:241-261 -> bridge_module.py:122-139 on the cached vision K/V)
// memory-bound for every prefix length up to 64.
//
// Layout facts this kernel rests on were measured with a one-CTA probe on B200 (now tests/gpu_checks/probe_umma_layouts.cu; profiles/r01_probe_tcgen05.jsonl):
//   * 32-byte-swizzle K-major operands (descriptor layout code 6) work for any K that is a multiple of
//     16 -- head dim 288 = 18 chunks of 16 -- with SBO = 256 B (8 rows x 32 B) and one k-step per chunk;
//   * an M = 64 accumulator (cta_group::1) keeps row r in TMEM lane 32*(r/16) + r%16, i.e. warp w finds
//     rows 16w..16w+15 in its lanes 0..15 (lanes 16..31 of every warp are unused).
//
// Cache layout ("tc" packing, b200b_kv_cache_pack_tc): per (image, block, head) a sequence of key tiles
// (32 keys; the last one padded to a multiple of 16 with zeros), each stored as the exact shared-memory
// image the MMAs read -- K tile as [18 chunks][keys][32 B] (B operand of S = Q K^T, keys are rows) followed
// by V^T tile as [keys/16 chunks][288][32 B] (B operand of O = P V, head-dim are rows, 16 keys per chunk),
// both 32-byte swizzled -- so a tile is ONE contiguous TMA bulk copy of keys x 1152 bytes.
//
// Warp-specialised: a control warp issues the TMA copies and, per 32-key tile, the 18 S MMAs (into one of
// two S buffers in TMEM, one tile ahead of the softmax) and the P V MMAs (two N = 144 halves per 16-key
// chunk, accumulating O in TMEM); four softmax warps read their rows of S from TMEM (one thread = one
// query row, so the softmax needs no shuffles), update the running max / sum, write P as bf16 into one of
// two A-operand images and rescale O in TMEM only if a row maximum jumped by more than 2^8. All hand-offs
// are mbarriers (tcgen05.commit on the MMA side). Four 36 KB stages: three tile copies are in flight
// while one tile is consumed.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "launch.h"

namespace b200b {

constexpr int kTcTile = 32;   // keys per tile
constexpr int kTcStages = 4;  // tiles resident in shared memory (3 in flight while one is consumed)
constexpr int kTcSBufs = 3;   // S accumulators in TMEM: the S MMAs run up to 3 tiles ahead of the softmax

__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(16u >> 4) << 16;    // LBO (unused for swizzled K-major)
  d |= (uint64_t)(256u >> 4) << 32;   // SBO: 8 rows x 32 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;             // SWIZZLE_32B
  return d;
}

// byte offset of element (row, k) in a [rows x K] K-major operand image with 32-byte swizzle:
// [K/16 chunks][rows][32 B], the two 16-byte units of a row swapped in rows 4..7 of every 8-row atom
__host__ __device__ __forceinline__ uint32_t sw32_offset(int row, int k, int rows) {
  return (uint32_t)(k >> 4) * (uint32_t)rows * 32u + (uint32_t)row * 32u +
         ((((uint32_t)(k >> 3) & 1u) ^ (((uint32_t)row >> 2) & 1u)) << 4) + ((uint32_t)k & 7u) * 2u;
}

struct TcParams {
  const __nv_bfloat16* q; long long ldq;
  const uint8_t* kv;            // this block's first (image 0, head 0) tile; see strides
  long long bstride, hstride;   // bytes between images / heads
  __nv_bfloat16* o; long long ldo;
  float* lse2;                  // [B, H, Lq] or null
  int B, H, Lq, Lk;
  float scale_log2;
  int debug_copy_only;          // -DB200B_DIAG builds only: stream the tiles, skip MMAs and softmax
};

template <int HD>
struct TcCfg {
  static constexpr int kChunks = HD / 16;
  static constexpr int kQBytes = 64 * HD * 2;
  static constexpr int kPBytes = 64 * kTcTile * 2;      // one P buffer (two are kept)
  static constexpr int kStageBytes = kTcTile * HD * 4;  // K image + V^T image of a full tile
  static constexpr int kNHalves = HD > 256 ? 2 : 1;
  static constexpr int kNHalf = HD / kNHalves;
  static constexpr int kOCol = kTcSBufs * kTcTile;      // TMEM: the S buffers, then O
  static constexpr uint32_t kTmemCols = (kOCol + HD) <= 128 ? 128 : ((kOCol + HD) <= 256 ? 256 : 512);
  static constexpr size_t kSmemBytes = kQBytes + 2 * kPBytes + (size_t)kTcStages * kStageBytes + 1024;
  static_assert(HD % 16 == 0 && kNHalf % 16 == 0 && kNHalf <= 256, "head dim");
};

static_assert(kTcSBufs < kTcStages, "S look-ahead must stay inside the resident K/V tiles");
constexpr int kTcThreads = 160;  // warps 0-3: softmax (one thread per query row in lanes 0-15), warp 4: TMA + MMA issue

template <int HD>
__global__ void __launch_bounds__(kTcThreads, 1) attn_decode_tc_kernel(const TcParams p) {
  using Cfg = TcCfg<HD>;
  extern __shared__ uint8_t smem_tc_raw[];
  // full: a K/V tile landed; s_full / s_free: an S buffer was produced / has been read;
  // p_full / p_free: a P buffer was written (and O rescaled if needed) / its P V has completed
  __shared__ __align__(8) uint64_t full_bar[kTcStages], s_full[kTcSBufs], s_free[kTcSBufs], p_full[2], p_free[2];
  __shared__ uint32_t tmem_base_smem;
  pdl_launch_dependents();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, h = blockIdx.x;
  const uint32_t s0 = (smem_u32(smem_tc_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_tc_raw + (s0 - smem_u32(smem_tc_raw));
  uint8_t* q_img = base;
  uint8_t* p_img = base + Cfg::kQBytes;                 // two buffers
  uint8_t* stage0 = p_img + 2 * Cfg::kPBytes;
  const int nq = min(64, p.Lq);
  const int nt = (p.Lk + kTcTile - 1) / kTcTile;
  const int rows_last = ((p.Lk - (nt - 1) * kTcTile) + 15) & ~15;
  const uint8_t* src = p.kv + (size_t)b * p.bstride + (size_t)h * p.hstride;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kTcStages; ++i) mbar_init(&full_bar[i], 1);
#pragma unroll
    for (int i = 0; i < kTcSBufs; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_wait();
  __syncthreads();   // barriers initialised before the first copy is credited to them
  auto issue_tile = [&](int t) {   // one thread
    const int rows = (t == nt - 1) ? rows_last : kTcTile;
    const uint32_t bytes = (uint32_t)rows * HD * 4;
    const int st = t % kTcStages;
    mbar_arrive_expect_tx(&full_bar[st], bytes);
    bulk_load_1d(stage0 + (size_t)st * Cfg::kStageBytes, src + (size_t)t * kTcTile * HD * 4, bytes, &full_bar[st]);
  };
  // The Q block's loads are issued first (all of a thread's loads in flight together), then the K/V
  // copies -- behind 147 KB of tile traffic per SM the small Q loads would wait microseconds.
  {
    // Q block -> A-operand image (rows past the block are zero)
    constexpr int kUnits = 64 * (HD / 8);
    constexpr int kPerThread = (kUnits + kTcThreads - 1) / kTcThreads;
    uint4 qv[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int idx = threadIdx.x + i * kTcThreads;
      const int r = idx / (HD / 8), u = idx % (HD / 8);
      qv[i] = make_uint4(0, 0, 0, 0);
      if (idx < kUnits && r < nq)
        qv[i] = __ldg(reinterpret_cast<const uint4*>(p.q + ((size_t)b * p.Lq + r) * p.ldq + (size_t)h * HD + u * 8));
    }
    if (threadIdx.x == 4 * 32)
      for (int t = 0; t < kTcStages && t < nt; ++t) issue_tile(t);
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int idx = threadIdx.x + i * kTcThreads;
      if (idx < kUnits) *reinterpret_cast<uint4*>(q_img + sw32_offset(idx / (HD / 8), (idx % (HD / 8)) * 8, 64)) = qv[i];
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_o = tmem_base + Cfg::kOCol;

#ifdef B200B_DIAG
  if (p.debug_copy_only) {
    if (warp == 4) {
      for (int t = 0; t < nt; ++t) {
        mbar_wait(&full_bar[t % kTcStages], (uint32_t)((t / kTcStages) & 1));
        if (t + kTcStages < nt && elect_one()) issue_tile(t + kTcStages);
        __syncwarp();
      }
    }
  } else
#endif
  if (warp == 4) {
    // ---------------------------- control warp: MMA issue and K/V refills ----------------------------
    const uint64_t q_desc0 = umma_desc_sw32(smem_u32(q_img));
    const uint64_t p_desc0 = umma_desc_sw32(smem_u32(p_img));
    const uint64_t st_desc0 = umma_desc_sw32(smem_u32(stage0));
    auto issue_s = [&](int t) {   // S(t) = Q K_t^T into S buffer t % kTcSBufs (warp-uniform; one lane issues)
      const int rows = (t == nt - 1) ? rows_last : kTcTile;
      const int st = t % kTcStages;
      const int sb = t % kTcSBufs;
      if (t >= kTcSBufs) mbar_wait(&s_free[sb], (uint32_t)((t / kTcSBufs - 1) & 1));
      mbar_wait(&full_bar[st], (uint32_t)((t / kTcStages) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_bf16(64, rows, false, false);
        const uint64_t kd = st_desc0 + (uint64_t)(((uint32_t)st * Cfg::kStageBytes) >> 4);
        const uint32_t kstep = (uint32_t)(rows * 32) >> 4;
        const uint32_t d = tmem_base + (uint32_t)(sb * kTcTile);
#pragma unroll
        for (int c = 0; c < Cfg::kChunks; ++c)
          umma_bf16(d, q_desc0 + (uint64_t)(c * ((64 * 32) >> 4)), kd + (uint64_t)(c * kstep), idesc, c > 0);
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
    };
    for (int u = 0; u < kTcSBufs && u < nt; ++u) issue_s(u);
    for (int t = 0; t < nt; ++t) {
      const int rows = (t == nt - 1) ? rows_last : kTcTile;
      const int st = t % kTcStages;
      mbar_wait(&p_full[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_bf16(64, Cfg::kNHalf, false, false);
        const uint64_t pd = p_desc0 + (uint64_t)(((uint32_t)(t & 1) * Cfg::kPBytes) >> 4);
        const uint64_t vd = st_desc0 + (uint64_t)(((uint32_t)st * Cfg::kStageBytes + (uint32_t)rows * HD * 2) >> 4);
#pragma unroll
        for (int kc = 0; kc < kTcTile / 16; ++kc)
          if (kc < rows / 16) {
#pragma unroll
            for (int nh = 0; nh < Cfg::kNHalves; ++nh)
              umma_bf16(tmem_o + (uint32_t)(nh * Cfg::kNHalf), pd + (uint64_t)(kc * ((64 * 32) >> 4)),
                        vd + (uint64_t)((kc * HD * 32 + nh * Cfg::kNHalf * 32) >> 4), idesc, (t > 0 || kc > 0));
          }
        umma_commit(&p_free[t & 1]);
      }
      __syncwarp();
      // the stage of tile t is free once its P V has completed: refill it kTcStages tiles ahead
      if (t + kTcStages < nt) {
        mbar_wait(&p_free[t & 1], (uint32_t)((t >> 1) & 1));
        if (elect_one()) issue_tile(t + kTcStages);
        __syncwarp();
      }
      // keep the S MMAs kTcSBufs tiles ahead: tile t + kTcSBufs is already resident (kTcSBufs < kTcStages)
      // and its S buffer is the one the softmax of tile t has just finished reading
      if (t + kTcSBufs < nt) issue_s(t + kTcSBufs);
    }
  } else {
    // ---------------------------- softmax warps: one thread per query row ----------------------------
    // The running maximum used for the exponentials follows the true maximum only when that moves by
    // more than 2^8, so the O accumulator in TMEM is almost never rescaled; p <= 256 is exact enough in
    // bf16 and the final division by l_run (same reference point) is exact.
    float m_run = -INFINITY, l_run = 0.f;
    const bool owner = lane < 16;   // this thread's TMEM lane holds query row warp*16 + lane
    const int row = warp * 16 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int t = 0; t < nt; ++t) {
      const int rows = (t == nt - 1) ? rows_last : kTcTile;
      const int sb = t % kTcSBufs;
      mbar_wait(&s_full[sb], (uint32_t)((t / kTcSBufs) & 1));
      tc_fence_after();
      float sv[kTcTile];
      float tmax = -INFINITY;
      {
        // all of the tile's 16-column loads are issued before the single tcgen05.wait::ld
        uint32_t v[kTcTile / 16][16];
#pragma unroll
        for (int g = 0; g < kTcTile / 16; ++g)
          if (g * 16 < rows) tmem_ld_32x32_x16(tmem_base + lane_off + (uint32_t)(sb * kTcTile + g * 16), v[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < kTcTile / 16; ++g)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = (g * 16 < rows && t * kTcTile + g * 16 + j < p.Lk) ? __uint_as_float(v[g][j]) * p.scale_log2
                                                                                : -INFINITY;
            sv[g * 16 + j] = x;
            tmax = fmaxf(tmax, x);
          }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sb]);   // the S buffer may be overwritten by S(t + kTcSBufs)
      float alpha = 1.0f;
      if (tmax > m_run + 8.0f) {
        alpha = exp2f(m_run - tmax);   // 0 on the first tile
        m_run = tmax;
      }
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < kTcTile; ++j) {
        sv[j] = exp2f(sv[j] - m_run);
        rs += sv[j];
      }
      l_run = l_run * alpha + rs;
      // P buffer t&1 was last read by P V of tile t-2
      if (t >= 2) mbar_wait(&p_free[t & 1], (uint32_t)(((t >> 1) - 1) & 1));
      if (owner) {
        uint8_t* pb = p_img + (size_t)(t & 1) * Cfg::kPBytes;
#pragma unroll
        for (int j = 0; j < kTcTile; j += 8) {
          if (j < rows) {
            uint4 u;
            u.x = pack_bf16(sv[j], sv[j + 1]);
            u.y = pack_bf16(sv[j + 2], sv[j + 3]);
            u.z = pack_bf16(sv[j + 4], sv[j + 5]);
            u.w = pack_bf16(sv[j + 6], sv[j + 7]);
            *reinterpret_cast<uint4*>(pb + sw32_offset(row, j, 64)) = u;
          }
        }
      }
      if (t >= 1) {
        const unsigned moved = __ballot_sync(0xffffffffu, owner && alpha != 1.0f);
        if (moved) {   // rare: a row maximum of this warp jumped; rescale its 16 rows of O after P V(t-1)
          mbar_wait(&p_free[(t - 1) & 1], (uint32_t)(((t - 1) >> 1) & 1));
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < HD; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32_x16(tmem_o + lane_off + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
            tmem_st_32x32_x16(tmem_o + lane_off + (uint32_t)c0, v);
          }
          tmem_st_wait();
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t & 1]);
    }
    mbar_wait(&p_free[(nt - 1) & 1], (uint32_t)(((nt - 1) >> 1) & 1));
    tc_fence_after();
    // tcgen05.ld is warp-collective: every lane executes it, only lanes that own a valid row store
    const bool store = owner && row < nq;
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = p.o + ((size_t)b * p.Lq + (store ? row : 0)) * p.ldo + (size_t)h * HD;
    auto put16 = [&](const uint32_t (&v)[16], int c0) {
      uint4 a, c;
      a.x = pack_bf16(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
      a.y = pack_bf16(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
      a.z = pack_bf16(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
      a.w = pack_bf16(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
      c.x = pack_bf16(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
      c.y = pack_bf16(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
      c.z = pack_bf16(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
      c.w = pack_bf16(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
      *reinterpret_cast<uint4*>(dst + c0) = a;
      *reinterpret_cast<uint4*>(dst + c0 + 8) = c;
    };
    constexpr int kGroup = (HD % 48 == 0) ? 48 : 32;   // columns fetched per tcgen05.wait::ld
#pragma unroll 1
    for (int c0 = 0; c0 < HD; c0 += kGroup) {
      uint32_t v0[16], v1[16], v2[16];
      tmem_ld_32x32_x16(tmem_o + lane_off + (uint32_t)c0, v0);
      tmem_ld_32x32_x16(tmem_o + lane_off + (uint32_t)(c0 + 16), v1);
      if constexpr (kGroup == 48) tmem_ld_32x32_x16(tmem_o + lane_off + (uint32_t)(c0 + 32), v2);
      tmem_ld_wait();
      if (store) {
        put16(v0, c0);
        put16(v1, c0 + 16);
        if constexpr (kGroup == 48) put16(v2, c0 + 32);
      }
    }
    if (store && p.lse2 != nullptr) p.lse2[((size_t)b * p.H + h) * p.Lq + row] = m_run + log2f(l_run);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// kv [B*Lk, nb*2*H*HD] (block i: K columns then V columns) -> tc packing, see the header comment
__global__ void kv_cache_pack_tc_kernel(const __nv_bfloat16* __restrict__ kv, long long ldkv, uint8_t* __restrict__ packed,
                                        int B, int Lk, int H, int HD, int nb) {
  pdl_prologue();
  const int nt = (Lk + kTcTile - 1) / kTcTile;
  const int rows_last = ((Lk - (nt - 1) * kTcTile) + 15) & ~15;
  const int lk_pad = (nt - 1) * kTcTile + rows_last;
  const long long per_head = (long long)lk_pad * HD * 4;              // bytes
  const long long units = (long long)B * nb * H * (per_head / 16);    // 16-byte units of output
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / (per_head / 16);
    const long long off = (i % (per_head / 16)) * 16;                 // byte offset inside the head's block
    const int h = (int)(r % H); r /= H;
    const int blk = (int)(r % nb);
    const int b = (int)(r / nb);
    // which tile, which image
    const long long full_tile_bytes = (long long)kTcTile * HD * 4;
    int t = (int)(off / full_tile_bytes);
    if (t > nt - 1) t = nt - 1;
    const long long toff = off - (long long)t * full_tile_bytes;
    const int rows = (t == nt - 1) ? rows_last : kTcTile;
    const long long k_bytes = (long long)rows * HD * 2;
    uint4 out = make_uint4(0, 0, 0, 0);
    const __nv_bfloat16* srcb = kv + (long long)(2 * blk) * H * HD + (long long)h * HD;
    if (toff < k_bytes) {
      // K image [HD/16 chunks][rows][32 B]: this unit = 8 consecutive head-dim elements of one key
      const int chunk = (int)(toff / (rows * 32));
      const int rem = (int)(toff % (rows * 32));
      const int key = rem / 32, usw = (rem % 32) / 16;
      const int u = usw ^ ((key >> 2) & 1);
      const int gk = t * kTcTile + key;
      if (gk < Lk) out = __ldg(reinterpret_cast<const uint4*>(srcb + ((long long)b * Lk + gk) * ldkv + chunk * 16 + u * 8));
    } else {
      // V^T image [rows/16 chunks][HD][32 B]: this unit = 8 consecutive keys of one head-dim element
      const long long voff = toff - k_bytes;
      const int chunk = (int)(voff / (HD * 32));
      const int rem = (int)(voff % (HD * 32));
      const int d = rem / 32, usw = (rem % 32) / 16;
      const int u = usw ^ ((d >> 2) & 1);
      const int key0 = t * kTcTile + chunk * 16 + u * 8;
      uint16_t e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gk = key0 + j;
        e[j] = gk < Lk ? __ldg(reinterpret_cast<const uint16_t*>(srcb + (long long)H * HD + ((long long)b * Lk + gk) * ldkv + d)) : (uint16_t)0;
      }
      out.x = e[0] | ((uint32_t)e[1] << 16);
      out.y = e[2] | ((uint32_t)e[3] << 16);
      out.z = e[4] | ((uint32_t)e[5] << 16);
      out.w = e[6] | ((uint32_t)e[7] << 16);
    }
    *reinterpret_cast<uint4*>(packed + i * 16) = out;
  }
}

static bool al16q(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int tc_lk_pad(int len_k) {
  const int nt = (len_k + kTcTile - 1) / kTcTile;
  return (nt - 1) * kTcTile + (((len_k - (nt - 1) * kTcTile) + 15) & ~15);
}

template <int HD>
static int launch_tc(const TcParams& p, cudaStream_t stream) {
  using Cfg = TcCfg<HD>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_last_error("attention_decode_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done = true;
  }
  launch_pdl(kPdlAttn, attn_decode_tc_kernel<HD>, dim3(p.H, p.B), dim3(kTcThreads), Cfg::kSmemBytes, stream, p);
  return check_launch("attn_decode_tc", stream);
}

}  // namespace b200b

using namespace b200b;

extern "C" size_t b200b_kv_cache_tc_bytes(int batch, int len_k, int heads, int head_dim, int num_blocks) {
  return (size_t)batch * num_blocks * heads * tc_lk_pad(len_k) * head_dim * 4;
}

extern "C" int b200b_kv_cache_pack_tc(const void* kv, int64_t ldkv, void* packed, int batch, int len_k, int heads,
                                      int head_dim, int num_blocks, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!kv || !packed || batch <= 0 || len_k <= 0 || heads <= 0 || num_blocks <= 0 || head_dim <= 0 || (head_dim % 16) ||
      (ldkv % 8) || !al16q(kv) || !al16q(packed)) {
    set_last_error("kv_cache_pack_tc: bad argument (16-byte aligned pointers, head_dim multiple of 16, ldkv of 8)");
    return B200B_ERR_ARG;
  }
  const size_t units = b200b_kv_cache_tc_bytes(batch, len_k, heads, head_dim, num_blocks) / 16;
  const int threads = 256;
  const size_t want = (units + threads - 1) / threads;
  const int blocks = (int)(want < (size_t)148 * 16 ? want : (size_t)148 * 16);
  launch_pdl(kPdlAttn, kv_cache_pack_tc_kernel, dim3(blocks), dim3(threads), 0, stream,
             reinterpret_cast<const __nv_bfloat16*>(kv), (long long)ldkv, reinterpret_cast<uint8_t*>(packed), batch, len_k,
             heads, head_dim, num_blocks);
  return check_launch("kv_cache_pack_tc", stream);
}

extern "C" int b200b_attention_decode_tc(const void* q, int64_t ldq, const void* kv_tc, int block_index, int num_blocks,
                                         void* o, int64_t ldo, float* lse, int batch, int heads, int len_q, int len_k,
                                         int head_dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !kv_tc || !o || batch <= 0 || heads <= 0 || len_q <= 0 || len_q > 64 || len_k <= 0 || block_index < 0 ||
      block_index >= num_blocks) {
    set_last_error("attention_decode_tc: bad argument (need 1 <= len_q <= 64, 0 <= block_index < num_blocks)");
    return B200B_ERR_ARG;
  }
  if (head_dim != 64 && head_dim != 128 && head_dim != 288) {
    set_last_error("attention_decode_tc: head_dim %d not built (64, 128, 288)", head_dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16q(q) || !al16q(kv_tc) || !al16q(o) || (ldq % 8) || (ldo % 8)) {
    set_last_error("attention_decode_tc: q/kv/o must be 16-byte aligned with row pitch multiple of 8 elements");
    return B200B_ERR_ALIGN;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  const long long per_head = (long long)tc_lk_pad(len_k) * head_dim * 4;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.ldq = ldq;
  p.kv = reinterpret_cast<const uint8_t*>(kv_tc) + (long long)block_index * heads * per_head;
  p.hstride = per_head;
  p.bstride = (long long)num_blocks * heads * per_head;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = ldo;
  p.lse2 = lse;
  p.B = batch; p.H = heads; p.Lq = len_q; p.Lk = len_k;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
#ifdef B200B_DIAG
  static const int dbg = [] { const char* e = getenv("B200B_DECODE_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug_copy_only = dbg & 1;
#endif
  switch (head_dim) {
    case 64: return launch_tc<64>(p, stream);
    case 128: return launch_tc<128>(p, stream);
    default: return launch_tc<288>(p, stream);
  }
}

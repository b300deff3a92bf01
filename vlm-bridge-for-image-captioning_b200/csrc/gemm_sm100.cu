// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled shared memory ->
// tcgen05.mma (cta_group::1, 128 x BLOCK_N x 16) -> double-buffered TMEM accumulators ->
// tcgen05.ld epilogue with fused bias / GELU / dropout / residual.
//
// Replaces every nn.Linear of the reference bridge (bridge_module.py:98-100,118,196-198,216,
// 292,295) and the dgrad / wgrad matmuls autograd derives from them.
//
// Roles (192 threads, one CTA per SM):
//   warp 0      TMA producer   : waits empty[s], arms full[s] with the stage byte count, issues
//                                the A and B tile loads
//   warp 1      MMA issuer     : owns TMEM (alloc/dealloc); waits full[s]; lane 0 issues 4
//                                tcgen05.mma per 64-wide k block and commits to empty[s]; after the
//                                last k block commits to tmem_full[acc]
//   warps 2..5  epilogue       : wait tmem_full[acc]; each warp drains its 32-lane TMEM quadrant
//                                with tcgen05.ld (thread = one output row, 32 columns at a time),
//                                applies the epilogue and stores; then arrives on tmem_empty[acc]
// The accumulator is double buffered (2 x BLOCK_N TMEM columns) so the epilogue of tile i
// overlaps the main loop of tile i+1.
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "../../include/b200_bridge.h"
#include "common.cuh"
#include "launch.h"

namespace b200b {

struct GemmKernelParams {
  int m, n, k;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  void* out;
  long long ldo;
  void* aux;
  long long ldaux;
  const float* bias;
  const float* resid;
  long long ldr;
  float beta;
  DropoutCfg drop;
  uint32_t drop_stream;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024;
};

// ---- epilogue for 8 consecutive columns of one row -------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue8(const GemmKernelParams& p, int row, int col,
                                          const uint32_t* v /*8 fp32 bit patterns*/) {
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[i]);

  if constexpr (EPI == B200B_EPI_BF16_BIAS || EPI == B200B_EPI_BF16_BIAS_GELU ||
                EPI == B200B_EPI_F32_BIAS_RESID) {
    if (p.bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
      f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
      f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
    }
  }

  uint4 bits = make_uint4(0, 0, 0, 0);
  const bool use_drop = (EPI == B200B_EPI_BF16_BIAS_GELU || EPI == B200B_EPI_F32_BIAS_RESID ||
                         EPI == B200B_EPI_BF16_DGELU) &&
                        p.drop.thr != 0;
  if (use_drop) {
    const uint64_t group = ((uint64_t)row * (uint64_t)p.n + (uint64_t)col) >> 3;
    bits = dropout_bits8(p.drop, p.drop_stream, group);
  }

  if constexpr (EPI == B200B_EPI_BF16_BIAS) {
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col) = o;
  } else if constexpr (EPI == B200B_EPI_BF16_BIAS_GELU) {
    float h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      f[i] = bf16_round(f[i]);          // the reference Linear output is bf16 under autocast
      h[i] = bf16_round(gelu_erf(f[i]));
      if (use_drop) h[i] = dropout_keep(bits, i, p.drop.thr) ? h[i] * p.drop.scale : 0.0f;
    }
    uint4 u, o;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
    u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    o.x = pack_bf16(h[0], h[1]); o.y = pack_bf16(h[2], h[3]);
    o.z = pack_bf16(h[4], h[5]); o.w = pack_bf16(h[6], h[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.aux) + (long long)row * p.ldaux + col) = u;
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col) = o;
  } else if constexpr (EPI == B200B_EPI_F32_BIAS_RESID) {
    const float* rp = p.resid + (long long)row * p.ldr + col;
    const float4 r0 = *reinterpret_cast<const float4*>(rp);
    const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
    const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = bf16_round(f[i]);
      if (use_drop) y = dropout_keep(bits, i, p.drop.thr) ? bf16_round(y * p.drop.scale) : 0.0f;
      o[i] = r[i] + y;
    }
    float* op = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
    *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
  } else if constexpr (EPI == B200B_EPI_BF16_DGELU) {
    const uint4 uu = *reinterpret_cast<const uint4*>(
        reinterpret_cast<const __nv_bfloat16*>(p.aux) + (long long)row * p.ldaux + col);
    const float u[8] = {bf16_lo(uu.x), bf16_hi(uu.x), bf16_lo(uu.y), bf16_hi(uu.y),
                        bf16_lo(uu.z), bf16_hi(uu.z), bf16_lo(uu.w), bf16_hi(uu.w)};
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float d = bf16_round(f[i]);
      if (use_drop) d = dropout_keep(bits, i, p.drop.thr) ? bf16_round(d * p.drop.scale) : 0.0f;
      g[i] = d * gelu_erf_grad(u[i]);
    }
    uint4 o;
    o.x = pack_bf16(g[0], g[1]); o.y = pack_bf16(g[2], g[3]);
    o.z = pack_bf16(g[4], g[5]); o.w = pack_bf16(g[6], g[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col) = o;
  } else {  // B200B_EPI_F32
    float* op = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
    if (p.beta != 0.0f) {
      const float4 o0 = *reinterpret_cast<const float4*>(op);
      const float4 o1 = *reinterpret_cast<const float4*>(op + 4);
      f[0] += p.beta * o0.x; f[1] += p.beta * o0.y; f[2] += p.beta * o0.z; f[3] += p.beta * o0.w;
      f[4] += p.beta * o1.x; f[5] += p.beta * o1.y; f[6] += p.beta * o1.z; f[7] += p.beta * o1.w;
    }
    *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const GemmKernelParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 128);
    mbar_init(&tmem_empty_bar[1], 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_idx = (tile % p.num_m_blocks) * kBlockM;
      const int n_idx = (tile / p.num_m_blocks) * BLOCK_N;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + (size_t)stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int k_idx = kb * kBlockK;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tm_a, &full_bar[stage], k_idx, m_idx);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j)                   // box {64 m, 64 k} per atom
              tma_load_2d(sa + j * 8192, &tm_a, &full_bar[stage], m_idx + 64 * j, k_idx);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tm_b, &full_bar[stage], k_idx, n_idx);  // box {64 k, BLOCK_N rows}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * 8192, &tm_b, &full_bar[stage], n_idx + 64 * j, k_idx);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N, A_MN, B_MN);
    // K-major: 8-row groups 1024 B apart; advancing 16 k elements = +32 B inside the swizzle row.
    // MN-major: 64-wide MN atoms 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO);
    //           advancing 16 k rows = +2048 B.
    constexpr uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
    constexpr uint32_t a_kstep = A_MN ? 2048u : 32u, b_kstep = B_MN ? 2048u : 32u;
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = umma_smem_desc(sa + k * a_kstep, a_lbo, 1024u);
            const uint64_t bdesc = umma_smem_desc(sb + k * b_kstep, b_lbo, 1024u);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // frees the smem stage
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator ready
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------------------
    const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are visible to this warp
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int m_idx = (tile % p.num_m_blocks) * kBlockM;
      const int n_idx = (tile / p.num_m_blocks) * BLOCK_N;
      const int row = m_idx + quad * 32 + lane;
      const bool row_ok = row < p.m;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        if (n_idx + c0 >= p.n) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = n_idx + c0 + 8 * j;
            if (col < p.n) epilogue8<EPI>(p, row, col, &v[8 * j]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// bf16 row-major 2-D tensor [outer, inner] (inner contiguous), 128B swizzle, box {64, box_outer}
static int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                          uint64_t ld_elems, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return B200B_ERR_DRIVER;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu box_outer=%u base=%p", (int)r,
                   (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_outer, base);
    return B200B_ERR_TENSORMAP;
  }
  return B200B_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p, int grid,
                       cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BLOCK_N, A_MN, B_MN, EPI>;
  static bool configured = false;  // benign race: attribute set is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)GemmCfg<BLOCK_N>::kSmemBytes);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm) failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  kern<<<grid, kGemmThreads, GemmCfg<BLOCK_N>::kSmemBytes, stream>>>(ta, tb, p);
  return check_launch("gemm_tcgen05_kernel", stream);
}

template <int BLOCK_N, bool A_MN, bool B_MN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmKernelParams& p, int grid,
                        cudaStream_t stream) {
  switch (epi) {
    case B200B_EPI_BF16_BIAS:
      return launch_gemm<BLOCK_N, A_MN, B_MN, B200B_EPI_BF16_BIAS>(ta, tb, p, grid, stream);
    case B200B_EPI_BF16_DGELU:
      return launch_gemm<BLOCK_N, A_MN, B_MN, B200B_EPI_BF16_DGELU>(ta, tb, p, grid, stream);
    case B200B_EPI_F32:
      return launch_gemm<BLOCK_N, A_MN, B_MN, B200B_EPI_F32>(ta, tb, p, grid, stream);
    default:
      break;
  }
  if constexpr (!A_MN && !B_MN) {
    switch (epi) {
      case B200B_EPI_BF16_BIAS_GELU:
        return launch_gemm<BLOCK_N, false, false, B200B_EPI_BF16_BIAS_GELU>(ta, tb, p, grid, stream);
      case B200B_EPI_F32_BIAS_RESID:
        return launch_gemm<BLOCK_N, false, false, B200B_EPI_F32_BIAS_RESID>(ta, tb, p, grid, stream);
      default:
        break;
    }
  }
  set_last_error("gemm: epilogue %d not built for a_major=%d b_major=%d", epi, (int)A_MN, (int)B_MN);
  return B200B_ERR_ARG;
}

template <int BLOCK_N>
static int dispatch_major(int a_mn, int b_mn, int epi, const CUtensorMap& ta, const CUtensorMap& tb,
                          const GemmKernelParams& p, int grid, cudaStream_t stream) {
  if (!a_mn && !b_mn) return dispatch_epi<BLOCK_N, false, false>(epi, ta, tb, p, grid, stream);
  if (!a_mn && b_mn) return dispatch_epi<BLOCK_N, false, true>(epi, ta, tb, p, grid, stream);
  if (a_mn && b_mn) return dispatch_epi<BLOCK_N, true, true>(epi, ta, tb, p, grid, stream);
  set_last_error("gemm: a_major=1,b_major=0 is not built");
  return B200B_ERR_ARG;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int choose_block_n(int m, int n, int num_sms) {
  const long long mb = (m + kBlockM - 1) / kBlockM;
  const long long t256 = mb * ((n + 255) / 256), t128 = mb * ((n + 127) / 128);
  const double c256 = (double)((t256 + num_sms - 1) / num_sms) * 256.0;
  const double c128 = (double)((t128 + num_sms - 1) / num_sms) * 128.0 * 1.08;  // 128-wide tiles read more smem per flop
  return c128 < c256 ? 128 : 256;
}

}  // namespace b200b

using namespace b200b;

extern "C" int b200b_gemm(const b200b_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (a == nullptr || a->a == nullptr || a->b == nullptr || a->out == nullptr) {
    set_last_error("gemm: null argument");
    return B200B_ERR_ARG;
  }
  // m and k may be ragged (TMA zero-fills out-of-bounds box elements; the epilogue masks rows);
  // n must be a multiple of 8 because the epilogue stores 8 columns at a time.
  if (a->m <= 0 || a->n <= 0 || a->k <= 0 || (a->n % 8) != 0) {
    set_last_error("gemm: need m,n,k > 0 and n a multiple of 8 (got m=%d n=%d k=%d)", a->m, a->n, a->k);
    return B200B_ERR_SHAPE;
  }
  const int epi = a->epilogue;
  if (epi < 0 || epi > B200B_EPI_F32) {
    set_last_error("gemm: bad epilogue %d", epi);
    return B200B_ERR_ARG;
  }
  const bool out_f32 = (epi == B200B_EPI_F32_BIAS_RESID || epi == B200B_EPI_F32);
  if (!aligned16(a->a) || !aligned16(a->b) || !aligned16(a->out) || (a->lda % 8) || (a->ldb % 8) ||
      (a->ldo % (out_f32 ? 4 : 8))) {
    set_last_error("gemm: operands must be 16-byte aligned with 16-byte aligned row pitch");
    return B200B_ERR_ALIGN;
  }
  if (a->bias != nullptr && !aligned16(a->bias)) {
    set_last_error("gemm: bias must be 16-byte aligned");
    return B200B_ERR_ALIGN;
  }
  if (epi == B200B_EPI_BF16_BIAS_GELU || epi == B200B_EPI_BF16_DGELU) {
    if (a->aux == nullptr || !aligned16(a->aux) || (a->ldaux % 8)) {
      set_last_error("gemm: epilogue %d needs a 16-byte aligned aux tensor", epi);
      return B200B_ERR_ARG;
    }
  }
  if (epi == B200B_EPI_F32_BIAS_RESID) {
    if (a->resid == nullptr || !aligned16(a->resid) || (a->ldr % 4)) {
      set_last_error("gemm: residual epilogue needs a 16-byte aligned resid tensor");
      return B200B_ERR_ARG;
    }
  }
  if (!(a->dropout_p >= 0.0f && a->dropout_p < 1.0f)) {
    set_last_error("gemm: dropout_p must be in [0,1)");
    return B200B_ERR_ARG;
  }
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc != B200B_OK) return rc;

  int block_n = a->block_n;
  if (block_n == 0) block_n = choose_block_n(a->m, a->n, num_sms);
  if (block_n != 128 && block_n != 256) {
    set_last_error("gemm: block_n must be 0, 128 or 256");
    return B200B_ERR_ARG;
  }

  CUtensorMap ta, tb;
  if (!a->a_major) rc = make_tmap_bf16(&ta, a->a, (uint64_t)a->k, (uint64_t)a->m, (uint64_t)a->lda, kBlockM);
  else             rc = make_tmap_bf16(&ta, a->a, (uint64_t)a->m, (uint64_t)a->k, (uint64_t)a->lda, kBlockK);
  if (rc != B200B_OK) return rc;
  if (!a->b_major) rc = make_tmap_bf16(&tb, a->b, (uint64_t)a->k, (uint64_t)a->n, (uint64_t)a->ldb, (uint32_t)block_n);
  else             rc = make_tmap_bf16(&tb, a->b, (uint64_t)a->n, (uint64_t)a->k, (uint64_t)a->ldb, kBlockK);
  if (rc != B200B_OK) return rc;

  GemmKernelParams p;
  p.m = a->m; p.n = a->n; p.k = a->k;
  p.num_m_blocks = (a->m + kBlockM - 1) / kBlockM;
  p.num_n_blocks = (a->n + block_n - 1) / block_n;
  p.num_k_blocks = (a->k + kBlockK - 1) / kBlockK;
  p.out = a->out; p.ldo = a->ldo;
  p.aux = a->aux; p.ldaux = a->ldaux;
  p.bias = a->bias;
  p.resid = a->resid; p.ldr = a->ldr;
  p.beta = a->beta;
  p.drop = make_dropout_cfg(a->dropout_p, a->seed);
  p.drop_stream = a->dropout_stream;

  const long long tiles = (long long)p.num_m_blocks * p.num_n_blocks;
  const int grid = (int)(tiles < num_sms ? tiles : num_sms);
  if (block_n == 256) return dispatch_major<256>(a->a_major, a->b_major, epi, ta, tb, p, grid, stream);
  return dispatch_major<128>(a->a_major, a->b_major, epi, ta, tb, p, grid, stream);
}

// Diagnostic probe, not part of the bridge path: one CTA issues tcgen05.mma (cta_group::1, kind::f16,
// bf16 x bf16 -> fp32) on operands it lays out itself in shared memory in the canonical K-major
// swizzled layout of a chosen swizzle width, then dumps ALL 128 TMEM lanes of the accumulator.
// It answers the layout questions a tcgen05 attention kernel for this bridge depends on -- head dim
// 288 is 9 x 32 = 18 x 16 elements but not a multiple of the 64-element rows of SWIZZLE_128B, and a
// decode step has at most 64 query rows (M = 64):
//   * do the 64-byte and 32-byte swizzle descriptors (layout codes 4 and 6) produce A B^T?
//   * where do the 64 rows of an M = 64 accumulator live among the 128 TMEM lanes?
// tests/gpu_checks/probe_tcgen05.py runs it against torch.matmul and prints the lane map.
#include <string.h>

#include "common.cuh"
#include "launch.h"

namespace b200b {

__device__ __forceinline__ uint64_t umma_smem_desc_mode(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                        uint32_t layout_code) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_code << 61;
  return d;
}

// byte offset of element (row, k) of a [rows x K] K-major operand stored as chunks of `w` bytes per
// row ([K*2/w chunks][rows][w bytes]), 16-byte units XOR-swizzled inside each 8-row atom
__device__ __forceinline__ uint32_t swz_offset(int row, int k, int rows, int w) {
  const int per = w / 2;  // elements per chunk row
  const int chunk = k / per, kin = k % per;
  const uint32_t unit = (uint32_t)(kin / 8);
  const uint32_t shift = (w == 128) ? 0 : (w == 64 ? 1 : 2);
  const uint32_t mask = (uint32_t)(w / 16 - 1);
  const uint32_t sw = (unit ^ (((uint32_t)row & 7u) >> shift)) & mask;
  return (uint32_t)chunk * rows * w + (uint32_t)row * w + sw * 16 + (uint32_t)(kin % 8) * 2;
}

__global__ void __launch_bounds__(128) probe_umma_kernel(const __nv_bfloat16* __restrict__ a,
                                                         const __nv_bfloat16* __restrict__ b, float* __restrict__ dump,
                                                         int m, int n, int k, int w, int reps,
                                                         long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_probe[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t s0 = (smem_u32(smem_probe) + 1023u) & ~1023u;
  uint8_t* base = smem_probe + (s0 - smem_u32(smem_probe));
  uint8_t* sa = base;
  uint8_t* sb = base + (((size_t)m * k * 2 + 1023) & ~(size_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < m * k; i += blockDim.x)
    *reinterpret_cast<__nv_bfloat16*>(sa + swz_offset(i / k, i % k, m, w)) = a[i];
  for (int i = threadIdx.x; i < n * k; i += blockDim.x)
    *reinterpret_cast<__nv_bfloat16*>(sb + swz_offset(i / k, i % k, n, w)) = b[i];
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  const uint32_t ncols = 2 * n <= 32 ? 32 : (2 * n <= 64 ? 64 : (2 * n <= 128 ? 128 : (2 * n <= 256 ? 256 : 512)));
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, ncols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (threadIdx.x == 0) {
    const uint32_t code = (w == 128) ? 2u : (w == 64 ? 4u : 6u);
    const uint32_t idesc = umma_idesc_bf16(m, n, false, false);
    const int per = w / 2;
    // timing: the same MMA sequence `reps` more times into a scratch accumulator (columns n..2n) first
    if (reps > 0 && cycles != nullptr) {
      __shared__ __align__(8) uint64_t tbar;
      mbar_init(&tbar, 1);
      fence_barrier_init();
      const long long t0 = clock64();
      // the same (first) k-step every time, descriptors hoisted: the loop body is one tcgen05.mma, so the
      // completion time measures the tensor pipe, not the issuing thread
      const uint64_t da0 = umma_smem_desc_mode(smem_u32(sa), 16, 8 * w, code);
      const uint64_t db0 = umma_smem_desc_mode(smem_u32(sb), 16, 8 * w, code);
      const uint32_t dscr = tmem_base + (uint32_t)n;
      const int total = reps * (k / 16);
#pragma unroll 8
      for (int r = 0; r < total; ++r) umma_bf16(dscr, da0, db0, idesc, 1u);
      umma_commit(&tbar);
      const long long t1 = clock64();
      mbar_wait(&tbar, 0);
      const long long t2 = clock64();
      cycles[0] = t1 - t0;   // issue time
      cycles[1] = t2 - t0;   // until the last MMA has completed
    }
    uint32_t acc = 0;
    for (int c = 0; c < k / per; ++c)
      for (int ks = 0; ks < per / 16; ++ks) {
        const uint64_t da = umma_smem_desc_mode(smem_u32(sa) + (uint32_t)c * m * w + ks * 32, 16, 8 * w, code);
        const uint64_t db = umma_smem_desc_mode(smem_u32(sb) + (uint32_t)c * n * w + ks * 32, 16, 8 * w, code);
        umma_bf16(tmem_base, da, db, idesc, acc);
        acc = 1;
      }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  // every warp dumps its 32 TMEM lanes x n columns
  for (int c0 = 0; c0 < n; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < n; ++j) dump[(size_t)(warp * 32 + lane) * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace b200b

using namespace b200b;

// a bf16 [m, k], b bf16 [n, k] row-major; dump fp32 [128, n] = the accumulator's 128 TMEM lanes.
// m in {64, 128}; n % 16 == 0, 16 <= n <= 256; k % (swizzle_bytes / 2) == 0; swizzle_bytes in {32, 64, 128}.
extern "C" int b200b_probe_umma(const void* a, const void* b, float* dump, int m, int n, int k, int swizzle_bytes,
                                int reps, long long* cycles2, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !b || !dump || (m != 64 && m != 128) || n < 16 || n > 256 || (n % 16) ||
      (swizzle_bytes != 32 && swizzle_bytes != 64 && swizzle_bytes != 128) || k <= 0 || (k % (swizzle_bytes / 2))) {
    set_last_error("probe_umma: bad argument");
    return B200B_ERR_ARG;
  }
  const size_t smem = (((size_t)m * k * 2 + 1023) & ~(size_t)1023) + (size_t)n * k * 2 + 2048;
  if (smem > 220 * 1024) {
    set_last_error("probe_umma: operands do not fit in shared memory");
    return B200B_ERR_SHAPE;
  }
  cudaError_t e = cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_error("probe_umma: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  probe_umma_kernel<<<1, 128, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                              reinterpret_cast<const __nv_bfloat16*>(b), dump, m, n, k, swizzle_bytes, reps, cycles2);
  return check_launch("probe_umma", stream);
}

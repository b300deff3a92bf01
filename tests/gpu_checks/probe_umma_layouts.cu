// Bring-up probe (test infrastructure, NOT part of libb200_bridge.so): one CTA issues tcgen05.mma
// (cta_group::1, kind::f16, bf16 x bf16 -> fp32) on operands it lays out itself, in every operand
// form the tcgen05 attention kernels (csrc/attention_tc_train.cu) rely on, and dumps all 128 TMEM
// lanes of the accumulator. tests/gpu_checks/probe_umma_layouts.py compares with torch.matmul.
//
//   D[M x N] = A[M x K] * B[N x K]^T
//   a_mode: 0 = shared memory, K-major   ([K/per chunks][M rows][w bytes], 16-byte units XOR-swizzled)
//           1 = shared memory, MN-major  ([M/per chunks][K rows][w bytes]; the M index is contiguous)
//           2 = tensor memory            (lane = row, column c holds elements 2c, 2c+1 of the row)
//   b_mode: 0 / 1 as above with N in place of M
//   w = swizzle width in bytes (32 / 64 / 128), per = w / 2 elements
//   mn_variant: 0 = descriptor LBO = chunk stride, SBO = 8 rows * w;  1 = the two swapped
//
// Build: tests/gpu_checks/build_probes.py -> tests/gpu_checks/libprobe_umma.so
#include <string.h>

#include "../../vlm-bridge-for-image-captioning_b200/csrc/common.cuh"

namespace b200b {

__device__ __forceinline__ uint64_t desc_mode(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_code) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_code << 61;
  return d;
}

// element (major index i, minor/contiguous index j) of an operand stored as [j / per chunks][rows = extent of i][w bytes]
__device__ __forceinline__ uint32_t swz_off(int i, int j, int rows, int w) {
  const int per = w / 2;
  const int chunk = j / per, jin = j % per;
  const uint32_t unit = (uint32_t)(jin / 8);
  const uint32_t shift = (w == 128) ? 0 : (w == 64 ? 1 : 2);
  const uint32_t mask = (uint32_t)(w / 16 - 1);
  const uint32_t sw = (unit ^ (((uint32_t)i & 7u) >> shift)) & mask;
  return (uint32_t)chunk * rows * w + (uint32_t)i * w + sw * 16 + (uint32_t)(jin % 8) * 2;
}

__device__ __forceinline__ void umma_bf16_tmem_a(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

struct ProbeArgs {
  const __nv_bfloat16* a;
  const __nv_bfloat16* b;
  float* dump;
  int m, n, k;
  int a_mode, b_mode, wa, wb, mn_variant;
};

__global__ void __launch_bounds__(128) probe_layouts_kernel(const ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_probe[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t s0 = (smem_u32(smem_probe) + 1023u) & ~1023u;
  uint8_t* base = smem_probe + (s0 - smem_u32(smem_probe));
  uint8_t* sa = base;
  uint8_t* sb = base + (((size_t)p.m * p.k * 2 + 1023) & ~(size_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = p.m, n = p.n, k = p.k;

  if (p.a_mode == 0)
    for (int i = threadIdx.x; i < m * k; i += blockDim.x)
      *reinterpret_cast<__nv_bfloat16*>(sa + swz_off(i / k, i % k, m, p.wa)) = p.a[i];
  else if (p.a_mode == 1)
    for (int i = threadIdx.x; i < m * k; i += blockDim.x)
      *reinterpret_cast<__nv_bfloat16*>(sa + swz_off(i % k, i / k, k, p.wa)) = p.a[i];
  if (p.b_mode == 0)
    for (int i = threadIdx.x; i < n * k; i += blockDim.x)
      *reinterpret_cast<__nv_bfloat16*>(sb + swz_off(i / k, i % k, n, p.wb)) = p.b[i];
  else
    for (int i = threadIdx.x; i < n * k; i += blockDim.x)
      *reinterpret_cast<__nv_bfloat16*>(sb + swz_off(i % k, i / k, k, p.wb)) = p.b[i];
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t a_col = 256;   // A operand in tensor memory: columns [256, 256 + k/2)

  if (p.a_mode == 2) {
    // thread = row (M = 128: lane = row); 8 packed words (16 elements) per store
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < k / 2; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat16 lo = p.a[(size_t)row * k + 2 * (c0 + j)], hi = p.a[(size_t)row * k + 2 * (c0 + j) + 1];
        v[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      tmem_st_32x32_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + a_col + (uint32_t)c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (threadIdx.x == 0) {
    auto code = [](int w) { return (w == 128) ? 2u : (w == 64 ? 4u : 6u); };
    const uint32_t idesc = umma_idesc_bf16(m, n, p.a_mode == 1, p.b_mode == 1);
    for (int s = 0; s < k / 16; ++s) {
      uint64_t da = 0, db = 0;
      if (p.a_mode == 0) {
        const int per = p.wa / 2;
        da = desc_mode(smem_u32(sa) + (uint32_t)((s * 16) / per) * m * p.wa + (uint32_t)(((s * 16) % per) * 2), 16,
                       8 * p.wa, code(p.wa));
      } else if (p.a_mode == 1) {
        const uint32_t lbo = (uint32_t)k * p.wa, sbo = 8u * p.wa;
        da = desc_mode(smem_u32(sa) + (uint32_t)s * 16 * p.wa, p.mn_variant ? sbo : lbo, p.mn_variant ? lbo : sbo,
                       code(p.wa));
      }
      if (p.b_mode == 0) {
        const int per = p.wb / 2;
        db = desc_mode(smem_u32(sb) + (uint32_t)((s * 16) / per) * n * p.wb + (uint32_t)(((s * 16) % per) * 2), 16,
                       8 * p.wb, code(p.wb));
      } else {
        const uint32_t lbo = (uint32_t)k * p.wb, sbo = 8u * p.wb;
        db = desc_mode(smem_u32(sb) + (uint32_t)s * 16 * p.wb, p.mn_variant ? sbo : lbo, p.mn_variant ? lbo : sbo,
                       code(p.wb));
      }
      if (p.a_mode == 2)
        umma_bf16_tmem_a(tmem_base, tmem_base + a_col + (uint32_t)(s * 8), db, idesc, s > 0);
      else
        umma_bf16(tmem_base, da, db, idesc, s > 0);
    }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < n; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < n; ++j) p.dump[(size_t)(warp * 32 + lane) * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace b200b

using namespace b200b;

// a bf16 [m, k], b bf16 [n, k] row-major; dump fp32 [128, n]. Returns 0, -1 (bad argument) or a cudaError_t.
extern "C" int probe_umma_layouts(const void* a, const void* b, float* dump, int m, int n, int k, int a_mode, int b_mode,
                                  int wa, int wb, int mn_variant, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !b || !dump || (m != 64 && m != 128) || n < 16 || n > 256 || (n % 16) || k <= 0 || (k % 16)) return -1;
  if (a_mode == 2 && (m != 128 || k > 512)) return -1;
  const size_t smem = (((size_t)m * k * 2 + 1023) & ~(size_t)1023) + (size_t)n * k * 2 + 2048;
  if (smem > 220 * 1024) return -1;
  cudaError_t e = cudaFuncSetAttribute(probe_layouts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProbeArgs p;
  p.a = reinterpret_cast<const __nv_bfloat16*>(a);
  p.b = reinterpret_cast<const __nv_bfloat16*>(b);
  p.dump = dump;
  p.m = m; p.n = n; p.k = k;
  p.a_mode = a_mode; p.b_mode = b_mode; p.wa = wa; p.wb = wb; p.mn_variant = mn_variant;
  probe_layouts_kernel<<<1, 128, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

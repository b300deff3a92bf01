// Gradient all-reduce over NVLink/NVSwitch with in-switch reduction (NVLS multimem), written for the
// data-parallel exchange of the bridge's gradient buckets (the reference has no
// distributed code, so there is no reference file to cite -- the operation is "average this slice
// of the gradient arena over all ranks").
//
// Every rank owns 1/world of each bucket ("shard"). Two-shot algorithm, in place on a symmetric
// buffer that all ranks have mapped both directly and through one multicast address:
//   barrier A   all ranks' producers of this bucket have finished (flags in peer memory)
//   reduce      multimem.ld_reduce over my shard: the switch fetches the 16-byte unit from every
//               rank, adds the values in fp32 and returns ONE result
//   broadcast   multimem.st of (scale * sum) to the multicast address: the switch writes it into
//               every rank's copy
//   barrier B   all ranks' broadcasts have landed
//   convert     (bf16 buckets, optional) this rank turns the whole averaged bucket into the fp32
//               .grad arena, block b converting exactly the units whose producers it has just
//               synchronised with in barrier B
// Per rank and direction the links carry ~(1 + 1/world)x the bucket (a ring carries 2*(world-1)/world).
//
// B200B_NVLS_OUT_MULTICAST: the broadcast goes out as fp32 to a second symmetric buffer (the .grad
// arena) instead of in place as bf16: 8-byte multimem.ld_reduce of 4 bf16 -> 16-byte multimem.st of
// 4 fp32, both fully coalesced. The incoming link carries 2x the broadcast bytes, but the separate
// bf16 -> fp32 pass over the whole arena (0.95 GB of HBM traffic per step, the single worst neighbour
// of the backward GEMMs: profiles/r01_exp_overlap_2gpu_v1.jsonl) disappears.
// B200B_NVLS_EXCLUSIVE_SMS: the grid is launched as CTA pairs (clusters of 2 = one TPC) that claim
// their SMs' whole shared memory, so no compute CTA is ever co-resident; together with
// b200b_set_sm_limit() on the compute side (the persistent GEMMs then leave those TPCs alone) the
// exchange neither slows the statically scheduled GEMM CTAs nor waits for them.
// The kernel uses no shared memory and ~32 registers per thread, so its CTAs share SMs with the
// 1-CTA-per-SM GEMMs of the backward pass instead of evicting them; the caller picks the shape of the
// grid -- many small CTAs spread the outstanding multimem requests thinly over all SMs, which matters
// because the compute kernels are statically scheduled and run at the pace of their slowest SM.
//
// Flags: rank r owns uint32 flags[kMaxBlocks][2][kMaxRanks] in symmetric memory; flags[b][ph][q] is
// written only by block b of rank q. Values are the monotonically increasing collective number
// (same sequence on every rank), so flags are never reset.
#include <stdio.h>
#include <string.h>

#include "../../include/b200_bridge.h"
#include "common.cuh"
#include "launch.h"

namespace b200b {

constexpr int kNvlsMaxThreads = 1024;
// 16-byte units in flight per thread: 4 (default), 8 or 16, chosen by the caller (bits 8..15 of `flags`). A
// multimem.ld_reduce round trip through the switch takes ~5 us, so the exchange is bound by bytes in flight:
// 32 CTAs x 512 threads x 4 units = 1 MB gives ~200 GB/s; fewer, fatter CTAs need the deeper unroll.

struct NvlsParams {
  uint64_t mc;                           // multicast address of the first unit of the bucket
  const uint4* local;                    // this rank's own copy of the bucket (same units)
  float4* out_f32;                       // fp32 destination of the whole bucket, or nullptr
  uint64_t mc_out;                       // B200B_NVLS_OUT_MULTICAST: multicast address of the fp32 destination
  uint32_t* flags[B200B_NVLS_MAX_RANKS]; // every rank's flag array as mapped in this process
  long long units;                       // 16-byte units in the bucket
  long long per;                         // units per shard
  float scale;
  uint32_t epoch;                        // collective number (added to *epoch_base when that is given)
  const uint32_t* epoch_base;            // device counter advanced by the caller once per step, or nullptr
  int rank, world;
  long long timeout_cycles;              // barrier wait limit in SM clock cycles
  uint32_t* error_word;                  // host-visible word that receives a code when the wait expires, or nullptr
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Blocks with the same index on all ranks meet here. Thread q < world signals rank q and waits for
// rank q's signal. Everything the block did before is ordered before the signal (bar.sync +
// fence), everything after the barrier is ordered after the peers' signals.
__device__ __forceinline__ void barrier_blocks(const NvlsParams& p, uint32_t epoch, int phase) {
  __syncthreads();
  if (threadIdx.x < p.world) {
    const int q = threadIdx.x;
    const int slot = (blockIdx.x * 2 + phase) * B200B_NVLS_MAX_RANKS;
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    st_release_sys(p.flags[q] + slot + p.rank, epoch);
    const uint32_t* mine = p.flags[p.rank] + slot + q;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > p.timeout_cycles) {
        // a rank never arrived within the limit (default: minutes, like a process-group timeout). With an error
        // word the failure is reported to the host, which raises at its next look (parallel.py), and this
        // collective is abandoned -- its output is garbage, but the CUDA context survives; without one, trap.
        if (p.error_word != nullptr) {
          // code: 1 + waiting-for rank | phase << 8 | block << 12
          const uint32_t code = (uint32_t)(q + 1) | ((uint32_t)phase << 8) | ((uint32_t)blockIdx.x << 12);
          asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.error_word), "r"(code) : "memory");
          break;
        }
        printf("b200b: nvls all-reduce barrier timed out (rank %d block %d phase %d waiting for rank %d)\n", p.rank,
               (int)blockIdx.x, phase, q);
        __trap();
      }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncthreads();
}

template <bool BF16>
__device__ __forceinline__ uint4 multimem_ld_reduce(uint64_t addr) {
  uint4 v;
  if constexpr (BF16) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(addr)
                 : "memory");
  } else {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(addr)
                 : "memory");
  }
  return v;
}
__device__ __forceinline__ void multimem_st(uint64_t addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

template <bool BF16>
__device__ __forceinline__ uint4 scale_unit(uint4 v, float s) {
  if (s == 1.0f) return v;
  if constexpr (BF16) {
    v.x = pack_bf16(bf16_lo(v.x) * s, bf16_hi(v.x) * s);
    v.y = pack_bf16(bf16_lo(v.y) * s, bf16_hi(v.y) * s);
    v.z = pack_bf16(bf16_lo(v.z) * s, bf16_hi(v.z) * s);
    v.w = pack_bf16(bf16_lo(v.w) * s, bf16_hi(v.w) * s);
  } else {
    v.x = __float_as_uint(__uint_as_float(v.x) * s);
    v.y = __float_as_uint(__uint_as_float(v.y) * s);
    v.z = __float_as_uint(__uint_as_float(v.z) * s);
    v.w = __float_as_uint(__uint_as_float(v.w) * s);
  }
  return v;
}

template <bool BF16, int kNvlsUnroll>
__global__ void __launch_bounds__(kNvlsMaxThreads) allreduce_nvls_kernel(const NvlsParams p) {
  const uint32_t epoch = p.epoch + (p.epoch_base != nullptr ? __ldcg(p.epoch_base) : 0u);
  barrier_blocks(p, epoch, 0);
  const int nthreads = (int)blockDim.x;
  const long long stride = (long long)gridDim.x * nthreads;
  {
    const long long lo = (long long)p.rank * p.per;
    const long long hi = min(lo + p.per, p.units);
    for (long long i = lo + (long long)blockIdx.x * nthreads + threadIdx.x; i < hi; i += kNvlsUnroll * stride) {
      uint4 v[kNvlsUnroll];
#pragma unroll
      for (int u = 0; u < kNvlsUnroll; ++u)
        if (i + u * stride < hi) v[u] = multimem_ld_reduce<BF16>(p.mc + 16ull * (unsigned long long)(i + u * stride));
#pragma unroll
      for (int u = 0; u < kNvlsUnroll; ++u)
        if (i + u * stride < hi)
          multimem_st(p.mc + 16ull * (unsigned long long)(i + u * stride), scale_unit<BF16>(v[u], p.scale));
    }
  }
  barrier_blocks(p, epoch, 1);
  if constexpr (BF16) {
    if (p.out_f32 != nullptr) {
      // block b converts, in every shard, the units block b of that shard's owner has broadcast
      for (int r = 0; r < p.world; ++r) {
        const long long lo = (long long)r * p.per;
        const long long hi = min(lo + p.per, p.units);
        for (long long i = lo + (long long)blockIdx.x * nthreads + threadIdx.x; i < hi; i += kNvlsUnroll * stride) {
          uint4 v[kNvlsUnroll];
#pragma unroll
          for (int u = 0; u < kNvlsUnroll; ++u)
            if (i + u * stride < hi) v[u] = __ldcg(p.local + i + u * stride);  // L2: the data arrived over NVLink
#pragma unroll
          for (int u = 0; u < kNvlsUnroll; ++u) {
            const long long j = i + u * stride;
            if (j < hi) {
              __stcs(p.out_f32 + 2 * j, make_float4(bf16_lo(v[u].x), bf16_hi(v[u].x), bf16_lo(v[u].y), bf16_hi(v[u].y)));
              __stcs(p.out_f32 + 2 * j + 1,
                     make_float4(bf16_lo(v[u].z), bf16_hi(v[u].z), bf16_lo(v[u].w), bf16_hi(v[u].w)));
            }
          }
        }
      }
    }
  }
}


__device__ __forceinline__ uint2 multimem_ld_reduce_bf16x4(uint64_t addr) {
  uint2 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v2.bf16x2 {%0,%1}, [%2];"
               : "=r"(v.x), "=r"(v.y)
               : "l"(addr)
               : "memory");
  return v;
}

// bf16 bucket in, fp32 broadcast out (B200B_NVLS_OUT_MULTICAST). Units are 8 bytes (4 bf16 -> 4 fp32).
constexpr int kBcastUnroll = 16;
__global__ void __launch_bounds__(kNvlsMaxThreads) allreduce_nvls_bcast32_kernel(const NvlsParams p) {
  const uint32_t epoch = p.epoch + (p.epoch_base != nullptr ? __ldcg(p.epoch_base) : 0u);
  barrier_blocks(p, epoch, 0);
  const int nthreads = (int)blockDim.x;
  const long long stride = (long long)gridDim.x * nthreads;
  const long long units = 2 * p.units;  // 8-byte units
  const long long per = (units + p.world - 1) / p.world;
  const long long lo = (long long)p.rank * per;
  const long long hi = min(lo + per, units);
  for (long long i = lo + (long long)blockIdx.x * nthreads + threadIdx.x; i < hi; i += kBcastUnroll * stride) {
    uint2 v[kBcastUnroll];
#pragma unroll
    for (int u = 0; u < kBcastUnroll; ++u)
      if (i + u * stride < hi) v[u] = multimem_ld_reduce_bf16x4(p.mc + 8ull * (unsigned long long)(i + u * stride));
#pragma unroll
    for (int u = 0; u < kBcastUnroll; ++u)
      if (i + u * stride < hi) {
        uint4 f;
        f.x = __float_as_uint(bf16_lo(v[u].x) * p.scale);
        f.y = __float_as_uint(bf16_hi(v[u].x) * p.scale);
        f.z = __float_as_uint(bf16_lo(v[u].y) * p.scale);
        f.w = __float_as_uint(bf16_hi(v[u].y) * p.scale);
        multimem_st(p.mc_out + 16ull * (unsigned long long)(i + u * stride), f);
      }
  }
  barrier_blocks(p, epoch, 1);
}

}  // namespace b200b

using namespace b200b;

extern "C" int b200b_allreduce_nvls(const b200b_nvls_comm* comm, int dtype, int64_t byte_offset, int64_t bytes,
                                    float scale, float* out_f32, uint32_t epoch, const uint32_t* epoch_base,
                                    int blocks, int threads, uint32_t flags, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (comm == nullptr || comm->multicast_base == nullptr || comm->local_base == nullptr) {
    set_last_error("allreduce_nvls: null communicator / buffer");
    return B200B_ERR_ARG;
  }
  if (comm->world < 2 || comm->world > B200B_NVLS_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world) {
    set_last_error("allreduce_nvls: need 2 <= world <= %d and 0 <= rank < world (rank=%d world=%d)",
                   B200B_NVLS_MAX_RANKS, comm->rank, comm->world);
    return B200B_ERR_ARG;
  }
  for (int q = 0; q < comm->world; ++q) {
    if (comm->flags[q] == nullptr) {
      set_last_error("allreduce_nvls: flag array of rank %d is null", q);
      return B200B_ERR_ARG;
    }
  }
  if (dtype != B200B_DTYPE_BF16 && dtype != B200B_DTYPE_F32) {
    set_last_error("allreduce_nvls: dtype must be B200B_DTYPE_BF16 or B200B_DTYPE_F32");
    return B200B_ERR_ARG;
  }
  if (bytes <= 0 || (bytes % 16) != 0 || (byte_offset % 16) != 0 || byte_offset < 0) {
    set_last_error("allreduce_nvls: offset and size must be multiples of 16 bytes (offset=%lld bytes=%lld)",
                   (long long)byte_offset, (long long)bytes);
    return B200B_ERR_ALIGN;
  }
  if (out_f32 != nullptr && (dtype != B200B_DTYPE_BF16 || (reinterpret_cast<uintptr_t>(out_f32) & 15))) {
    set_last_error("allreduce_nvls: out_f32 needs a bf16 bucket and a 16-byte aligned destination");
    return B200B_ERR_ARG;
  }
  const bool out_mc = (flags & B200B_NVLS_OUT_MULTICAST) != 0;
  const bool exclusive = (flags & B200B_NVLS_EXCLUSIVE_SMS) != 0;
  if (out_mc && out_f32 == nullptr) {
    set_last_error("allreduce_nvls: B200B_NVLS_OUT_MULTICAST needs the multicast address of the fp32 destination");
    return B200B_ERR_ARG;
  }
  if (exclusive && (blocks % 2) != 0) {
    set_last_error("allreduce_nvls: B200B_NVLS_EXCLUSIVE_SMS launches CTA pairs: blocks must be even");
    return B200B_ERR_ARG;
  }
  if (blocks <= 0 || blocks > B200B_NVLS_MAX_BLOCKS || threads < 32 || threads > kNvlsMaxThreads || (threads % 32)) {
    set_last_error("allreduce_nvls: blocks must be in 1..%d and threads a multiple of 32 in 32..%d", B200B_NVLS_MAX_BLOCKS,
                   kNvlsMaxThreads);
    return B200B_ERR_ARG;
  }
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc != B200B_OK) return rc;
  NvlsParams p;
  p.mc = reinterpret_cast<uint64_t>(comm->multicast_base) + (uint64_t)byte_offset;
  p.local = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(comm->local_base) + byte_offset);
  p.out_f32 = reinterpret_cast<float4*>(out_f32);
  for (int q = 0; q < B200B_NVLS_MAX_RANKS; ++q)
    p.flags[q] = q < comm->world ? reinterpret_cast<uint32_t*>(comm->flags[q]) : nullptr;
  p.units = bytes / 16;
  p.per = (p.units + comm->world - 1) / comm->world;
  p.scale = scale;
  p.epoch = epoch;
  p.epoch_base = epoch_base;
  p.rank = comm->rank;
  p.world = comm->world;
  p.mc_out = out_mc ? reinterpret_cast<uint64_t>(out_f32) : 0;
  if (out_mc) p.out_f32 = nullptr;
  const int unroll = (int)((flags >> 8) & 0xffu);
  if (unroll != 0 && unroll != 4 && unroll != 8 && unroll != 16) {
    set_last_error("allreduce_nvls: units in flight per thread (flags bits 8..15) must be 0 (= 4), 4, 8 or 16");
    return B200B_ERR_ARG;
  }
  void (*kern)(const NvlsParams);
  if (out_mc) kern = allreduce_nvls_bcast32_kernel;
  else if (dtype == B200B_DTYPE_BF16)
    kern = unroll == 16 ? allreduce_nvls_kernel<true, 16> : (unroll == 8 ? allreduce_nvls_kernel<true, 8> : allreduce_nvls_kernel<true, 4>);
  else
    kern = unroll == 16 ? allreduce_nvls_kernel<false, 16> : (unroll == 8 ? allreduce_nvls_kernel<false, 8> : allreduce_nvls_kernel<false, 4>);
  {
    // wait limit of the cross-rank barriers: comm->timeout_s seconds (<= 0: 600 s), converted with the SM clock
    static int khz = 0;
    if (khz == 0) {
      int dev = 0, v = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev);
      khz = v > 0 ? v : 1965000;
    }
    const long long secs = comm->timeout_s > 0 ? comm->timeout_s : 600;
    p.timeout_cycles = secs * (long long)khz * 1000LL;
    p.error_word = reinterpret_cast<uint32_t*>(comm->error_word);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (exclusive) {
    // claim the SM: no CTA that needs more than the few KB left can become co-resident
    constexpr int kHogBytes = 200 * 1024;
    static bool attr_done[7] = {false, false, false, false, false, false, false};
    const int ui = unroll == 16 ? 2 : (unroll == 8 ? 1 : 0);
    const int ki = out_mc ? 0 : (dtype == B200B_DTYPE_BF16 ? 1 + ui : 4 + ui);
    if (!attr_done[ki]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kHogBytes);
      if (e != cudaSuccess) {
        set_last_error("allreduce_nvls: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return (int)e;
      }
      attr_done[ki] = true;
    }
    cfg.dynamicSmemBytes = kHogBytes;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) {
    set_last_error("allreduce_nvls: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return check_launch(out_mc ? "allreduce_nvls_bcast32" : "allreduce_nvls", stream);
}

extern "C" size_t b200b_allreduce_nvls_flag_bytes(void) {
  return (size_t)B200B_NVLS_MAX_BLOCKS * 2 * B200B_NVLS_MAX_RANKS * sizeof(uint32_t);
}

// HBM-bound row kernels of the bridge: LayerNorm forward / backward, column reductions for the
// bias and LayerNorm-affine gradients, fp32 -> bf16 casts (weights, residual-stream gradients with
// the dropout-backward mask). All are vectorised (16-byte accesses) and coalesced; none reuses data
// beyond a row, so they use registers rather than shared memory.
//
// Reference arithmetic: nn.LayerNorm(2304) (bridge_module.py:282,288,298 / calls :316,326,331),
// autograd of the same, bias gradients of every nn.Linear, autocast's fp32->bf16 casts.
#include "common.cuh"
#include "launch.h"

namespace b200b {

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, the row cached in registers (NV float4 per lane)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ y,
                                                            float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int rows, int dim,
                                                            float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * dim;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < dim) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / (float)dim;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < dim) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)dim + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  __nv_bfloat16* yr = y + (size_t)row * dim;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < dim) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      uint2 o;
      o.x = pack_bf16((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      o.y = pack_bf16((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      *reinterpret_cast<uint2*>(yr + c) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward (input gradient): dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),
// g = dy * gamma. One warp per row; x and dy cached in registers.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                            const float* __restrict__ x,
                                                            const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in,
                                                            const float* __restrict__ gamma,
                                                            const float* dres, float* dx, int rows, int dim) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * dim;
  const __nv_bfloat16* dyr = dy + (size_t)row * dim;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float4 xh[NV];  // xhat
  float4 g[NV];   // dy * gamma
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < dim) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      const uint2 d = *reinterpret_cast<const uint2*>(dyr + c);
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      g[i] = make_float4(bf16_lo(d.x) * gm.x, bf16_hi(d.x) * gm.y, bf16_lo(d.y) * gm.z, bf16_hi(d.y) * gm.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
  }
  const float m1 = warp_sum(s1) / (float)dim;
  const float m2 = warp_sum(s2) / (float)dim;
  float* dxr = dx + (size_t)row * dim;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < dim) {
      float4 o = make_float4(rstd * (g[i].x - m1 - xh[i].x * m2), rstd * (g[i].y - m1 - xh[i].y * m2),
                             rstd * (g[i].z - m1 - xh[i].z * m2), rstd * (g[i].w - m1 - xh[i].w * m2));
      if (dres != nullptr) {
        const float4 r = *reinterpret_cast<const float4*>(dres + (size_t)row * dim + c);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      *reinterpret_cast<float4*>(dxr + c) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Column reductions over rows of a bf16 matrix dy[rows, cols] (pitch ld):
//   sum[c]   = sum_r dy[r,c]                       (bias gradients, LayerNorm dbeta)
//   gsum[c]  = sum_r dy[r,c] * (x[r,c]-mean[r])*rstd[r]   (LayerNorm dgamma; only if x != NULL)
// Block = 32 column-threads (4 columns each, 8-byte loads -> 256 B per warp row) x 8 row-threads.
// Grid = (col blocks, row chunks); partials go to a workspace and a second kernel sums the chunks
// in a fixed order (deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __nv_bfloat16* __restrict__ dy, long long ld,
                                                             const float* __restrict__ x,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float* __restrict__ psum,
                                                             float* __restrict__ pgsum, int rows, int cols,
                                                             int rows_per_chunk) {
  pdl_prologue_late_trigger();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 4;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), gs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const uint2 d = *reinterpret_cast<const uint2*>(dy + (size_t)r * ld + c);
      const float d0 = bf16_lo(d.x), d1 = bf16_hi(d.x), d2 = bf16_lo(d.y), d3 = bf16_hi(d.y);
      s.x += d0; s.y += d1; s.z += d2; s.w += d3;
      if (x != nullptr) {
        const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)r * cols + c);
        const float m = mean[r], rs = rstd[r];
        gs.x += d0 * (xv.x - m) * rs; gs.y += d1 * (xv.y - m) * rs;
        gs.z += d2 * (xv.z - m) * rs; gs.w += d3 * (xv.w - m) * rs;
      }
    }
  }
  __shared__ float4 sh[2][8][32];
  sh[0][ty][tx] = s;
  sh[1][ty][tx] = gs;
  __syncthreads();
  if (ty == 0 && c < cols) {
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      const float4 a = sh[0][j][tx], b = sh[1][j][tx];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      gs.x += b.x; gs.y += b.y; gs.z += b.z; gs.w += b.w;
    }
    *reinterpret_cast<float4*>(psum + (size_t)blockIdx.y * cols + c) = s;
    if (x != nullptr) *reinterpret_cast<float4*>(pgsum + (size_t)blockIdx.y * cols + c) = gs;
  }
}

__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ psum,
                                                           const float* __restrict__ pgsum, float* __restrict__ out_sum,
                                                           float* __restrict__ out_gsum, int cols, int chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f, g = 0.f;
  for (int j = 0; j < chunks; ++j) {
    s += psum[(size_t)j * cols + c];
    if (pgsum != nullptr) g += pgsum[(size_t)j * cols + c];
  }
  if (out_sum != nullptr) out_sum[c] = s;
  if (out_gsum != nullptr) out_gsum[c] = g;
}

static int colsum_chunks(int rows, int cols, int num_sms) {
  const int col_blocks = (cols + 127) / 128;
  int chunks = (2 * num_sms + col_blocks - 1) / col_blocks;
  const int max_chunks = (rows + 31) / 32;  // at least 32 rows per chunk
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  return chunks;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast of a contiguous range, optionally applying a dropout-backward mask
// (element i of `stream` kept iff forward kept it; kept values are scaled by 1/(1-p) after the
// bf16 rounding, as autograd does on the bf16 gradient).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                        long long n8, DropoutCfg drop_in, uint32_t stream) {
  pdl_prologue_late_trigger();
  const DropoutCfg drop = dropout_resolve(drop_in);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += stride) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * g);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * g + 1);
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (drop.thr != 0) {
      const uint4 bits = dropout_bits8(drop, stream, (uint64_t)g);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        f[i] = dropout_keep(bits, i, drop.thr) ? bf16_round(bf16_round(f[i]) * drop.scale) : 0.0f;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
    o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    reinterpret_cast<uint4*>(out)[g] = o;
  }
}

// ================================================================================================
// CTA-per-row kernels (used by the whole-block entry points). At the bridge's row counts
// (T = 1024..2048) one warp per row leaves ~7 warps per SM and ~1 TB/s; here a CTA of 192 threads
// owns a row (3 float4 per thread at D = 2304), rows are strided over the grid, and everything that
// can be derived from the row while it is in registers is produced in the same pass.
// ================================================================================================
constexpr int kRowThreads = 192;

// sum of (a, b) over the CTA; `slot` alternates per row so one __syncthreads per row suffices
__device__ __forceinline__ float2 block_sum2(float a, float b, float2 (*red)[kRowThreads / 32], int slot) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[slot][warp] = make_float2(a, b);
  __syncthreads();
  float2 t = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < kRowThreads / 32; ++w) {
    const float2 v = red[slot][w];
    t.x += v.x;
    t.y += v.y;
  }
  return t;
}

template <int VPT>
__global__ void __launch_bounds__(kRowThreads) layernorm_fwd_row_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta,
                                                                        __nv_bfloat16* __restrict__ y,
                                                                        float* __restrict__ mean_out,
                                                                        float* __restrict__ rstd_out, int rows,
                                                                        int dim, float eps) {
  pdl_prologue_late_trigger();
  __shared__ float2 red[4][kRowThreads / 32];
  float4 g[VPT], b[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * kRowThreads + threadIdx.x) * 4;
    g[i] = c < dim ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    b[i] = c < dim ? __ldg(reinterpret_cast<const float4*>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int it = 0;
  for (int row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
    const float* xr = x + (size_t)row * dim;
    float4 v[VPT];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * kRowThreads + threadIdx.x) * 4;
      v[i] = c < dim ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = block_sum2(s, 0.f, red, (2 * it) & 3).x / (float)dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * kRowThreads + threadIdx.x) * 4;
      if (c < dim) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rstd = rsqrtf(block_sum2(q, 0.f, red, (2 * it + 1) & 3).x / (float)dim + eps);
    if (threadIdx.x == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    __nv_bfloat16* yr = y + (size_t)row * dim;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * kRowThreads + threadIdx.x) * 4;
      if (c < dim) {
        uint2 o;
        o.x = pack_bf16((v[i].x - mean) * rstd * g[i].x + b[i].x, (v[i].y - mean) * rstd * g[i].y + b[i].y);
        o.y = pack_bf16((v[i].z - mean) * rstd * g[i].z + b[i].z, (v[i].w - mean) * rstd * g[i].w + b[i].w);
        *reinterpret_cast<uint2*>(yr + c) = o;
      }
    }
  }
}

// LayerNorm backward, fused with what follows it in the bridge's backward pass:
//   g = dy*gamma ; dx = dres + rstd*(g - mean(g) - xhat*mean(g*xhat))      (input gradient, fp32)
//   dy_next = bf16(dx)                                                     (operand of the previous sub-layer's GEMMs)
//   partials[cta][0][c] += dy[r,c]            -> LayerNorm dbeta
//   partials[cta][1][c] += dy[r,c]*xhat[r,c]  -> LayerNorm dgamma
//   partials[cta][2][c] += dy_next[r,c]       -> bias gradient of the previous sub-layer's output projection
// dx / dy_next may be NULL (block 0 when the text embeddings need no gradient): only dgamma/dbeta then.
template <int VPT>
struct LnBwdRow {
  float4 x[VPT], r[VPT];
  uint2 d[VPT];
  float mean, rstd;
};

template <int VPT>
__device__ __forceinline__ void ln_bwd_load(LnBwdRow<VPT>& o, const __nv_bfloat16* __restrict__ dy,
                                            const float* __restrict__ x, const float* __restrict__ mean_in,
                                            const float* __restrict__ rstd_in, const float* dres, bool want_res, int row,
                                            int dim) {
  o.mean = mean_in[row];
  o.rstd = rstd_in[row];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * kRowThreads + threadIdx.x) * 4;
    o.x[i] = o.r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    o.d[i] = make_uint2(0u, 0u);
    if (c < dim) {
      o.x[i] = *reinterpret_cast<const float4*>(x + (size_t)row * dim + c);
      o.d[i] = *reinterpret_cast<const uint2*>(dy + (size_t)row * dim + c);
      if (want_res) o.r[i] = *reinterpret_cast<const float4*>(dres + (size_t)row * dim + c);
    }
  }
}

template <int VPT>
__global__ void __launch_bounds__(kRowThreads) layernorm_bwd_row_kernel(
    const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_in,
    const float* __restrict__ rstd_in, const float* __restrict__ gamma, const float* dres, float* dx,
    __nv_bfloat16* __restrict__ dy_next, float* __restrict__ partials, int rows, int dim) {
  pdl_prologue_late_trigger();
  __shared__ float2 red[2][kRowThreads / 32];
  float4 gm[VPT], pb[VPT], pg[VPT], pc[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * kRowThreads + threadIdx.x) * 4;
    gm[i] = c < dim ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    pb[i] = pg[i] = pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bool want_res = dx != nullptr && dres != nullptr;
  // The next row's operands are in flight while this row is reduced and written. dres may alias dx
  // (in-place residual gradient): a row is only ever read and written by the CTA that owns it, and
  // its loads complete before its stores are issued.
  LnBwdRow<VPT> cur, nxt;
  int row = blockIdx.x;
  if (row < rows) ln_bwd_load<VPT>(cur, dy, x, mean_in, rstd_in, dres, want_res, row, dim);
  int it = 0;
  for (; row < rows; row += gridDim.x, ++it) {
    const int nrow = row + gridDim.x;
    if (nrow < rows) ln_bwd_load<VPT>(nxt, dy, x, mean_in, rstd_in, dres, want_res, nrow, dim);
    const float mean = cur.mean, rstd = cur.rstd;
    float4 xh[VPT], g[VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * kRowThreads + threadIdx.x) * 4;
      xh[i] = g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < dim) {
        const float4 xv = cur.x[i];
        const uint2 d = cur.d[i];
        const float d0 = bf16_lo(d.x), d1 = bf16_hi(d.x), d2 = bf16_lo(d.y), d3 = bf16_hi(d.y);
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        g[i] = make_float4(d0 * gm[i].x, d1 * gm[i].y, d2 * gm[i].z, d3 * gm[i].w);
        pb[i].x += d0; pb[i].y += d1; pb[i].z += d2; pb[i].w += d3;
        pg[i].x += d0 * xh[i].x; pg[i].y += d1 * xh[i].y; pg[i].z += d2 * xh[i].z; pg[i].w += d3 * xh[i].w;
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      }
    }
    if (dx != nullptr) {  // uniform
      const float2 t = block_sum2(s1, s2, red, it & 1);
      const float m1 = t.x / (float)dim, m2 = t.y / (float)dim;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int c = (i * kRowThreads + threadIdx.x) * 4;
        if (c < dim) {
          const float4 r = cur.r[i];
          const float4 o = make_float4(r.x + rstd * (g[i].x - m1 - xh[i].x * m2), r.y + rstd * (g[i].y - m1 - xh[i].y * m2),
                                       r.z + rstd * (g[i].z - m1 - xh[i].z * m2), r.w + rstd * (g[i].w - m1 - xh[i].w * m2));
          *reinterpret_cast<float4*>(dx + (size_t)row * dim + c) = o;
          if (dy_next != nullptr) {
            uint2 ob;
            ob.x = pack_bf16(o.x, o.y);
            ob.y = pack_bf16(o.z, o.w);
            *reinterpret_cast<uint2*>(dy_next + (size_t)row * dim + c) = ob;
            pc[i].x += bf16_lo(ob.x); pc[i].y += bf16_hi(ob.x); pc[i].z += bf16_lo(ob.y); pc[i].w += bf16_hi(ob.y);
          }
        }
      }
    }
    cur = nxt;
  }
  float* pr = partials + (size_t)blockIdx.x * 3 * dim;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * kRowThreads + threadIdx.x) * 4;
    if (c < dim) {
      *reinterpret_cast<float4*>(pr + c) = pb[i];
      *reinterpret_cast<float4*>(pr + dim + c) = pg[i];
      *reinterpret_cast<float4*>(pr + 2 * dim + c) = pc[i];
    }
  }
}

// fp32 -> bf16 cast of a [rows, dim] matrix with the dropout-backward mask of `stream`
// (same element indexing as cast_bf16_kernel) fused with the column sums of the bf16 result:
// partials[cta][c] = sum over this CTA's rows. Thread = one 8-element dropout group per row pass.
template <int GPT>
__global__ void __launch_bounds__(288) cast_colsum_row_kernel(const float* __restrict__ in,
                                                              __nv_bfloat16* __restrict__ out,
                                                              float* __restrict__ partials, int rows, int dim,
                                                              DropoutCfg drop_in, uint32_t stream) {
  pdl_prologue_late_trigger();
  const DropoutCfg drop = dropout_resolve(drop_in);
  float ps[GPT][8];
#pragma unroll
  for (int i = 0; i < GPT; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) ps[i][e] = 0.f;
  const int groups = dim >> 3;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
#pragma unroll
    for (int i = 0; i < GPT; ++i) {
      const int gcol = i * 288 + threadIdx.x;
      if (gcol < groups) {
        const size_t gidx = (size_t)row * groups + gcol;
        const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * gidx);
        const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * gidx + 1);
        float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (drop.thr != 0) {
          const uint4 bits = dropout_bits8(drop, stream, (uint64_t)gidx);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = dropout_keep(bits, e, drop.thr) ? bf16_round(bf16_round(f[e]) * drop.scale) : 0.0f;
        }
        uint4 o;
        o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
        o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
        reinterpret_cast<uint4*>(out)[gidx] = o;
        ps[i][0] += bf16_lo(o.x); ps[i][1] += bf16_hi(o.x); ps[i][2] += bf16_lo(o.y); ps[i][3] += bf16_hi(o.y);
        ps[i][4] += bf16_lo(o.z); ps[i][5] += bf16_hi(o.z); ps[i][6] += bf16_lo(o.w); ps[i][7] += bf16_hi(o.w);
      }
    }
  }
  float* pr = partials + (size_t)blockIdx.x * dim;
#pragma unroll
  for (int i = 0; i < GPT; ++i) {
    const int gcol = i * 288 + threadIdx.x;
    if (gcol < groups) {
      *reinterpret_cast<float4*>(pr + 8 * gcol) = make_float4(ps[i][0], ps[i][1], ps[i][2], ps[i][3]);
      *reinterpret_cast<float4*>(pr + 8 * gcol + 4) = make_float4(ps[i][4], ps[i][5], ps[i][6], ps[i][7]);
    }
  }
}

// One launch that finishes any number of partial column-sum sets:
//   out[c] = sum_{j < chunks} partials[j * chunk_stride + c],  fixed order (deterministic).
struct FinalizeTasks {
  b200b_colsum_task t[B200B_MAX_COLSUM_TASKS];
  int n;
};
// Block = 128 columns (32 lanes x float4) x 8 chunk-lanes: lane ty sums chunks ty, ty+8, ... with 8
// independent 16-byte loads in flight, then the 8 lanes are combined through shared memory in a
// fixed order (deterministic).
__global__ void __launch_bounds__(256) colsum_finalize_multi_kernel(const FinalizeTasks tasks) {
  pdl_prologue_late_trigger();
  const b200b_colsum_task& t = tasks.t[blockIdx.y];
  if (blockIdx.x * 128 >= t.cols) return;  // whole block out of range for this (shorter) task
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + 4 * tx;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < t.cols) {
    const float* base = t.partials + c;
    int j = ty;
    for (; j + 56 < t.chunks; j += 64) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(base + (size_t)(j + 8 * u) * t.chunk_stride));
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; j < t.chunks; j += 8) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(base + (size_t)j * t.chunk_stride));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  __shared__ float4 sh[8][32];
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < t.cols) {
    float4 r = sh[0][tx];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 v = sh[k][tx]; r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w; }
    *reinterpret_cast<float4*>(t.out + c) = r;
  }
}

// out f32 = scale * in bf16 (exchanged gradient bucket -> .grad)
__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const uint4* __restrict__ in, float4* __restrict__ out,
                                                          long long n8, float scale) {
  pdl_prologue_late_trigger();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += stride) {
    const uint4 v = __ldcs(in + g);
    __stcs(out + 2 * g, make_float4(bf16_lo(v.x) * scale, bf16_hi(v.x) * scale, bf16_lo(v.y) * scale, bf16_hi(v.y) * scale));
    __stcs(out + 2 * g + 1, make_float4(bf16_lo(v.z) * scale, bf16_hi(v.z) * scale, bf16_lo(v.w) * scale, bf16_hi(v.w) * scale));
  }
}

template <typename K, typename... Args>
static int launch_rows(K kern, int rows, cudaStream_t stream, const char* what, Args... args) {
  const int warps = 8;
  const int grid = (rows + warps - 1) / warps;
  kern<<<grid, warps * 32, 0, stream>>>(args...);
  return check_launch(what, stream);
}

}  // namespace b200b

using namespace b200b;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#define B200B_DISPATCH_NV(dim, CALL)                                   \
  do {                                                                 \
    const int nv_ = ((dim) + 127) / 128;                               \
    if (nv_ <= 1) { CALL(1); }                                         \
    else if (nv_ <= 2) { CALL(2); }                                    \
    else if (nv_ <= 4) { CALL(4); }                                    \
    else if (nv_ <= 8) { CALL(8); }                                    \
    else if (nv_ <= 18) { CALL(18); }                                  \
    else if (nv_ <= 32) { CALL(32); }                                  \
    else {                                                             \
      set_last_error("layernorm: dim %d > 4096 not supported", (dim)); \
      return B200B_ERR_SHAPE;                                          \
    }                                                                  \
  } while (0)

extern "C" int b200b_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* mean,
                                   float* rstd, int rows, int dim, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y_bf16 || !mean || !rstd) {
    set_last_error("layernorm_fwd: null argument");
    return B200B_ERR_ARG;
  }
  if (rows <= 0 || dim <= 0 || (dim % 4) != 0) {
    set_last_error("layernorm_fwd: need rows > 0 and dim %% 4 == 0 (rows=%d dim=%d)", rows, dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16(x) || !al16(gamma) || !al16(beta) || (reinterpret_cast<uintptr_t>(y_bf16) & 7)) {
    set_last_error("layernorm_fwd: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
#define CALL(NV)                                                                                              \
  return launch_rows(layernorm_fwd_kernel<NV>, rows, stream, "layernorm_fwd", x, gamma, beta,                 \
                     reinterpret_cast<__nv_bfloat16*>(y_bf16), mean, rstd, rows, dim, eps)
  B200B_DISPATCH_NV(dim, CALL);
#undef CALL
  return B200B_OK;
}

extern "C" int b200b_layernorm_bwd(const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                                   const float* gamma, const float* dres, float* dx, int rows, int dim,
                                   void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dy_bf16 || !x || !mean || !rstd || !gamma || !dx) {
    set_last_error("layernorm_bwd: null argument");
    return B200B_ERR_ARG;
  }
  if (rows <= 0 || dim <= 0 || (dim % 4) != 0) {
    set_last_error("layernorm_bwd: need rows > 0 and dim %% 4 == 0 (rows=%d dim=%d)", rows, dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16(x) || !al16(gamma) || !al16(dx) || (dres && !al16(dres)) || (reinterpret_cast<uintptr_t>(dy_bf16) & 7)) {
    set_last_error("layernorm_bwd: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
#define CALL(NV)                                                                                              \
  return launch_rows(layernorm_bwd_kernel<NV>, rows, stream, "layernorm_bwd",                                 \
                     reinterpret_cast<const __nv_bfloat16*>(dy_bf16), x, mean, rstd, gamma, dres, dx, rows, dim)
  B200B_DISPATCH_NV(dim, CALL);
#undef CALL
  return B200B_OK;
}

extern "C" size_t b200b_colsum_workspace_bytes(int rows, int cols) {
  // sized for the largest chunk count colsum_chunks() can pick (64), both partial arrays
  (void)rows;
  return (size_t)2 * 64 * (size_t)cols * sizeof(float);
}

extern "C" int b200b_colsum(const void* dy_bf16, int64_t ld, const float* x, const float* mean, const float* rstd,
                            float* out_sum, float* out_gsum, int rows, int cols, void* workspace, size_t ws_bytes,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dy_bf16 || (!out_sum && !out_gsum) || !workspace) {
    set_last_error("colsum: null argument");
    return B200B_ERR_ARG;
  }
  if (x != nullptr && (!mean || !rstd || !out_gsum)) {
    set_last_error("colsum: LayerNorm mode needs x, mean, rstd and out_gsum");
    return B200B_ERR_ARG;
  }
  if (rows <= 0 || cols <= 0 || (cols % 4) != 0 || (ld % 4) != 0) {
    set_last_error("colsum: need rows > 0, cols %% 4 == 0, ld %% 4 == 0 (rows=%d cols=%d ld=%lld)", rows, cols,
                   (long long)ld);
    return B200B_ERR_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(dy_bf16) & 7) || (x && !al16(x)) || !al16(workspace)) {
    set_last_error("colsum: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc != B200B_OK) return rc;
  const int chunks = colsum_chunks(rows, cols, num_sms);
  if (ws_bytes < (size_t)2 * chunks * cols * sizeof(float)) {
    set_last_error("colsum: workspace too small (%zu < %zu)", ws_bytes, (size_t)2 * chunks * cols * sizeof(float));
    return B200B_ERR_WORKSPACE;
  }
  float* psum = reinterpret_cast<float*>(workspace);
  float* pgsum = psum + (size_t)chunks * cols;
  const int rows_per_chunk = (rows + chunks - 1) / chunks;
  dim3 grid((cols + 127) / 128, chunks);
  launch_pdl(kPdlRows, colsum_partial_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), (long long)ld, x,
                                                  mean, rstd, psum, pgsum, rows, cols, rows_per_chunk);
  rc = check_launch("colsum_partial", stream);
  if (rc != B200B_OK) return rc;
  colsum_final_kernel<<<(cols + 255) / 256, 256, 0, stream>>>(psum, x ? pgsum : nullptr, out_sum, out_gsum, cols,
                                                              chunks);
  return check_launch("colsum_final", stream);
}

extern "C" int b200b_cast_bf16(const float* in, void* out_bf16, int64_t n, float dropout_p, uint64_t seed,
                               uint32_t dropout_stream, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!in || !out_bf16) {
    set_last_error("cast_bf16: null argument");
    return B200B_ERR_ARG;
  }
  if (n <= 0 || (n % 8) != 0) {
    set_last_error("cast_bf16: n must be a positive multiple of 8 (n=%lld)", (long long)n);
    return B200B_ERR_SHAPE;
  }
  if (!al16(in) || !al16(out_bf16)) {
    set_last_error("cast_bf16: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
  if (!(dropout_p >= 0.0f && dropout_p < 1.0f)) {
    set_last_error("cast_bf16: dropout_p must be in [0,1)");
    return B200B_ERR_ARG;
  }
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc != B200B_OK) return rc;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)num_sms * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  const DropoutCfg dc = make_dropout_cfg(dropout_p, seed, &dropout_stream);
  launch_pdl(kPdlRows, cast_bf16_kernel, dim3((int)blocks), dim3(256), 0, stream, in, reinterpret_cast<__nv_bfloat16*>(out_bf16), n8, dc,
                                                    dropout_stream);
  return check_launch("cast_bf16", stream);
}

// ------------------------------------------------------------------------------------------------
// CTA-per-row entry points
// ------------------------------------------------------------------------------------------------
#define B200B_DISPATCH_VPT(dim, CALL)                                        \
  do {                                                                       \
    const int v_ = ((dim) / 4 + kRowThreads - 1) / kRowThreads;              \
    if (v_ <= 1) { CALL(1); }                                                \
    else if (v_ <= 2) { CALL(2); }                                           \
    else if (v_ <= 3) { CALL(3); }                                           \
    else if (v_ <= 6) { CALL(6); }                                           \
    else {                                                                   \
      set_last_error("row kernels: dim %d > 4608 not supported", (dim));     \
      return B200B_ERR_SHAPE;                                                \
    }                                                                        \
  } while (0)

extern "C" int b200b_row_chunks(int rows) {
  int num_sms = 0;
  if (device_sm_count(&num_sms) != B200B_OK) return 0;
  const int cap = 2 * num_sms;
  return rows < cap ? (rows > 0 ? rows : 0) : cap;
}

extern "C" int b200b_layernorm_fwd_rows(const float* x, const float* gamma, const float* beta, void* y_bf16,
                                        float* mean, float* rstd, int rows, int dim, float eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y_bf16 || !mean || !rstd) {
    set_last_error("layernorm_fwd_rows: null argument");
    return B200B_ERR_ARG;
  }
  if (rows <= 0 || dim <= 0 || (dim % 4) != 0) {
    set_last_error("layernorm_fwd_rows: need rows > 0 and dim %% 4 == 0 (rows=%d dim=%d)", rows, dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16(x) || !al16(gamma) || !al16(beta) || (reinterpret_cast<uintptr_t>(y_bf16) & 7)) {
    set_last_error("layernorm_fwd_rows: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc != B200B_OK) return rc;
  const int grid = rows < 8 * num_sms ? rows : 8 * num_sms;
#define CALL(V)                                                                                                   \
  launch_pdl(kPdlRows, layernorm_fwd_row_kernel<V>, dim3(grid), dim3(kRowThreads), 0, stream, x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y_bf16), \
                                                                mean, rstd, rows, dim, eps)
  B200B_DISPATCH_VPT(dim, CALL);
#undef CALL
  return check_launch("layernorm_fwd", stream);
}

extern "C" int b200b_layernorm_bwd_fused(const void* dy_bf16, const float* x, const float* mean, const float* rstd,
                                         const float* gamma, const float* dres, float* dx, void* dy_next_bf16,
                                         float* partials, int rows, int dim, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dy_bf16 || !x || !mean || !rstd || !gamma || !partials) {
    set_last_error("layernorm_bwd_fused: null argument");
    return B200B_ERR_ARG;
  }
  if (dx == nullptr && dy_next_bf16 != nullptr) {
    set_last_error("layernorm_bwd_fused: dy_next needs dx");
    return B200B_ERR_ARG;
  }
  if (rows <= 0 || dim <= 0 || (dim % 4) != 0) {
    set_last_error("layernorm_bwd_fused: need rows > 0 and dim %% 4 == 0 (rows=%d dim=%d)", rows, dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16(x) || !al16(gamma) || (dx && !al16(dx)) || (dres && !al16(dres)) || !al16(partials) ||
      (reinterpret_cast<uintptr_t>(dy_bf16) & 7) || (reinterpret_cast<uintptr_t>(dy_next_bf16) & 7)) {
    set_last_error("layernorm_bwd_fused: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
  const int grid = b200b_row_chunks(rows);
  if (grid <= 0) return B200B_ERR_DEVICE;
#define CALL(V)                                                                                              \
  launch_pdl(kPdlRows, layernorm_bwd_row_kernel<V>, dim3(grid), dim3(kRowThreads), 0, stream, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), x, mean, \
                                                                rstd, gamma, dres, dx,                       \
                                                                reinterpret_cast<__nv_bfloat16*>(dy_next_bf16), partials, rows, dim)
  B200B_DISPATCH_VPT(dim, CALL);
#undef CALL
  return check_launch("layernorm_bwd", stream);
}

extern "C" int b200b_cast_bf16_colsum(const float* in, void* out_bf16, float* partials, int rows, int dim,
                                      float dropout_p, uint64_t seed, uint32_t dropout_stream, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!in || !out_bf16 || !partials) {
    set_last_error("cast_bf16_colsum: null argument");
    return B200B_ERR_ARG;
  }
  if (rows <= 0 || dim <= 0 || (dim % 8) != 0 || dim > 8 * 288 * 4) {
    set_last_error("cast_bf16_colsum: need rows > 0, dim %% 8 == 0, dim <= 9216 (rows=%d dim=%d)", rows, dim);
    return B200B_ERR_SHAPE;
  }
  if (!al16(in) || !al16(out_bf16) || !al16(partials)) {
    set_last_error("cast_bf16_colsum: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
  if (!(dropout_p >= 0.0f && dropout_p < 1.0f)) {
    set_last_error("cast_bf16_colsum: dropout_p must be in [0,1)");
    return B200B_ERR_ARG;
  }
  const int grid = b200b_row_chunks(rows);
  if (grid <= 0) return B200B_ERR_DEVICE;
  const DropoutCfg dc = make_dropout_cfg(dropout_p, seed, &dropout_stream);
  const int gpt = (dim / 8 + 287) / 288;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (gpt <= 1) launch_pdl(kPdlRows, cast_colsum_row_kernel<1>, dim3(grid), dim3(288), 0, stream, in, o, partials, rows, dim, dc, dropout_stream);
  else if (gpt <= 2) launch_pdl(kPdlRows, cast_colsum_row_kernel<2>, dim3(grid), dim3(288), 0, stream, in, o, partials, rows, dim, dc, dropout_stream);
  else launch_pdl(kPdlRows, cast_colsum_row_kernel<4>, dim3(grid), dim3(288), 0, stream, in, o, partials, rows, dim, dc, dropout_stream);
  return check_launch("cast_colsum", stream);
}

extern "C" int b200b_colsum_partials(const void* dy_bf16, int64_t ld, int rows, int cols, float* partials,
                                     int* chunks_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dy_bf16 || !partials || !chunks_out) {
    set_last_error("colsum_partials: null argument");
    return B200B_ERR_ARG;
  }
  if (rows <= 0 || cols <= 0 || (cols % 4) != 0 || (ld % 4) != 0) {
    set_last_error("colsum_partials: need rows > 0, cols %% 4 == 0, ld %% 4 == 0");
    return B200B_ERR_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(dy_bf16) & 7) || !al16(partials)) {
    set_last_error("colsum_partials: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc != B200B_OK) return rc;
  const int chunks = colsum_chunks(rows, cols, num_sms);
  const int rows_per_chunk = (rows + chunks - 1) / chunks;
  dim3 grid((cols + 127) / 128, chunks);
  launch_pdl(kPdlRows, colsum_partial_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), (long long)ld, nullptr,
                                                  nullptr, nullptr, partials, nullptr, rows, cols, rows_per_chunk);
  *chunks_out = chunks;
  return check_launch("colsum_partial", stream);
}

extern "C" int b200b_colsum_finalize(const b200b_colsum_task* tasks, int ntasks, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!tasks || ntasks <= 0 || ntasks > B200B_MAX_COLSUM_TASKS) {
    set_last_error("colsum_finalize: need 1..%d tasks", B200B_MAX_COLSUM_TASKS);
    return B200B_ERR_ARG;
  }
  FinalizeTasks ft;
  ft.n = ntasks;
  int max_cols = 0;
  for (int i = 0; i < ntasks; ++i) {
    if (!tasks[i].partials || !tasks[i].out || tasks[i].cols <= 0 || tasks[i].chunks <= 0 || (tasks[i].cols % 4) ||
        (tasks[i].chunk_stride % 4) || !al16(tasks[i].partials) || !al16(tasks[i].out)) {
      set_last_error("colsum_finalize: bad task %d", i);
      return B200B_ERR_ARG;
    }
    ft.t[i] = tasks[i];
    if (tasks[i].cols > max_cols) max_cols = tasks[i].cols;
  }
  dim3 grid((max_cols + 127) / 128, ntasks);
  launch_pdl(kPdlRows, colsum_finalize_multi_kernel, dim3(grid), dim3(256), 0, stream, ft);
  return check_launch("colsum_final", stream);
}

extern "C" int b200b_bf16_to_f32(const void* in_bf16, float* out, int64_t n, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!in_bf16 || !out) {
    set_last_error("bf16_to_f32: null argument");
    return B200B_ERR_ARG;
  }
  if (n <= 0 || (n % 8) != 0) {
    set_last_error("bf16_to_f32: n must be a positive multiple of 8 (n=%lld)", (long long)n);
    return B200B_ERR_SHAPE;
  }
  if (!al16(in_bf16) || !al16(out)) {
    set_last_error("bf16_to_f32: misaligned pointer");
    return B200B_ERR_ALIGN;
  }
  int num_sms = 0;
  int rc = device_sm_count(&num_sms);
  if (rc != B200B_OK) return rc;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)num_sms * 8;
  if (blocks > cap) blocks = cap;
  launch_pdl(kPdlRows, bf16_to_f32_kernel, dim3((int)blocks), dim3(256), 0, stream, reinterpret_cast<const uint4*>(in_bf16),
                                                      reinterpret_cast<float4*>(out), n8, scale);
  return check_launch("bf16_to_f32", stream);
}

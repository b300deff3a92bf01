// Whole-block forward / backward of the bridge: the host-side sequencing of the kernels in
// gemm_sm100.cu, attention.cu and elementwise.cu for one BridgeBlock
// (reference: BridgeBlock.forward, bridge_module.py:300-335, and its autograd backward).
// Everything is enqueued on the caller's stream; nothing here allocates or synchronises.
#include <string.h>

#include "common.cuh"
#include "launch.h"

namespace b200b {

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += al256(count * sizeof(T));
    return p;
  }
};

struct Dims {
  int B, L, Nv, D, Dv, F, Hc, Hs, nb, flags;
  size_t T, Tv;
  int dc, ds;  // head dims
};

static int read_dims(const b200b_bridge_dims* d, Dims* o, const char* what) {
  if (d == nullptr) {
    set_last_error("%s: null dims", what);
    return B200B_ERR_ARG;
  }
  o->B = d->batch; o->L = d->len_text; o->Nv = d->len_vision;
  o->D = d->dim; o->Dv = d->dim_vision; o->F = d->dim_ffn;
  o->Hc = d->heads_cross; o->Hs = d->heads_self; o->nb = d->num_blocks; o->flags = d->flags;
  if (o->B <= 0 || o->L <= 0 || o->Nv <= 0 || o->D <= 0 || o->Dv <= 0 || o->F <= 0 || o->Hc <= 0 || o->Hs <= 0 ||
      o->nb <= 0) {
    set_last_error("%s: all dims must be positive", what);
    return B200B_ERR_SHAPE;
  }
  if ((o->D % 8) || (o->Dv % 8) || (o->F % 8) || (o->D % o->Hc) || (o->D % o->Hs)) {
    set_last_error("%s: dim/dim_vision/dim_ffn must be multiples of 8 and dim divisible by the head counts", what);
    return B200B_ERR_SHAPE;
  }
  o->T = (size_t)o->B * o->L;
  o->Tv = (size_t)o->B * o->Nv;
  o->dc = o->D / o->Hc;
  o->ds = o->D / o->Hs;
  if (o->T > 0x7fffffff || o->Tv > 0x7fffffff) {
    set_last_error("%s: too many rows", what);
    return B200B_ERR_SHAPE;
  }
  return B200B_OK;
}

// activations a block keeps for its backward
struct Saved {
  __nv_bfloat16 *xn1, *q, *o1, *xn2, *qkv, *o2, *xn3, *u, *h;
  float *x1, *x2;
  float *mean1, *rstd1, *mean2, *rstd2, *mean3, *rstd3;
  float *lse1, *lse2;
  size_t bytes;
};

static Saved carve_saved(const Dims& d, void* base) {
  Carver c(base);
  Saved s;
  const size_t TD = d.T * d.D, TF = d.T * d.F;
  s.xn1 = c.take<__nv_bfloat16>(TD);
  s.q = c.take<__nv_bfloat16>(TD);
  s.o1 = c.take<__nv_bfloat16>(TD);
  s.xn2 = c.take<__nv_bfloat16>(TD);
  s.qkv = c.take<__nv_bfloat16>(3 * TD);
  s.o2 = c.take<__nv_bfloat16>(TD);
  s.xn3 = c.take<__nv_bfloat16>(TD);
  s.u = c.take<__nv_bfloat16>(TF);
  s.h = c.take<__nv_bfloat16>(TF);
  s.x1 = c.take<float>(TD);
  s.x2 = c.take<float>(TD);
  s.mean1 = c.take<float>(d.T); s.rstd1 = c.take<float>(d.T);
  s.mean2 = c.take<float>(d.T); s.rstd2 = c.take<float>(d.T);
  s.mean3 = c.take<float>(d.T); s.rstd3 = c.take<float>(d.T);
  s.lse1 = c.take<float>((size_t)d.B * d.Hc * d.L);
  s.lse2 = c.take<float>((size_t)d.B * d.Hs * d.L);
  s.bytes = c.off;
  return s;
}

// transient buffers of the backward pass
constexpr int kMaxRowChunks = 320;  // >= 2 x SM count: upper bound of b200b_row_chunks()
struct BwdWs {
  __nv_bfloat16 *dy, *du, *dxn, *dattn, *dqkv;
  float* dx;
  uint8_t* attn_ws; size_t attn_ws_bytes;
  // per-CTA partial column sums, finished by ONE b200b_colsum_finalize launch per block
  float *p_cast, *p_du, *p_lnf, *p_dqkv, *p_lns, *p_dq, *p_lnc, *p_dkv;
  size_t bytes;
};

static BwdWs carve_bwd(const Dims& d, void* base) {
  Carver c(base);
  BwdWs w;
  const size_t TD = d.T * d.D, TF = d.T * d.F;
  w.dy = c.take<__nv_bfloat16>(TD);
  w.du = c.take<__nv_bfloat16>(TF);
  w.dxn = c.take<__nv_bfloat16>(TD);
  w.dattn = c.take<__nv_bfloat16>(TD);
  w.dqkv = c.take<__nv_bfloat16>(3 * TD);
  w.dx = c.take<float>(TD);
  const size_t a1 = b200b_attention_bwd_workspace_bytes(d.B, d.Hc, d.L, d.Nv);
  const size_t a2 = b200b_attention_bwd_workspace_bytes(d.B, d.Hs, d.L, d.L);
  w.attn_ws_bytes = a1 > a2 ? a1 : a2;
  w.attn_ws = c.take<uint8_t>(w.attn_ws_bytes);
  const size_t rc = d.T < (size_t)kMaxRowChunks ? d.T : (size_t)kMaxRowChunks;  // row-kernel chunks
  const size_t D = d.D;
  w.p_cast = c.take<float>(rc * D);
  w.p_du = c.take<float>((size_t)64 * d.F);
  w.p_lnf = c.take<float>(rc * 3 * D);
  w.p_dqkv = c.take<float>((size_t)64 * 3 * D);
  w.p_lns = c.take<float>(rc * 3 * D);
  w.p_dq = c.take<float>((size_t)64 * D);
  w.p_lnc = c.take<float>(rc * 3 * D);
  w.p_dkv = c.take<float>((size_t)64 * 2 * D * d.nb);
  w.bytes = c.off;
  return w;
}

// list of partial column-sum sets to be finished at the end of a block's backward
struct Finalizer {
  b200b_colsum_task t[B200B_MAX_COLSUM_TASKS];
  int n = 0;
  void add(const float* partials, float* out, int cols, int chunks, long long stride) {
    t[n].partials = partials; t[n].out = out; t[n].cols = cols; t[n].chunks = chunks; t[n].chunk_stride = stride;
    ++n;
  }
};

#define B200B_TRY(expr)        \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != B200B_OK) return rc_; \
  } while (0)

static int gemm(const void* a, int a_major, long long lda, const void* b, int b_major, long long ldb, int m, int n,
                int k, int epi, void* out, long long ldo, const float* bias, const float* resid, void* aux,
                long long ldaux, float p, uint64_t seed, uint32_t dstream, cudaStream_t st) {
  b200b_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = a; g.b = b; g.a_major = a_major; g.b_major = b_major;
  g.m = m; g.n = n; g.k = k; g.lda = lda; g.ldb = ldb;
  g.epilogue = epi; g.block_n = 0;
  g.out = out; g.ldo = ldo; g.aux = aux; g.ldaux = ldaux;
  g.bias = bias; g.resid = resid; g.ldr = ldo;
  g.beta = 0.f; g.dropout_p = p; g.seed = seed; g.dropout_stream = dstream;
  return b200b_gemm(&g, st);
}

// dX = dY W (bf16 [rows, n_in]) and dW = dY^T X ([n_out, n_in], fp32 or bf16 by `wgrad_epi`) of one Linear as one grouped
// launch where both fit 256 x 128 pair tiles (b200b_gemm_dual), else as two launches
static int grad_pair(const void* dy, long long lddy, const void* w, const void* x, long long ldx, int rows, int n_out,
                     int n_in, void* dx, long long lddx, void* dw, int wgrad_epi, cudaStream_t st) {
  b200b_gemm_args dg, wg;
  memset(&dg, 0, sizeof(dg));
  memset(&wg, 0, sizeof(wg));
  dg.a = dy; dg.a_major = 0; dg.lda = lddy;            // A = dY [rows, n_out]
  dg.b = w; dg.b_major = 1; dg.ldb = n_in;             // B = W stored [n_out (k), n_in (n)]
  dg.m = rows; dg.n = n_in; dg.k = n_out;
  dg.epilogue = B200B_EPI_BF16_BIAS; dg.out = dx; dg.ldo = lddx;
  wg.a = dy; wg.a_major = 1; wg.lda = lddy;            // A = dY stored [rows (k), n_out (m)]
  wg.b = x; wg.b_major = 1; wg.ldb = ldx;              // B = X stored [rows (k), n_in (n)]
  wg.m = n_out; wg.n = n_in; wg.k = rows;
  wg.epilogue = wgrad_epi; wg.out = dw; wg.ldo = n_in;
  return b200b_gemm_dual(&dg, &wg, st);
}

static int attn(bool bwd, const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                void* o, long long ldo, float* lse, const void* d_o, void* dq, long long lddq, void* dk, long long lddk,
                void* dv, long long lddv, void* ws, size_t ws_bytes, int B, int H, int Lq, int Lk, int hd, float p,
                uint64_t seed, uint32_t dstream, cudaStream_t st) {
  b200b_attn_args a;
  memset(&a, 0, sizeof(a));
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.o = o; a.ldo = ldo; a.lse = lse;
  a.d_o = d_o; a.lddo = ldo; a.dq = dq; a.lddq = lddq; a.dk = dk; a.lddk = lddk; a.dv = dv; a.lddv = lddv;
  a.workspace = ws; a.workspace_bytes = ws_bytes;
  a.batch = B; a.heads = H; a.len_q = Lq; a.len_k = Lk; a.head_dim = hd;
  a.dropout_p = p; a.seed = seed; a.dropout_stream = dstream;
  return bwd ? b200b_attention_bwd(&a, st) : b200b_attention_fwd(&a, st);
}

// dropout stream ids of block i
static inline uint32_t ds_cross(int i) { return 16u * i + 0; }
static inline uint32_t ds_self(int i) { return 16u * i + 1; }
static inline uint32_t ds_ffn_h(int i) { return 16u * i + 2; }
static inline uint32_t ds_ffn_o(int i) { return 16u * i + 3; }

}  // namespace b200b

using namespace b200b;

extern "C" size_t b200b_bridge_block_saved_bytes(const b200b_bridge_dims* dims) {
  Dims d;
  if (read_dims(dims, &d, "block_saved_bytes") != B200B_OK) return 0;
  return carve_saved(d, nullptr).bytes;
}

extern "C" size_t b200b_bridge_backward_workspace_bytes(const b200b_bridge_dims* dims) {
  Dims d;
  if (read_dims(dims, &d, "backward_workspace_bytes") != B200B_OK) return 0;
  return carve_bwd(d, nullptr).bytes;
}

extern "C" int b200b_bridge_kv_project(const b200b_bridge_dims* dims, const float* vision_f32, const void* wkv_all,
                                       const float* bkv_all, void* vision_bf16, void* kv, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  Dims d;
  B200B_TRY(read_dims(dims, &d, "kv_project"));
  if (!vision_f32 || !wkv_all || !bkv_all || !vision_bf16 || !kv) {
    set_last_error("kv_project: null argument");
    return B200B_ERR_ARG;
  }
  const int n = 2 * d.D * d.nb;
  B200B_TRY(b200b_cast_bf16(vision_f32, vision_bf16, (int64_t)d.Tv * d.Dv, 0.f, 0, 0, st));
  return gemm(vision_bf16, 0, d.Dv, wkv_all, 0, d.Dv, (int)d.Tv, n, d.Dv, B200B_EPI_BF16_BIAS, kv, n, bkv_all, nullptr,
              nullptr, 0, 0.f, 0, 0, st);
}

extern "C" int b200b_bridge_block_forward(const b200b_bridge_dims* dims, int i, const b200b_block_weights* w,
                                          const float* x_in, const void* kv, float* x_out, void* saved,
                                          size_t saved_bytes, float p, uint64_t seed, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  Dims d;
  B200B_TRY(read_dims(dims, &d, "block_forward"));
  if (!w || !x_in || !kv || !x_out || !saved || i < 0 || i >= d.nb) {
    set_last_error("block_forward: null argument or bad block index");
    return B200B_ERR_ARG;
  }
  Saved s = carve_saved(d, saved);
  if (saved_bytes < s.bytes) {
    set_last_error("block_forward: saved arena too small (%zu < %zu)", saved_bytes, s.bytes);
    return B200B_ERR_WORKSPACE;
  }
  const int T = (int)d.T, D = d.D, F = d.F;
  // B200B_BRIDGE_SEED_INDIRECT: `seed` holds a device pointer; every kernel reads the seed itself
  const uint32_t ind = (d.flags & B200B_BRIDGE_SEED_INDIRECT) ? B200B_SEED_INDIRECT : 0u;
  const long long ldkv = 2LL * D * d.nb;
  const __nv_bfloat16* kblk = reinterpret_cast<const __nv_bfloat16*>(kv) + (size_t)2 * D * i;
  const float eps = 1e-5f;

  // B200B_BRIDGE_PART_*: decode runs the position-independent cross-attention sub-layer once per text
  // position and the rest of the block on the whole prefix
  const bool do_cross = !(d.flags & B200B_BRIDGE_PART_REST), do_rest = !(d.flags & B200B_BRIDGE_PART_CROSS);
  if (!do_cross && !do_rest) {
    set_last_error("block_forward: PART_CROSS and PART_REST are exclusive");
    return B200B_ERR_ARG;
  }
  if ((!do_cross || !do_rest) && p != 0.f) {
    set_last_error("block_forward: running a part of a block needs dropout_p == 0");
    return B200B_ERR_ARG;
  }
  float* x1w = do_rest ? s.x1 : x_out;                 // where the cross sub-layer writes
  const float* x1 = do_cross ? x1w : x_in;             // what the rest of the block reads

  // 1. cross-attention: x1 = x + W_o * SDPA(W_q * LN(x), K_i, V_i)          (bridge_module.py:316-323)
  if (do_cross) {
    B200B_TRY(b200b_layernorm_fwd_rows(x_in, w->ln_c_g, w->ln_c_b, s.xn1, s.mean1, s.rstd1, T, D, eps, st));
    B200B_TRY(gemm(s.xn1, 0, D, w->wq_c, 0, D, T, D, D, B200B_EPI_BF16_BIAS, s.q, D, w->bq_c, nullptr, nullptr, 0, 0.f,
                   0, 0, st));
    if (d.flags & (B200B_BRIDGE_KV_PACKED | B200B_BRIDGE_KV_TC)) {
      if (p != 0.f || d.L > 64) {
        set_last_error("block_forward: a packed K/V cache needs dropout_p == 0 and len_text <= 64");
        return B200B_ERR_ARG;
      }
      if (d.flags & B200B_BRIDGE_KV_TC)
        B200B_TRY(b200b_attention_decode_tc(s.q, D, kv, i, d.nb, s.o1, D, s.lse1, d.B, d.Hc, d.L, d.Nv, d.dc, st));
      else
        B200B_TRY(b200b_attention_decode_packed(s.q, D, kv, i, d.nb, s.o1, D, s.lse1, d.B, d.Hc, d.L, d.Nv, d.dc, st));
    } else {
      B200B_TRY(attn(false, s.q, D, kblk, ldkv, kblk + D, ldkv, s.o1, D, s.lse1, nullptr, nullptr, 0, nullptr, 0,
                     nullptr, 0, nullptr, 0, d.B, d.Hc, d.L, d.Nv, d.dc, p, seed, (ds_cross(i) | ind), st));
    }
    B200B_TRY(gemm(s.o1, 0, D, w->wo_c, 0, D, T, D, D, B200B_EPI_F32_BIAS_RESID, x1w, D, w->bo_c, x_in, nullptr, 0, 0.f,
                   0, 0, st));
  }
  if (!do_rest) return B200B_OK;
  // 2. self-attention (non-causal, unmasked)                                 (:326-328)
  B200B_TRY(b200b_layernorm_fwd_rows(x1, w->ln_s_g, w->ln_s_b, s.xn2, s.mean2, s.rstd2, T, D, eps, st));
  B200B_TRY(gemm(s.xn2, 0, D, w->wqkv_s, 0, D, T, 3 * D, D, B200B_EPI_BF16_BIAS, s.qkv, 3 * D, w->bqkv_s, nullptr,
                 nullptr, 0, 0.f, 0, 0, st));
  B200B_TRY(attn(false, s.qkv, 3 * D, s.qkv + D, 3 * D, s.qkv + 2 * D, 3 * D, s.o2, D, s.lse2, nullptr, nullptr, 0,
                 nullptr, 0, nullptr, 0, nullptr, 0, d.B, d.Hs, d.L, d.L, d.ds, p, seed, (ds_self(i) | ind), st));
  B200B_TRY(gemm(s.o2, 0, D, w->wo_s, 0, D, T, D, D, B200B_EPI_F32_BIAS_RESID, s.x2, D, w->bo_s, x1, nullptr, 0, 0.f,
                 0, 0, st));
  // 3. FFN: x3 = x2 + drop(W_2 * drop(gelu(W_1 * LN(x2))))                    (:331-333)
  B200B_TRY(b200b_layernorm_fwd_rows(s.x2, w->ln_f_g, w->ln_f_b, s.xn3, s.mean3, s.rstd3, T, D, eps, st));
  B200B_TRY(gemm(s.xn3, 0, D, w->w1, 0, D, T, F, D, B200B_EPI_BF16_BIAS_GELU, s.h, F, w->b1, nullptr, s.u, F, p, seed,
                 (ds_ffn_h(i) | ind), st));
  B200B_TRY(gemm(s.h, 0, F, w->w2, 0, F, T, D, F, B200B_EPI_F32_BIAS_RESID, x_out, D, w->b2, s.x2, nullptr, 0, p, seed,
                 (ds_ffn_o(i) | ind), st));
  return B200B_OK;
}

extern "C" int b200b_bridge_block_backward(const b200b_bridge_dims* dims, int i, const b200b_block_weights* w,
                                           const float* x_in, const void* kv, const void* saved, const float* d_out,
                                           float* d_in, void* dkv, const b200b_block_grads* g, void* workspace,
                                           size_t workspace_bytes, float p, uint64_t seed,
                                           const b200b_grad_notify* notify, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  Dims d;
  B200B_TRY(read_dims(dims, &d, "block_backward"));
  if (!w || !x_in || !kv || !saved || !d_out || !dkv || !g || !workspace || i < 0 || i >= d.nb) {
    set_last_error("block_backward: null argument or bad block index");
    return B200B_ERR_ARG;
  }
  Saved s = carve_saved(d, const_cast<void*>(saved));
  BwdWs ws = carve_bwd(d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_last_error("block_backward: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes);
    return B200B_ERR_WORKSPACE;
  }
  const int T = (int)d.T, D = d.D, F = d.F;
  // B200B_BRIDGE_SEED_INDIRECT: `seed` holds a device pointer; every kernel reads the seed itself
  const uint32_t ind = (d.flags & B200B_BRIDGE_SEED_INDIRECT) ? B200B_SEED_INDIRECT : 0u;
  const long long ldkv = 2LL * D * d.nb;
  const __nv_bfloat16* kblk = reinterpret_cast<const __nv_bfloat16*>(kv) + (size_t)2 * D * i;
  __nv_bfloat16* dkblk = reinterpret_cast<__nv_bfloat16*>(dkv) + (size_t)2 * D * i;
  const int EB = B200B_EPI_BF16_BIAS;
  // weight gradients: fp32, or bf16 (no bias operand) when the caller exchanges them in bf16
  const int EW = (d.flags & B200B_BRIDGE_WGRAD_BF16) ? B200B_EPI_BF16_BIAS : B200B_EPI_F32;
  auto ready = [&](const void* grad, long long elems) {
    if (notify != nullptr && notify->fn != nullptr) notify->fn(notify->user, grad, elems);
  };

  const int rchunks = b200b_row_chunks(T);
  if (rchunks <= 0 || rchunks > kMaxRowChunks) {
    set_last_error("block_backward: no usable device (row chunks = %d)", rchunks);
    return B200B_ERR_DEVICE;
  }
  Finalizer fin;
  int ch = 0;

  // ---- FFN ----
  // d(ffn.3 out) = dropout-bwd(bf16(d_out)); its column sums are the ffn.3 bias gradient
  B200B_TRY(b200b_cast_bf16_colsum(d_out, ws.dy, ws.p_cast, T, D, p, seed, (ds_ffn_o(i) | ind), st));
  fin.add(ws.p_cast, g->b2, D, rchunks, D);
  B200B_TRY(gemm(ws.dy, 1, D, s.h, 1, F, D, F, T, EW, g->w2, F, nullptr, nullptr, nullptr, 0, 0.f, 0, 0, st));
  ready(g->w2, (long long)D * F);
  B200B_TRY(gemm(ws.dy, 0, D, w->w2, 1, F, T, F, D, B200B_EPI_BF16_DGELU, ws.du, F, nullptr, nullptr, s.u, F, p, seed,
                 (ds_ffn_h(i) | ind), st));
  B200B_TRY(b200b_colsum_partials(ws.du, F, T, F, ws.p_du, &ch, st));
  fin.add(ws.p_du, g->b1, F, ch, F);
  B200B_TRY(gemm(ws.du, 1, F, s.xn3, 1, D, F, D, T, EW, g->w1, D, nullptr, nullptr, nullptr, 0, 0.f, 0, 0, st));
  ready(g->w1, (long long)F * D);
  B200B_TRY(gemm(ws.du, 0, F, w->w1, 1, D, T, D, F, EB, ws.dxn, D, nullptr, nullptr, nullptr, 0, 0.f, 0, 0, st));
  // ln_ffn backward (+ residual gradient d_out) -> dx; bf16(dx) = gradient of the self-attention W_o output
  B200B_TRY(b200b_layernorm_bwd_fused(ws.dxn, s.x2, s.mean3, s.rstd3, w->ln_f_g, d_out, ws.dx, ws.dy, ws.p_lnf, T, D, st));
  fin.add(ws.p_lnf, g->ln_f_b, D, rchunks, 3LL * D);
  fin.add(ws.p_lnf + D, g->ln_f_g, D, rchunks, 3LL * D);
  fin.add(ws.p_lnf + 2 * D, g->bo_s, D, rchunks, 3LL * D);
  // ---- self-attention ----
  B200B_TRY(grad_pair(ws.dy, D, w->wo_s, s.o2, D, T, D, D, ws.dattn, D, g->wo_s, EW, st));
  ready(g->wo_s, (long long)D * D);
  B200B_TRY(attn(true, s.qkv, 3 * D, s.qkv + D, 3 * D, s.qkv + 2 * D, 3 * D, s.o2, D, s.lse2, ws.dattn, ws.dqkv, 3 * D,
                 ws.dqkv + D, 3 * D, ws.dqkv + 2 * D, 3 * D, ws.attn_ws, ws.attn_ws_bytes, d.B, d.Hs, d.L, d.L, d.ds, p,
                 seed, (ds_self(i) | ind), st));
  B200B_TRY(b200b_colsum_partials(ws.dqkv, 3 * D, T, 3 * D, ws.p_dqkv, &ch, st));
  fin.add(ws.p_dqkv, g->bqkv_s, 3 * D, ch, 3LL * D);
  B200B_TRY(grad_pair(ws.dqkv, 3 * D, w->wqkv_s, s.xn2, D, T, 3 * D, D, ws.dxn, D, g->wqkv_s, EW, st));
  ready(g->wqkv_s, 3LL * D * D);
  B200B_TRY(b200b_layernorm_bwd_fused(ws.dxn, s.x1, s.mean2, s.rstd2, w->ln_s_g, ws.dx, ws.dx, ws.dy, ws.p_lns, T, D, st));
  fin.add(ws.p_lns, g->ln_s_b, D, rchunks, 3LL * D);
  fin.add(ws.p_lns + D, g->ln_s_g, D, rchunks, 3LL * D);
  fin.add(ws.p_lns + 2 * D, g->bo_c, D, rchunks, 3LL * D);
  // ---- cross-attention ----
  B200B_TRY(grad_pair(ws.dy, D, w->wo_c, s.o1, D, T, D, D, ws.dattn, D, g->wo_c, EW, st));
  ready(g->wo_c, (long long)D * D);
  __nv_bfloat16* dq = ws.dqkv;  // [T, D]
  B200B_TRY(attn(true, s.q, D, kblk, ldkv, kblk + D, ldkv, s.o1, D, s.lse1, ws.dattn, dq, D, dkblk, ldkv, dkblk + D, ldkv,
                 ws.attn_ws, ws.attn_ws_bytes, d.B, d.Hc, d.L, d.Nv, d.dc, p, seed, (ds_cross(i) | ind), st));
  B200B_TRY(b200b_colsum_partials(dq, D, T, D, ws.p_dq, &ch, st));
  fin.add(ws.p_dq, g->bq_c, D, ch, D);
  B200B_TRY(grad_pair(dq, D, w->wq_c, s.xn1, D, T, D, D, ws.dxn, D, g->wq_c, EW, st));
  ready(g->wq_c, (long long)D * D);
  // ln_cross backward: d_in (if wanted) = dx + LN input gradient; dgamma / dbeta always
  B200B_TRY(b200b_layernorm_bwd_fused(ws.dxn, x_in, s.mean1, s.rstd1, w->ln_c_g, ws.dx, d_in, nullptr, ws.p_lnc, T, D, st));
  fin.add(ws.p_lnc, g->ln_c_b, D, rchunks, 3LL * D);
  fin.add(ws.p_lnc + D, g->ln_c_g, D, rchunks, 3LL * D);
  return b200b_colsum_finalize(fin.t, fin.n, st);
}

extern "C" int b200b_bridge_kv_backward(const b200b_bridge_dims* dims, const void* vision_bf16, const void* dkv,
                                        void* dwkv_all, float* dbkv_all, void* workspace, size_t workspace_bytes,
                                        void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  Dims d;
  B200B_TRY(read_dims(dims, &d, "kv_backward"));
  if (!vision_bf16 || !dkv || !dwkv_all || !dbkv_all || !workspace) {
    set_last_error("kv_backward: null argument");
    return B200B_ERR_ARG;
  }
  BwdWs ws = carve_bwd(d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_last_error("kv_backward: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes);
    return B200B_ERR_WORKSPACE;
  }
  const int n = 2 * d.D * d.nb;
  int ch = 0;
  B200B_TRY(b200b_colsum_partials(dkv, n, (int)d.Tv, n, ws.p_dkv, &ch, st));
  b200b_colsum_task t;
  t.partials = ws.p_dkv; t.out = dbkv_all; t.cols = n; t.chunks = ch; t.chunk_stride = n;
  B200B_TRY(b200b_colsum_finalize(&t, 1, st));
  const int EW = (d.flags & B200B_BRIDGE_WGRAD_BF16) ? B200B_EPI_BF16_BIAS : B200B_EPI_F32;
  return gemm(dkv, 1, n, vision_bf16, 1, d.Dv, n, d.Dv, (int)d.Tv, EW, dwkv_all, d.Dv, nullptr, nullptr,
              nullptr, 0, 0.f, 0, 0, st);
}

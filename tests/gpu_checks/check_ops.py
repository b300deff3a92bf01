"""On-GPU bring-up check of the row kernels and the fused attention (run under gpurun)."""
from __future__ import annotations

import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _harness  # noqa: E402

# ("ln", rows, dim) | ("colsum", rows, cols) | ("cast", n, p) | ("attn", B, H, Lq, Lk, d, p)
CASES = [
    ("ln", 1024, 2304), ("ln", 37, 2304), ("ln", 5, 128), ("ln", 64, 1000),
    ("colsum", 1024, 2304), ("colsum", 2056, 9216), ("colsum", 7, 64),
    ("cast", 1 << 20, 0.0), ("cast", 2304 * 64, 0.25),
    ("attn", 1, 1, 16, 32, 64, 0.0),
    ("attn", 1, 1, 64, 32, 128, 0.0),
    ("attn", 2, 3, 40, 100, 64, 0.0),
    ("attn", 2, 18, 64, 64, 128, 0.0),     # self-attention shape (C1)
    ("attn", 2, 8, 64, 257, 288, 0.0),     # cross-attention shape (C1)
    ("attn", 8, 8, 128, 257, 288, 0.0),    # C2 cross
    ("attn", 8, 18, 128, 128, 128, 0.0),   # C2 self
    ("attn", 1, 8, 5, 257, 288, 0.0),      # decode-like short prefix
    ("attn", 32, 8, 1, 257, 288, 0.0),     # C4 decode steps (K/V-streaming kernel): 1, 2, 3, 4 query row groups
    ("attn", 32, 8, 17, 257, 288, 0.0),
    ("attn", 4, 8, 33, 257, 288, 0.0),
    ("attn", 4, 8, 48, 257, 288, 0.0),
    ("attn", 32, 8, 64, 257, 288, 0.0),
    ("attn", 2, 8, 20, 5, 288, 0.0),       # one key tile: the second key split sees no key
    ("attn", 2, 18, 3, 3, 128, 0.0),       # self-attention of a 3-token prefix
    ("attn", 2, 8, 64, 1370, 288, 0.0),    # long K/V: many ring wrap-arounds
    ("attn", 2, 8, 128, 1370, 288, 0.0),   # C5 vision length
    ("attn", 2, 8, 64, 257, 288, 0.1),     # dropout
    ("attn", 2, 18, 64, 64, 128, 0.1),
]
NAMES = ["_".join(str(x) for x in c) for c in CASES]


def relerr(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-30)


def run_case(i: int) -> dict:
    import torch

    from vlm_bridge_b200 import ops

    c = CASES[i]
    torch.manual_seed(7 + i)
    dev = "cuda"
    res: dict = {}
    if c[0] == "ln":
        _, rows, dim = c
        x = torch.randn(rows, dim, device=dev) * 2 + 0.5
        g = torch.randn(dim, device=dev)
        b = torch.randn(dim, device=dev)
        y, mean, rstd = ops.layernorm_fwd(x, g, b)
        xr = x.clone().requires_grad_()
        yr = torch.nn.functional.layer_norm(xr, (dim,), g, b, 1e-5)
        e_y = relerr(y, yr.detach())
        dy = torch.randn(rows, dim, device=dev).bfloat16()
        dres = torch.randn(rows, dim, device=dev)
        dx = ops.layernorm_bwd(dy, x, mean, rstd, g, dres)
        gr = g.clone().requires_grad_()
        br = b.clone().requires_grad_()
        yr2 = torch.nn.functional.layer_norm(xr, (dim,), gr, br, 1e-5)
        yr2.backward(dy.float())
        e_dx = relerr(dx, xr.grad + dres)
        dbeta, dgamma = ops.colsum(dy, x=x, mean=mean, rstd=rstd)
        e_dg = relerr(dgamma, gr.grad)
        e_db = relerr(dbeta, br.grad)
        # in-place form (dx aliases dres)
        d2 = dres.clone()
        ops.layernorm_bwd(dy, x, mean, rstd, g, d2, out=d2)
        e_alias = relerr(d2, dx)
        # CTA-per-row variants used by the whole-block entry points
        y2, mean2, rstd2 = ops.layernorm_fwd_rows(x, g, b)
        e_rows = max(relerr(y2, yr.detach()), relerr(mean2, mean), relerr(rstd2, rstd))
        fdx, fdy, fdb, fdg, fcs = ops.layernorm_bwd_fused(dy, x, mean, rstd, g, dres)
        e_f = max(relerr(fdx, xr.grad + dres), relerr(fdg, gr.grad), relerr(fdb, br.grad),
                  relerr(fdy, (xr.grad + dres).bfloat16()), relerr(fcs, fdy.float().sum(0)))
        _, _, fdb2, fdg2, _ = ops.layernorm_bwd_fused(dy, x, mean, rstd, g, None, want_dx=False)
        e_f = max(e_f, relerr(fdg2, gr.grad), relerr(fdb2, br.grad))
        res.update(e_y=e_y, e_dx=e_dx, e_dgamma=e_dg, e_dbeta=e_db, e_alias=e_alias, e_rows=e_rows, e_fused=e_f)
        res["ok"] = (e_y < 6e-3 and e_dx < 1e-4 and e_dg < 1e-4 and e_db < 1e-4 and e_alias == 0 and e_rows < 6e-3
                     and e_f < 8e-3)
    elif c[0] == "colsum":
        _, rows, cols = c
        big = torch.randn(rows, cols + 64, device=dev).bfloat16()
        dy = big[:, :cols]  # pitched view
        s = ops.colsum(dy)
        e = relerr(s, dy.float().sum(0))
        e2 = relerr(ops.colsum_two_stage(dy), dy.float().sum(0))
        res.update(e=e, e_two_stage=e2)
        res["ok"] = e < 1e-4 and e2 < 1e-4
    elif c[0] == "cast":
        _, n, p = c
        x = torch.randn(n, device=dev)
        y = ops.cast_bf16(x, dropout_p=p, seed=99, dropout_stream=4)
        if p == 0:
            res["ok"] = bool(torch.equal(y, x.bfloat16()))
        else:
            keep = y != 0
            ref = (x.bfloat16().float() / (1 - p)).bfloat16()
            res["keep_frac"] = keep.float().mean().item()
            res["ok"] = bool(torch.equal(y[keep], ref[keep])) and abs(res["keep_frac"] - (1 - p)) < 0.01
            # the mask must equal the one the GEMM residual epilogue draws for the same stream/index
            M, N = n // 2304, 2304
            a = torch.zeros(M, 64, device=dev, dtype=torch.bfloat16)
            w = torch.zeros(N, 64, device=dev, dtype=torch.bfloat16)
            bias = torch.ones(N, device=dev)
            resid = torch.zeros(M, N, device=dev)
            o = ops.gemm(a, w, epilogue=ops.EPI_F32_BIAS_RESID, bias=bias, resid=resid, dropout_p=p, seed=99,
                         dropout_stream=4)
            res["mask_matches_gemm"] = bool(torch.equal(o.reshape(-1) != 0, keep))
            res["ok"] = res["ok"] and res["mask_matches_gemm"]
            # fused cast + column sums (the block backward's first kernel): same bytes, same mask
            y2, cs = ops.cast_bf16_colsum(x.view(M, N), dropout_p=p, seed=99, dropout_stream=4)
            res["fused_equal"] = bool(torch.equal(y2.reshape(-1), y))
            res["e_colsum"] = relerr(cs, y2.float().sum(0))
            res["ok"] = res["ok"] and res["fused_equal"] and res["e_colsum"] < 1e-4
    else:
        _, B, H, Lq, Lk, d, p = c
        D = H * d
        # q/k/v live inside wider fused buffers to exercise row pitches
        qb = (torch.randn(B * Lq, D + 64, device=dev)).bfloat16()
        kvb = (torch.randn(B * Lk, 2 * D, device=dev)).bfloat16()
        q, k, v = qb[:, :D], kvb[:, :D], kvb[:, D:]
        o, lse = ops.attention_fwd(q, k, v, batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=d, dropout_p=p,
                                   seed=5, dropout_stream=1)
        qr = q.float().reshape(B, Lq, H, d).transpose(1, 2).requires_grad_()
        kr = k.float().reshape(B, Lk, H, d).transpose(1, 2).requires_grad_()
        vr = v.float().reshape(B, Lk, H, d).transpose(1, 2).requires_grad_()
        s = (qr @ kr.transpose(-1, -2)) / math.sqrt(d)
        pr = torch.softmax(s, -1)
        d_o = torch.randn(B * Lq, D, device=dev).bfloat16()
        dq = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16)
        dkv = torch.empty(B * Lk, 2 * D, device=dev, dtype=torch.bfloat16)
        if p == 0:
            oref = (pr @ vr).transpose(1, 2).reshape(B * Lq, D)
            e_o = relerr(o, oref.detach())
            e_lse = (lse * math.log(2) - torch.logsumexp(s, -1)).abs().max().item()
            ops.attention_bwd(d_o, q, k, v, o, lse, dq, dkv[:, :D], dkv[:, D:], batch=B, heads=H, len_q=Lq,
                              len_k=Lk, head_dim=d)
            oref.backward(d_o.float())
            e_dq = relerr(dq, qr.grad.transpose(1, 2).reshape(B * Lq, D))
            e_dk = relerr(dkv[:, :D], kr.grad.transpose(1, 2).reshape(B * Lk, D))
            e_dv = relerr(dkv[:, D:], vr.grad.transpose(1, 2).reshape(B * Lk, D))
            res.update(e_o=e_o, e_lse=e_lse, e_dq=e_dq, e_dk=e_dk, e_dv=e_dv)
            res["ok"] = e_o < 1e-2 and e_lse < 1e-3 and e_dq < 2e-2 and e_dk < 2e-2 and e_dv < 2e-2
        else:
            # recover the mask from a run with V = identity-like probes is costly; instead check
            # (a) determinism, (b) E[o] ~ undropped o, (c) fwd/bwd mask consistency through a
            # finite-difference-free identity: with dO = o_drop-independent random, compare against a
            # torch reference that uses the mask recovered from P_drop = o when V = I (d >= Lk only).
            o2, _ = ops.attention_fwd(q, k, v, batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=d, dropout_p=p,
                                      seed=5, dropout_stream=1)
            res["deterministic"] = bool(torch.equal(o, o2))
            o0, _ = ops.attention_fwd(q, k, v, batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=d)
            res["mean_ratio"] = (o.float().abs().mean() / o0.float().abs().mean()).item()
            ok = res["deterministic"] and 0.8 < res["mean_ratio"] < 1.6
            if Lk <= d:
                # V = [I; 0] per head -> o[:, :Lk] = P_drop exactly (bf16), so the mask is observable
                eye = torch.zeros(B * Lk, D, device=dev, dtype=torch.bfloat16)
                for h in range(H):
                    for j in range(Lk):
                        eye[j::Lk][:, h * d + j] = 0  # placeholder to keep shapes obvious
                ev = eye.reshape(B, Lk, H, d)
                idx = torch.arange(Lk, device=dev)
                ev[:, idx, :, idx] = 1
                ev2 = ev.reshape(B * Lk, D)
                opd, lse_d = ops.attention_fwd(q, k, ev2, batch=B, heads=H, len_q=Lq, len_k=Lk, head_dim=d,
                                               dropout_p=p, seed=5, dropout_stream=1)
                pd = opd.float().reshape(B, Lq, H, d).transpose(1, 2)[..., :Lk]   # [B,H,Lq,Lk]
                mask = (pd != 0).float()
                res["keep_frac"] = mask.mean().item()
                e_p = relerr(pd, (pr.detach() * mask / (1 - p)))
                # backward against torch autograd with that mask
                evr = ev2.float().reshape(B, Lk, H, d).transpose(1, 2).requires_grad_()
                oref = ((pr * mask / (1 - p)) @ evr).transpose(1, 2).reshape(B * Lq, D)
                ops.attention_bwd(d_o, q, k, ev2, opd, lse_d, dq, dkv[:, :D], dkv[:, D:], batch=B, heads=H,
                                  len_q=Lq, len_k=Lk, head_dim=d, dropout_p=p, seed=5, dropout_stream=1)
                oref.backward(d_o.float())
                e_dq = relerr(dq, qr.grad.transpose(1, 2).reshape(B * Lq, D))
                e_dk = relerr(dkv[:, :D], kr.grad.transpose(1, 2).reshape(B * Lk, D))
                e_dv = relerr(dkv[:, D:], evr.grad.transpose(1, 2).reshape(B * Lk, D))
                res.update(e_p=e_p, e_dq=e_dq, e_dk=e_dk, e_dv=e_dv)
                ok = ok and e_p < 1e-2 and e_dq < 2e-2 and e_dk < 2e-2 and e_dv < 2e-2 and \
                    abs(res["keep_frac"] - (1 - p)) < 0.02
            res["ok"] = bool(ok)
    torch.cuda.synchronize()
    return res


if __name__ == "__main__":
    sys.exit(_harness.main(os.path.abspath(__file__), NAMES, run_case, "check_ops.jsonl"))

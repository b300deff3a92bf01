"""On-GPU bring-up check of the tcgen05 GEMM (run under gpurun; not a pytest file).

Each case runs in its own subprocess with a timeout so that a wedged kernel cannot stall the
whole call. Results are appended to gpurun_out/check_gemm.jsonl.
Usage: python tests/gpu_checks/check_gemm.py            (driver: all cases)
       python tests/gpu_checks/check_gemm.py --case i   (one case, in-process)
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# (name, M, N, K, a_major, b_major, epilogue, block_n)
CASES = [
    ("kk_small_128", 128, 128, 64, 0, 0, 0, 128),
    ("kk_small_256", 128, 256, 64, 0, 0, 0, 256),
    ("kk_k256_128", 128, 128, 256, 0, 0, 0, 128),
    ("kk_multi_tile_128", 384, 384, 512, 0, 0, 0, 128),
    ("kk_multi_tile_256", 384, 512, 512, 0, 0, 0, 256),
    ("kk_persistent_128", 2048, 2304, 2304, 0, 0, 0, 128),     # > 148 tiles: 2nd accumulator stage + wraparound
    ("kk_persistent_256", 2048, 9216, 2304, 0, 0, 0, 256),
    ("kk_ragged_m", 2056, 4608, 1024, 0, 0, 0, 256),           # M tail (2056 = 16*128 + 8)
    ("kk_ragged_small", 40, 72, 136, 0, 0, 0, 128),            # M, N, K tails
    ("kk_gelu", 1024, 9216, 2304, 0, 0, 1, 256),
    ("kk_gelu_drop", 256, 512, 256, 0, 0, 1, 256),
    ("kk_resid", 1024, 2304, 9216, 0, 0, 2, 128),
    ("kk_resid_drop", 256, 256, 256, 0, 0, 2, 128),
    ("kk_dgelu", 256, 512, 256, 0, 0, 3, 256),
    ("kk_f32", 2304, 1024, 1024, 0, 0, 4, 256),
    ("kk_f32_beta", 256, 256, 128, 0, 0, 4, 128),
    ("kmn_small", 128, 128, 64, 0, 1, 0, 128),                 # dgrad form: B stored [K, N]
    ("kmn_small_256", 128, 256, 128, 0, 1, 0, 256),
    ("kmn_dgrad", 1024, 2304, 9216, 0, 1, 0, 128),
    ("kmn_dgelu", 1024, 9216, 2304, 0, 1, 3, 256),
    ("mnmn_small", 128, 128, 64, 1, 1, 4, 128),                # wgrad form: A stored [K, M], B stored [K, N]
    ("mnmn_small_256", 128, 256, 128, 1, 1, 4, 256),
    ("mnmn_wgrad", 9216, 2304, 1024, 1, 1, 4, 256),
    ("mnmn_wgrad_ragged_k", 4608, 1024, 2056, 1, 1, 4, 256),   # K tail (2056 vision rows)
    ("kk_auto", 1024, 6912, 2304, 0, 0, 0, 0),
]
CASES = [c + (1,) for c in CASES]
# CTA-pair (cta_group::2) variants: 256 x block_n tiles
CASES += [
    ("pair_kk_small_128", 256, 128, 64, 0, 0, 0, 128, 2),
    ("pair_kk_small_256", 256, 256, 128, 0, 0, 0, 256, 2),
    ("pair_kk_multi_128", 768, 384, 512, 0, 0, 0, 128, 2),
    ("pair_kk_persistent_256", 2048, 9216, 2304, 0, 0, 0, 256, 2),
    ("pair_kk_persistent_128", 2048, 2304, 2304, 0, 0, 0, 128, 2),
    ("pair_kk_ragged", 2056, 4608, 1024, 0, 0, 0, 256, 2),
    ("pair_kk_ragged_small", 40, 72, 136, 0, 0, 0, 128, 2),
    ("pair_kk_gelu_drop", 512, 512, 256, 0, 0, 1, 256, 2),
    ("pair_kk_resid", 1024, 2304, 9216, 0, 0, 2, 128, 2),
    ("pair_kmn_dgrad", 1024, 2304, 9216, 0, 1, 0, 128, 2),
    ("pair_kmn_dgelu", 1024, 9216, 2304, 0, 1, 3, 256, 2),
    ("pair_mnmn_small_128", 256, 128, 64, 1, 1, 4, 128, 2),
    ("pair_mnmn_wgrad", 9216, 2304, 1024, 1, 1, 4, 256, 2),
    ("pair_mnmn_wgrad_128", 2304, 2304, 1024, 1, 1, 4, 128, 2),
    ("pair_mnmn_ragged_k", 4608, 1024, 2056, 1, 1, 4, 256, 2),
    ("auto_ffn_up", 1024, 9216, 2304, 0, 0, 0, 0, 0),
    ("auto_proj", 1024, 2304, 2304, 0, 0, 0, 0, 0),
]


def run_case(i: int) -> dict:
    import torch

    from vlm_bridge_b200 import ops

    name, M, N, K, am, bm, epi, bn, cg = CASES[i]
    torch.manual_seed(100 + i)
    dev = "cuda"
    a_log = (torch.randn(M, K, device=dev) * 0.5).bfloat16()   # logical A [M,K]
    b_log = (torch.randn(N, K, device=dev) * 0.5).bfloat16()   # logical B [N,K]
    a = a_log.t().contiguous() if am else a_log
    b = b_log.t().contiguous() if bm else b_log
    acc = a_log.float() @ b_log.float().t()
    bias = torch.randn(N, device=dev)
    p = 0.25 if "drop" in name else 0.0
    seed = 1234
    res = {"case": name, "M": M, "N": N, "K": K}
    if epi == 0:
        out = ops.gemm(a, b, a_major=am, b_major=bm, epilogue=0, bias=bias, block_n=bn, cta_group=cg)
        ref = acc + bias
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        tol = 1e-2
    elif epi == 1:
        out, u = ops.gemm(a, b, a_major=am, b_major=bm, epilogue=1, bias=bias, block_n=bn, cta_group=cg,
                          dropout_p=p, seed=seed, dropout_stream=3)
        u_ref = (acc + bias).bfloat16().float()
        h_ref = torch.nn.functional.gelu(u_ref).bfloat16().float()
        err_u = (u.float() - u_ref).abs().max().item() / u_ref.abs().max().item()
        if p > 0:
            keep = out != 0
            res["keep_frac"] = keep.float().mean().item()
            err = ((out.float() - h_ref / (1 - p)) * keep).abs().max().item() / h_ref.abs().max().item()
            # mask must be reproducible
            out2, _ = ops.gemm(a, b, a_major=am, b_major=bm, epilogue=1, bias=bias, block_n=bn, cta_group=cg,
                               dropout_p=p, seed=seed, dropout_stream=3)
            res["mask_repro"] = bool(torch.equal(out, out2))
        else:
            err = (out.float() - h_ref).abs().max().item() / h_ref.abs().max().item()
        err = max(err, err_u)
        tol = 1.5e-2
    elif epi == 2:
        resid = torch.randn(M, N, device=dev)
        out = ops.gemm(a, b, a_major=am, b_major=bm, epilogue=2, bias=bias, resid=resid, block_n=bn, cta_group=cg,
                       dropout_p=p, seed=seed, dropout_stream=5)
        y = (acc + bias).bfloat16().float()
        if p > 0:
            d = out - resid
            keep = d.abs() > 1e-6
            res["keep_frac"] = keep.float().mean().item()
            err = ((d - y / (1 - p)) * keep).abs().max().item() / y.abs().max().item()
        else:
            err = (out - (resid + y)).abs().max().item() / y.abs().max().item()
        tol = 1.5e-2
    elif epi == 3:
        u = torch.randn(M, N, device=dev).bfloat16()
        out = ops.gemm(a, b, a_major=am, b_major=bm, epilogue=3, aux=u, block_n=bn, cta_group=cg)
        uf = u.float().requires_grad_()
        torch.nn.functional.gelu(uf).backward(acc.bfloat16().float())
        ref = uf.grad
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        tol = 1.5e-2
    else:
        beta = 0.5 if "beta" in name else 0.0
        out0 = torch.randn(M, N, device=dev)
        out = out0.clone()
        ops.gemm(a, b, a_major=am, b_major=bm, epilogue=4, out=out, beta=beta, block_n=bn, cta_group=cg)
        ref = beta * out0 + acc
        err = (out - ref).abs().max().item() / ref.abs().max().item()
        tol = 1e-4
    torch.cuda.synchronize()
    res["rel_err"] = err
    res["ok"] = bool(err < tol) and res.get("mask_repro", True)
    if res["ok"] and M * N * K >= 1024 * 2304 * 1024 and epi in (0, 4):
        # quick throughput sample (CUDA events, 20 launches after 3 warm-ups)
        kw = dict(a_major=am, b_major=bm, epilogue=epi, block_n=bn, cta_group=cg)
        if epi == 0:
            kw["bias"] = bias
            o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        else:
            o = torch.empty(M, N, device=dev, dtype=torch.float32)
        for _ in range(3):
            ops.gemm(a, b, out=o, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm(a, b, out=o, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        # cuBLAS on the same logical problem for context
        for _ in range(3):
            torch.matmul(a_log, b_log.t())
        e0.record()
        for _ in range(20):
            torch.matmul(a_log, b_log.t())
        e1.record()
        torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / 20) / 1e9
    return res


def main() -> int:
    if "--from" in sys.argv:
        # in-process: run cases i.. until one raises (a CUDA fault poisons the context -> exit)
        i0 = int(sys.argv[sys.argv.index("--from") + 1])
        for i in range(i0, len(CASES)):
            print(f"START {i}", flush=True)
            try:
                res = run_case(i)
            except Exception as e:  # noqa: BLE001
                print("RESULT " + json.dumps({"case": CASES[i][0], "idx": i, "ok": False,
                                              "error": repr(e)[:600]}), flush=True)
                return 1
            res["idx"] = i
            print("RESULT " + json.dumps(res), flush=True)
        return 0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out_path = os.path.join(ROOT, "gpurun_out", "check_gemm.jsonl")
    results: dict[int, dict] = {}
    nxt = 0
    while nxt < len(CASES):
        started = nxt
        try:
            r = subprocess.run([sys.executable, __file__, "--from", str(nxt)], capture_output=True,
                               text=True, timeout=420)
            stdout, stderr, note = r.stdout, r.stderr, f"rc={r.returncode}"
        except subprocess.TimeoutExpired as e:
            stdout = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            stderr = (e.stderr or b"").decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
            note = "timeout"
        last_start = None
        for line in stdout.splitlines():
            if line.startswith("START "):
                last_start = int(line[6:])
            elif line.startswith("RESULT "):
                res = json.loads(line[7:])
                results[res["idx"]] = res
        if last_start is not None and last_start not in results:
            results[last_start] = {"case": CASES[last_start][0], "idx": last_start, "ok": False,
                                   "error": note, "stderr": stderr[-1500:], "stdout": stdout[-600:]}
        done = max(results) if results else started
        if last_start is None and started not in results:
            results[started] = {"case": CASES[started][0], "idx": started, "ok": False,
                                "error": "no output " + note, "stderr": stderr[-1500:]}
            done = started
        nxt = done + 1
    n_ok = 0
    with open(out_path, "w") as f:
        for i in sorted(results):
            n_ok += bool(results[i].get("ok"))
            f.write(json.dumps(results[i]) + "\n")
            print(json.dumps(results[i]), flush=True)
    print(f"check_gemm: {n_ok}/{len(CASES)} ok")
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Runs csrc/probe_tcgen05.cu: tcgen05.mma with 32 / 64 / 128-byte swizzled K-major operands and M = 64 / 128,
compared with torch.matmul; prints which TMEM lanes hold which accumulator rows."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from vlm_bridge_b200 import _lib

lib = _lib.lib()
torch.manual_seed(0)
for m, n, k, w in ((64, 32, 288, 32), (64, 32, 288, 64), (64, 32, 256, 128), (128, 32, 288, 32), (128, 32, 288, 64),
                   (64, 64, 288, 32), (64, 64, 288, 64), (64, 144, 32, 32), (64, 144, 32, 64), (64, 144, 64, 128),
                   (64, 128, 288, 64), (128, 128, 288, 64), (128, 256, 256, 128), (64, 16, 288, 32)):
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    dump = torch.full((128, n), float("nan"), device="cuda")
    cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
    reps = 20
    rc = lib.b200b_probe_umma(a.data_ptr(), b.data_ptr(), dump.data_ptr(), m, n, k, w, reps, cyc.data_ptr(),
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    res = {"m": m, "n": n, "k": k, "swizzle": w, "rc": rc}
    if rc == 0:
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()                       # [m, n]
        # for every accumulator row, the TMEM lane whose dump matches it best
        err = (dump[:, None, :] - ref[None, :, :]).abs().amax(-1)     # [128 lanes, m rows]
        err = torch.nan_to_num(err, nan=1e30)
        best = err.argmin(0)                                   # lane per row
        worst = float(err.min(0).values.max())
        lanes = best.tolist()
        res["max_abs_err_at_best_lane"] = worst
        res["ok"] = worst < 1e-2 * float(ref.abs().max())
        res["lane_of_row_0_15_16_31_32_47_48_63"] = [lanes[i] for i in (0, 15, 16, 31, 32, 47, 48, min(63, m - 1))]
        res["identity_map"] = lanes == list(range(m))
        n_mma = reps * (k // 16)
        res["cycles_per_mma_issue"] = round(float(cyc[0]) / n_mma, 1)
        res["cycles_per_mma_complete"] = round(float(cyc[1]) / n_mma, 1)
    else:
        res["error"] = lib.b200b_last_error().decode()
    print(json.dumps(res), flush=True)

"""Runs probe_umma_layouts.cu (see its header) against torch.matmul for every operand form the
tcgen05 attention kernels use: K-major / MN-major shared-memory operands with 64- and 128-byte
swizzle, A from tensor memory, M = 64 and 128. One JSON line per case in gpurun_out/probe_umma_layouts.jsonl."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _harness  # noqa: E402

# (m, n, k, a_mode, b_mode, wa, wb, mn_variant)
CASES = [
    (128, 64, 288, 0, 0, 64, 64, 0),     # control: K-major both (validated in round 1)
    (128, 128, 64, 0, 1, 128, 128, 0),   # control: B MN-major 128B swizzle (the GEMM's dgrad form)
    (128, 128, 64, 1, 1, 128, 128, 0),   # control: A and B MN-major 128B (the GEMM's wgrad form)
    (128, 160, 64, 0, 1, 64, 64, 0),     # B MN-major 64B swizzle: P V / dS K
    (128, 160, 64, 0, 1, 64, 64, 1),
    (128, 128, 48, 0, 1, 64, 64, 0),
    (128, 32, 16, 0, 1, 64, 64, 0),
    (128, 160, 128, 1, 1, 64, 64, 0),    # A and B MN-major 64B: dV = P^T dO, dK = dS^T Q (M = 128 keys)
    (128, 160, 128, 1, 1, 64, 64, 1),
    (64, 160, 128, 1, 1, 64, 64, 0),     # the same with M = 64 keys
    (64, 128, 128, 1, 1, 64, 64, 0),
    (128, 64, 64, 2, 0, 64, 64, 0),      # A from tensor memory, B K-major
    (128, 160, 64, 2, 1, 64, 64, 0),     # A from tensor memory, B MN-major
    (128, 96, 288, 0, 0, 64, 64, 0),     # S tile of 96 keys
    (128, 48, 288, 0, 0, 64, 64, 0),
    (128, 16, 288, 0, 0, 64, 64, 0),     # last key tile (1 valid key padded to 16)
]
NAMES = ["m%d n%d k%d a%d b%d wa%d wb%d v%d" % c for c in CASES]


def run_case(i: int) -> dict:
    import torch

    lib = C.CDLL(os.path.join(HERE, "libprobe_umma.so"))
    m, n, k, am, bm, wa, wb, var = CASES[i]
    torch.manual_seed(i)
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    dump = torch.full((128, n), float("nan"), device="cuda")
    rc = lib.probe_umma_layouts(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(dump.data_ptr()), m, n, k,
                                am, bm, wa, wb, var, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    res = {"m": m, "n": n, "k": k, "a_mode": am, "b_mode": bm, "wa": wa, "wb": wb, "mn_variant": var, "rc": rc}
    if rc != 0:
        res["ok"] = False
        return res
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = (dump[:, None, :] - ref[None, :, :]).abs().amax(-1)     # [128 lanes, m rows]
    err = torch.nan_to_num(err, nan=1e30)
    best = err.argmin(0).tolist()
    worst = float(err.min(0).values.max())
    res["max_abs_err_at_best_lane"] = worst
    res["ok"] = worst < 1e-2 * float(ref.abs().max())
    res["identity_lane_map"] = best == list(range(m))
    if m == 64:
        res["lane_of_row_0_15_16_31_32_47_48_63"] = [best[j] for j in (0, 15, 16, 31, 32, 47, 48, 63)]
    return res


if __name__ == "__main__":
    sys.exit(_harness.main(os.path.abspath(__file__), NAMES, run_case, "probe_umma_layouts.jsonl", worker_timeout=240))

"""bench.py prints ONE JSON line on stdout whatever libraries write to fd 1 (NCCL prints its version
banner there on some boxes); host-side helpers of the bench (no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_carries_only_the_json_line():
    code = (
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "bench.own_stdout()\n"
        "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')\n"      # a C library writing to fd 1
        "print('python-level noise')\n"
        "bench.emit({'metric': 'x', 'value': 1.5})\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout) == {"metric": "x", "value": 1.5} and r.stdout.count("\n") == 1
    assert "NCCL version" in r.stderr and "python-level noise" in r.stderr


def test_algorithmic_work_figures():
    sys.path.insert(0, ROOT)
    import bench

    # C2 GEMM flops of a fwd+bwd step (text linears x3, K/V projections x2)
    assert abs(bench.gemm_flops_per_step(8, 128, 257) / 1e9 - 990.83) < 0.01
    # decode bytes: 2 blocks x (K,V read + Q read + O written); with position rows block 0 has one query position
    D, B, Nv = 2304, 32, 257
    kv = B * 2 * Nv * D * 2
    assert bench.decode_bytes_per_step(B, 64, Nv) == 2 * (kv + 2 * B * 64 * D * 2)
    assert bench.decode_bytes_per_step(B, 64, Nv, position_rows=True) == 2 * kv + 2 * B * D * 2 * (1 + 64)
    total = sum(bench.decode_bytes_per_step(B, s, Nv) for s in range(1, 65))
    assert abs(total / 1e9 - 10.93) < 0.01

"""Shared driver for the on-GPU bring-up checks (run under gpurun; not pytest files).

Cases run in-process one after another; a CUDA fault poisons the context, so on the first exception
the worker exits and the driver restarts a fresh worker at the next case. A hung kernel is bounded
by the worker timeout (and by the mbarrier watchdog inside the kernels).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(script: str, names: list[str], run_case, out_name: str, worker_timeout: int = 420) -> int:
    if "--from" in sys.argv:
        i0 = int(sys.argv[sys.argv.index("--from") + 1])
        for i in range(i0, len(names)):
            print(f"START {i}", flush=True)
            try:
                res = run_case(i)
            except Exception as e:  # noqa: BLE001
                print("RESULT " + json.dumps({"case": names[i], "idx": i, "ok": False,
                                              "error": repr(e)[:500],
                                              "tb": traceback.format_exc()[-800:]}), flush=True)
                return 1
            res["idx"] = i
            res.setdefault("case", names[i])
            print("RESULT " + json.dumps(res), flush=True)
        return 0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results: dict[int, dict] = {}
    nxt = 0
    while nxt < len(names):
        started = nxt
        try:
            r = subprocess.run([sys.executable, script, "--from", str(nxt)], capture_output=True, text=True,
                               timeout=worker_timeout)
            stdout, stderr, note = r.stdout, r.stderr, f"rc={r.returncode}"
        except subprocess.TimeoutExpired as e:
            stdout = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            stderr = e.stderr.decode() if isinstance(e.stderr, bytes) else (e.stderr or "")
            note = "timeout"
        last_start = None
        for line in stdout.splitlines():
            if line.startswith("START "):
                last_start = int(line[6:])
            elif line.startswith("RESULT "):
                res = json.loads(line[7:])
                results[res["idx"]] = res
        if last_start is not None and last_start not in results:
            results[last_start] = {"case": names[last_start], "idx": last_start, "ok": False, "error": note,
                                   "stderr": stderr[-1500:], "stdout": stdout[-600:]}
        if last_start is None and started not in results:
            results[started] = {"case": names[started], "idx": started, "ok": False,
                                "error": "no output " + note, "stderr": stderr[-1500:]}
        nxt = max(max(results), started) + 1
    n_ok = 0
    with open(os.path.join(ROOT, "gpurun_out", out_name), "w") as f:
        for i in sorted(results):
            n_ok += bool(results[i].get("ok"))
            f.write(json.dumps(results[i]) + "\n")
            print(json.dumps(results[i]), flush=True)
    print(f"{out_name}: {n_ok}/{len(names)} ok")
    return 0

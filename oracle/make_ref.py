"""Recipe that makes the UNMODIFIED reference importable on the GPU box -- TEST / BASELINE INFRASTRUCTURE ONLY.

/root/reference exists only in the build container. This script copies the reference's own Python package
(src/vlm_bridge, .py files only, byte for byte) from where it lies under /root/reference into
oracle/_ref/vlm_bridge. oracle/_ref/ is git-ignored (no reference source ever enters the history) but not
gpurun-ignored, so it travels to the GPU box like the built .so files. Users of the copy:
  * bench.py --impl reference and bench.py's cpu_baseline leg: the reference `BridgeLite` itself (train mode,
    dropout 0.1, fp32 on the host cores) instead of the oracle port  -> `cpu_baseline.kind == "reference"`;
  * tests/test_inloop_gpu.py: the reference `FullModel` / training step with this repository's module swapped in.
The product package never imports it (tests/test_no_oracle_in_product.py).

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/vlm_bridge"
DST = os.path.join(HERE, "_ref", "vlm_bridge")


def make(verbose: bool = False) -> str | None:
    """Returns the destination package directory, or None when /root/reference is absent (GPU box)."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, SRC)
        out = os.path.join(DST, rel) if rel != "." else DST
        os.makedirs(out, exist_ok=True)
        for f in files:
            if not f.endswith(".py"):
                continue
            a, b = os.path.join(root, f), os.path.join(out, f)
            if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                shutil.copyfile(a, b)
                n += 1
    if verbose:
        print(f"oracle/_ref: {n} file(s) refreshed from {SRC}")
    return DST


def bridge_module_path() -> str | None:
    p = os.path.join(DST, "model_architecture", "bridge_module.py")
    return p if os.path.exists(p) else None


def load_reference_bridge_module():
    """The reference's bridge_module.py imported by file path (it imports only torch), or None."""
    import importlib.util

    p = bridge_module_path()
    if p is None:
        return None
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_bridge_module", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(make(verbose=True))

"""The oracle is test infrastructure: nothing in the product package may import, call or execute
anything under oracle/ (or /root/reference), and bench.py may only touch it in its CPU legs."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vlm-bridge-for-image-captioning_b200")


def _sources(top):
    for d, _, fs in os.walk(top):
        if "build" in d.split(os.sep):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(d, f)


def test_package_never_touches_oracle_or_reference():
    bad = []
    for path in list(_sources(PKG)) + [os.path.join(ROOT, "vlm_bridge_b200", "__init__.py")]:
        src = open(path).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "bridge_oracle" in src or "/root/reference" in src:
            bad.append(path)
    assert not bad, bad


def test_package_has_no_cpu_or_eager_fallback():
    src = open(os.path.join(PKG, "bridge.py")).read()
    assert "F.scaled_dot_product_attention" not in src and "torch.nn.functional" not in src
    assert "F.linear" not in src and "F.layer_norm" not in src


def test_bench_uses_oracle_only_in_cpu_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    # two imports, each inside its own helper: _cpu_oracle() (the port) and _cpu_reference_module() (the unmodified
    # reference copied to oracle/_ref by oracle/make_ref.py); only the CPU legs (functions named cpu_*) call them
    uses = [m.start() for m in re.finditer(r"from oracle import", src)]
    assert len(uses) == 2
    helpers = ("_cpu_oracle", "_cpu_reference_module")
    for u in uses:
        enclosing = re.findall(r"^def (\w+)", src[:u], flags=re.M)[-1]
        assert enclosing in helpers, enclosing
    for h in helpers:
        for m in re.finditer(h + r"\(\)", src):
            if src[m.start() - 4:m.start()] == "def ":
                continue
            enclosing = re.findall(r"^def (\w+)", src[:m.start()], flags=re.M)[-1]
            assert enclosing.startswith("cpu_"), enclosing
    assert "/root/reference" not in src

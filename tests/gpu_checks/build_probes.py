"""Builds the bring-up probes under tests/gpu_checks/ (test infrastructure; none of it ships in
libb200_bridge.so): nvcc -gencode arch=compute_100a,code=sm_100a -> tests/gpu_checks/libprobe_umma.so."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libprobe_umma.so")


def build() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(HERE, "probe_umma_layouts.cu")]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-shared", "-o", OUT, *srcs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build())

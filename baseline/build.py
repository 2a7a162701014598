"""Builds baseline/lib/libupstream_structure.so (the "upstream-structure" GPU baseline, see upstream_structure.cu).
It links against the product library for the per-Gaussian kernels it shares with it."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "lib" / "libupstream_structure.so"


def build(force: bool = False) -> Path:
    from ggrt_official_b200 import build as pb

    product = pb.build()
    src = HERE / "upstream_structure.cu"
    if not force and LIB.exists() and LIB.stat().st_mtime >= max(src.stat().st_mtime, product.stat().st_mtime):
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    cmd = [pb.find_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", str(LIB), str(src), f"-L{product.parent}", "-lggrt_raster",
           "-Xlinker", "-rpath=$ORIGIN/../../ggrt_official_b200/lib"]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))

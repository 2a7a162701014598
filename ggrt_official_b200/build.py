"""Builds libggrt_raster.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libggrt_raster.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build libggrt_raster.so)")


def sources():
    return sorted(CSRC.glob("*.cu"))


STAMP = LIB_DIR / "libggrt_raster.sha256"


def source_hash() -> str:
    """Content hash of every file the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "ggrt_raster.h"]
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    return h.hexdigest()


def is_stale() -> bool:
    if not LIB.exists() or not STAMP.exists():
        return True
    return STAMP.read_text().strip() != source_hash()


def _replace_atomically(path: Path, data: str) -> None:
    tmp = path.with_name(f".{path.name}.{os.getpid()}.tmp")
    tmp.write_text(data)
    os.replace(tmp, path)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if the sources changed.  Safe under torchrun: every rank calls this lazily, so the
    staleness check + compile run under an exclusive file lock, nvcc writes to a private temporary file that is
    renamed into place (a concurrent CDLL never sees a half-written .so), and the stamp is replaced the same way
    after the library."""
    if not force and not is_stale():
        return LIB
    import fcntl

    LIB_DIR.mkdir(exist_ok=True)
    with open(LIB_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while this one waited for the lock
                return LIB
            tmp = LIB_DIR / f".libggrt_raster.{os.getpid()}.so.tmp"
            cmd = [find_nvcc(), *NVCC_FLAGS, "-o", str(tmp), *map(str, sources())]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            res = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if verbose:
                print(res.stdout, res.stderr)
            if res.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
            os.replace(tmp, LIB)
            _replace_atomically(STAMP, source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def build_variant(name: str, defines: dict, verbose: bool = False) -> Path:
    """Tuning experiments: the same sources with other -D knobs -> gpurun_variants/<name>/libggrt_raster.so
    (git-ignored, shipped to the GPU box; load it with GGRT_RASTER_LIB=<path>)."""
    out_dir = PKG.parent / "gpurun_variants" / name
    out_dir.mkdir(parents=True, exist_ok=True)
    out = out_dir / "libggrt_raster.so"
    cmd = [find_nvcc(), *NVCC_FLAGS, *[f"-D{k}={v}" for k, v in defines.items()], "-o", str(out), *map(str, sources())]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose:
        print(res.stdout, res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Ahead-of-time build of libyabpe.so (nvcc, sm_100a only) next to the Python host layer.

The shared object is built IN-TREE (yet-another-bpe_b200/yabpe/libyabpe.so) so that it travels
with the repository snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "csrc" / "yabpe.cu"
OUT = HERE / "yabpe" / "libyabpe.so"
DEPS = sorted((HERE / "csrc").glob("*.cu*")) + [HERE / "csrc" / "unicode_tables.inc", HERE.parent / "include" / "yabpe.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


HOST_SRC = HERE / "csrc" / "hostlist.c"
HOST_OUT = HERE / "yabpe" / "_hostlist.so"


def build_hostlist(force: bool = False) -> Path | None:
    """gcc -> yabpe/_hostlist.so: the Python objects of a training result built with the C API (host glue, no compute).
    Optional: without Python.h the package uses its Python construction of the same objects."""
    import sysconfig
    if not force and HOST_OUT.exists() and HOST_OUT.stat().st_mtime >= HOST_SRC.stat().st_mtime:
        return HOST_OUT
    inc = sysconfig.get_paths()["include"]
    if not (Path(inc) / "Python.h").exists():
        return None
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-I", inc, "-o", str(HOST_OUT), str(HOST_SRC)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        return None
    return HOST_OUT


def build(force: bool = False, verbose: bool = False) -> Path:
    build_hostlist(force)
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return OUT
    import os
    import shlex
    extra = shlex.split(os.environ.get("YABPE_NVCC_EXTRA", ""))      # tuning experiments only
    cmd = ["nvcc", *NVCC_FLAGS, *extra] + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(OUT), str(SRC)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libyabpe.so")
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

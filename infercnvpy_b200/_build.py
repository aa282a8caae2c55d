"""In-tree nvcc build of ``libicnv.so`` for sm_100a.

The shared library lands next to this file (``infercnvpy_b200/libicnv.so``) so it
travels to the GPU box with the repo snapshot.  ``build()`` is a no-op when the
library is newer than every source.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libicnv.so"
OBJ_DIR = PKG / "csrc" / "_obj"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas",
    "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found; libicnv.so cannot be built")
    return exe


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple[str, ...] = (), out: Path | None = None) -> Path:
    """``defines`` / ``out`` build an experimental variant next to the product library (A/B timing
    on one box through ICNV_LIB_PATH, see tools/ab.sh); the default call builds ``libicnv.so``."""
    global LIB
    if out is None and not force and not _stale():
        return LIB
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    logs = {}
    tag = "" if out is None else "_" + Path(out).stem
    target = LIB if out is None else Path(out)

    def compile_one(src: Path):
        obj = OBJ_DIR / (src.stem + tag + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src.name] = r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(4, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    tmp = target.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    tmp.replace(target)
    (OBJ_DIR / f"ptxas{tag}.log").write_text("\n".join(f"==== {k}\n{v}" for k, v in logs.items()))
    if verbose:
        print(f"built {target} ({target.stat().st_size / 1e6:.1f} MB)")
    return target


if __name__ == "__main__":
    build(force=True, verbose=True)

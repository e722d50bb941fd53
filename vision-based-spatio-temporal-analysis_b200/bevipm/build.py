"""In-tree nvcc build of libbevipm.so for sm_100a (no JIT cache: the .so travels with the tree).

Every csrc/*.cu is compiled to its own object (in parallel, re-compiled only when it or a header it may include
changed) and the objects are linked into one shared library.
"""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG.parent / "csrc"
INCLUDE = PKG.parents[1] / "include"
LIB = PKG / "libbevipm.so"
OBJ = PKG.parent / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # no -use_fast_math and no default contraction games: every rounding is written out with
    # _rn intrinsics in the sources; -fmad=false keeps nvcc from fusing anything we did not write
    "-fmad=false",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _headers_mtime() -> float:
    deps = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h"))
    return max((p.stat().st_mtime for p in deps), default=0.0)


def _obj(src: Path) -> Path:
    return OBJ / (src.stem + ".o")


def _src_mtime(src: Path) -> float:
    """A source's own time stamp, or that of a .cu it includes (bevipm_run_plan.cu compiles bevipm_run.cu a second time)."""
    t = src.stat().st_mtime
    for line in src.read_text().splitlines():
        if line.startswith('#include "') and line.rstrip().endswith('.cu"'):
            inc = CSRC / line.split('"')[1]
            if inc.exists():
                t = max(t, inc.stat().st_mtime)
    return t


def _stale_objects():
    hm = _headers_mtime()
    out = []
    for s in sources():
        o = _obj(s)
        if not o.exists() or o.stat().st_mtime < max(_src_mtime(s), hm):
            out.append(s)
    return out


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return _headers_mtime() > t or any(s.stat().st_mtime > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    OBJ.mkdir(exist_ok=True)
    todo = sources() if force else _stale_objects()

    def compile_one(src: Path):
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", "-o", str(_obj(src)), str(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, todo))
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *[str(_obj(s)) for s in sources()]]
    subprocess.run(link, check=True)
    return LIB


if __name__ == "__main__":
    import sys
    build(force="-f" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)

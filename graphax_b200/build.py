"""In-tree build of libgraphax_b200.so for sm_100a (nvcc cross-compiles without a GPU).

The library holds the C ABI, the generic interpreter kernel and one ahead-of-time,
plan-specialised kernel per named (workload, order) pair of ``configs.py``.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
from pathlib import Path
from typing import List

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
GEN = CSRC / "gen"
OBJ = CSRC / "build"
LIB = Path(os.environ.get("GX_LIB") or PKG / "libgraphax_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-cudart", "static", f"-I{CSRC}"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _compile(src: Path, obj: Path, extra: List[str]) -> str:
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def generate_aot_sources() -> List[Path]:
    from .codegen import emit_source, plan_key
    from .configs import named_plans
    GEN.mkdir(parents=True, exist_ok=True)
    for old in GEN.glob("aot_*.cu"):
        old.unlink()
    paths, seen = [], set()
    for name, plan in named_plans():
        key = plan_key(plan)
        if key in seen:
            continue
        seen.add(key)
        path = GEN / f"aot_{key:016x}.cu"
        path.write_text(f"// {name}\n" + emit_source(plan, nvrtc=False))
        paths.append(path)
    return paths


def build_library(verbose: bool = False, with_aot: bool = True) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    for old in OBJ.glob("*.o"):
        old.unlink()
    sources = [CSRC / "gx_abi.cu", CSRC / "gx_gemm.cu"] + (generate_aot_sources() if with_aot else [])
    objs = [OBJ / (s.stem + ".o") for s in sources]
    extra = ["-Xptxas", "-v"] if verbose else []
    workers = max(1, min(len(sources), len(os.sched_getaffinity(0))))
    with cf.ThreadPoolExecutor(workers) as ex:
        logs = list(ex.map(lambda so: _compile(so[0], so[1], extra), zip(sources, objs)))
    if verbose:
        for s, log in zip(sources, logs):
            print(f"== {s.name}\n{log}")
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static", "-Xcompiler", "-fPIC", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(verbose="-v" in sys.argv))

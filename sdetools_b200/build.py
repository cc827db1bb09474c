"""Builds sdetools_b200/libsdegpu.so in-tree with nvcc for sm_100a (no JIT cache:
the built .so travels to the GPU box with the repo snapshot).

    python -m sdetools_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libsdegpu.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", INCLUDE]
SCALAR_MODELS = (0, 1, 2, 3, 4, 5, 7, 8)        # sdegpu_model ids compiled from ensemble_model.cu (6 = linear has its own files)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsdegpu.so cannot be built")


def _units():
    units = [("api.o", "api.cu", []), ("integrals.o", "integrals.cu", []), ("linear.o", "linear.cu", []),
             ("linear_dmma.o", "linear_dmma.cu", []), ("laws.o", "laws.cu", []), ("linlaws.o", "linlaws.cu", []), ("linear_tc.o", "linear_tc.cu", [])]
    units += [(f"ensemble_model_{m}.o", "ensemble_model.cu", [f"-DSDEGPU_TU_MODEL={m}"]) for m in SCALAR_MODELS]
    return [u for u in units if os.path.exists(os.path.join(CSRC, u[1]))]


def _sources_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "sdegpu.h"), __file__]
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _sources_mtime():
        return SO
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)      # the image exports a CC that nvcc should not pick up

    def compile_one(u):
        obj, src, defs = u
        extra = os.environ.get("SDEGPU_NVCC_EXTRA", "").split()     # tuning experiments, e.g. -DSDEGPU_TERMINAL_THREADS=512
        cmd = [nvcc, *NVCC_FLAGS, *defs, *extra, "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {defs}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    units = _units()
    with ThreadPoolExecutor(max_workers=min(8, len(units))) as ex:
        logs = list(ex.map(compile_one, units))
    if verbose:
        sys.stderr.write("\n".join(logs))
    objs = [os.path.join(OBJ, u[0]) for u in units]
    cmd = [nvcc, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-Xlinker", "--exclude-libs,ALL", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

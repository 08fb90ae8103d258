"""Build libbsi_b200.so (sm_100a) in-tree with nvcc.  `python -m bsi_b200.build [--force]`."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(PKG, "libbsi_b200.so")
SOURCES = ["elementwise.cu", "reduce.cu", "layernorm.cu", "attention.cu", "attention_sm100.cu", "gemm_sm100.cu", "dit_engine.cu", "unet_kernels.cu", "unet_engine.cu", "optim.cu", "wgrad_sm100.cu", "train_kernels.cu", "attention_bwd.cu", "attention_bwd_sm100.cu", "exact_kernels.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libbsi_b200.so cannot be built on this machine")
    return cand


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(os.path.dirname(PKG), "include", "bsi_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def is_stale() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


def build(force: bool = False, verbose: bool = False, variant: str | None = None, defines: tuple[str, ...] = ()) -> str:
    """Build the product library, or (variant given) an experimental copy libbsi_b200_<variant>.so compiled with extra -D flags."""
    lib_path = LIB if variant is None else os.path.join(PKG, f"libbsi_b200_{variant}.so")
    obj_dir = OBJ if variant is None else os.path.join(OBJ, variant)
    if variant is None and not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    if "--variant" in sys.argv:
        name = sys.argv[sys.argv.index("--variant") + 1]
        print(build(force=True, verbose=True, variant=name, defines=tuple(a for a in sys.argv if a.startswith("-D"))))
    else:
        print(build(force="--force" in sys.argv, verbose=True))

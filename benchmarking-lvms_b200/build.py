"""Build libblvm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).  Importable without the
library being present (it is what creates it): `python benchmarking-lvms_b200/build.py` or `__graft_entry__.build()`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libblvm_b200.so")
SOURCES = ["blvm_b200.cu", "blvm_dmol_f32.cu", "blvm_dmol_f16.cu", "blvm_dmol_bf16.cu", "blvm_linear.cu"]   # one object per source, compiled in parallel
HEADERS = ["blvm_math.cuh", "ptx_sm100.cuh", "dmol_kernels.cuh", "dmol_stream_kernel.cuh", "dmol_dispatch.cuh", "kl_kernels.cuh", "misc_kernels.cuh", "sample_kernels.cuh", "linear_dmol_kernel.cuh", "host_common.h",
           os.path.join("..", "..", "include", "blvm_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return res.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    objs = [os.path.join(obj_dir, os.path.splitext(s)[0] + ".o") for s in SOURCES]
    cmds = [[_nvcc()] + NVCC_FLAGS + extra + ["-c", "-o", o, os.path.join(CSRC, s)] for s, o in zip(SOURCES, objs)]
    with ThreadPoolExecutor(max_workers=len(cmds)) as pool:
        logs = list(pool.map(_run, cmds))
    _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    if verbose:
        sys.stderr.write("".join(logs))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

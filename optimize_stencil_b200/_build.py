"""Build recipe for the CUDA library (nvcc, sm_100a only) -- used by __graft_entry__.build().

The shared object is built IN-TREE (optimize_stencil_b200/csrc/libstencil_b200.so) so that it
travels to the GPU box with the repository snapshot.  It is git-ignored.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libstencil_b200.so")
SOURCES = ["stencil_b200_capi.cu"]
DEPENDS = SOURCES + ["dispersion_kernels.cuh", os.path.join("..", "..", "include", "stencil_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libstencil_b200.so")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPENDS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libstencil_b200.so for sm_100a.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed with exit code %d" % res.returncode)
    with open(os.path.join(CSRC, "ptxas_info.txt"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

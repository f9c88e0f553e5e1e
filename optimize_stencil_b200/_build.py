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
# The same library with sqrtarg in the reference's operation order at EVERY k-point (-DSB200_FAST_A=0): the
# differential check of the regrouped kernel in tests/test_gpu_parity.py loads it through SB200_LIB.
LIB_REFORDER = os.path.join(CSRC, "libstencil_b200_reforder.so")
SOURCES = ["stencil_b200_capi.cu"]
DEPENDS = SOURCES + ["dispersion_kernels.cuh", "objective_fa.cuh", os.path.join("..", "..", "include", "stencil_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libstencil_b200.so")


def is_stale(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPENDS)


def _compile(lib, defines, verbose, log):
    cmd = [find_nvcc()] + NVCC_FLAGS + list(defines) + ["-o", lib] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed with exit code %d" % res.returncode)
    if log:
        with open(os.path.join(CSRC, log), "w") as f:
            f.write(res.stderr)
    return lib


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libstencil_b200.so for sm_100a.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    return _compile(LIB, [], verbose, "ptxas_info.txt")


def build_reforder(force=False, verbose=False):
    """The reference-order variant (test infrastructure: the product never loads it)."""
    if not force and not is_stale(LIB_REFORDER):
        return LIB_REFORDER
    return _compile(LIB_REFORDER, ["-DSB200_FAST_A=0"], verbose, None)


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_reforder(force=True))

"""Builds libmgym.so (sm_100a only) in-tree with nvcc; see csrc/Makefile."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmgym.so")
SOURCES = ["mgym_api.cu", "mgym_kernels.cuh", "mgym_device.cuh", "Makefile"]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "..", "include", "mgym.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile the CUDA extension for sm_100a.  nvcc cross-compiles without a GPU."""
    if force or is_stale():
        r = subprocess.run(["make", "-C", CSRC] + (["-B"] if force else []), capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout)
            print(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("building libmgym.so failed; see output above")
    return LIB_PATH

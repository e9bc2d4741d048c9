"""In-tree build of librsb200.so (sm_100a only) with nvcc.

The library is a plain C-ABI shared object (include/rsb200.h); there is no
torch / pybind dependency, so this is a single nvcc invocation.  nvcc
cross-compiles without a GPU, which is what ``__graft_entry__.build()`` relies
on.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librsb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--shared",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "rsb200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into recstudio_b200/librsb200.so."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build librsb200.so")
    tmp = LIB + ".tmp"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + sources()
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + "\n".join(l for l in log.splitlines() if "error" in l.lower())[:4000])
    os.replace(tmp, LIB)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

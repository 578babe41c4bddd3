"""Build libemoport.so (all sm_100a kernels + the C-ABI) in-tree with nvcc.

    python -m emoportraits_b200.csrc.build [--force]

The library is built next to the sources so that it travels with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import pathlib
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
LIB = HERE / "libemoport.so"
SOURCES = ["conv_igemm.cu", "grid_sample.cu", "norm.cu", "misc.cu"]
HEADERS = ["common.cuh", "pose_math.cuh", "conv_igemm_kernel.inc", "apply_kernel.inc", "../../include/emoportraits_b200.h"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in [os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"]:
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + ["build.py"]:
        h.update((HERE / f).read_bytes())
    h.update(os.environ.get("EMO_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> pathlib.Path:
    """debug=True: the instrumented build (-DEMO_CONV_DEBUG: parts of the conv kernel can be switched off at run time,
    tools/conv_bound_probe.py) as libemoport_dbg.so next to the product library; loaded only through EMO_LIB=<path>."""
    LIB = HERE / ("libemoport_dbg.so" if debug else "libemoport.so")
    stamp = HERE / (".build_stamp_dbg" if debug else ".build_stamp")
    dig = _digest()
    if LIB.exists() and not force and stamp.exists() and stamp.read_text() == dig:
        return LIB
    cmd = [_nvcc(), *ARCH_FLAGS, "-O3", "-std=c++17", "-lineinfo"]  # no --use_fast_math: expf/tanhf/division stay IEEE-accurate
    cmd += ["-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]
    cmd += os.environ.get("EMO_NVCC_EXTRA", "").split()
    if debug:
        cmd += ["-DEMO_CONV_DEBUG"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    tmp = LIB.with_name(LIB.name + ".tmp")  # link into a temporary name, then rename: a reader (or a gpurun snapshot)
    cmd += ["-o", str(tmp)] + [str(HERE / s) for s in SOURCES]  # never sees a half-written library
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        tmp.unlink(missing_ok=True)
        raise RuntimeError("nvcc failed building libemoport.so")
    os.replace(tmp, LIB)
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv)
    print(p)

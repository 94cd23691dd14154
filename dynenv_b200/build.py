"""Builds the C-ABI shared library (dynenv_b200/libdynenv_b200.so) with nvcc for sm_100a.

In-tree build: the .so is git-ignored but travels to the GPU box with the repo snapshot.
nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libdynenv_b200.so")
SOURCES = ["csrc/dynenv_api.cu", "csrc/rc_slow.cu"]
HEADERS = ["csrc/dynenv_common.cuh", "csrc/rc_sim.cuh", "csrc/drv_sim.cuh", "csrc/rc_raster.cuh", "csrc/rc_obs.cuh", "csrc/rc_phys2.cuh", "csrc/rc_contact.cuh", "csrc/rc_slow.h", "csrc/arrange.cuh", "csrc/rollout.cuh", "csrc/raster_compose.cuh", "csrc/raster_prims.cuh", "../include/dynenv_b200.h", "../include/dynenv_rng.h"]
# --use_fast_math: float32 division / square root / exp as the 2-ulp hardware approximations (MUFU.RCP, MUFU.RSQ,
# MUFU.EX2 + a multiply) with denormals flushed, instead of the IEEE sequences and their slow-path calls.  The step
# kernels are bound by instruction fetch: the flag shrinks them by ~20 % (rc_phys_kernel 105 -> 82 KB of SASS) and buys
# 11-12 % env-steps/s (measured: IEEE 15.2 ms / step, -prec-div/-prec-sqrt=false 14.1, + -ftz 14.0, fast math 13.6);
# every GPU parity test passes unchanged (tolerance 1e-4).  The two-float sums use the explicit
# round-to-nearest intrinsics and float64 code (car angles, image projection) is not affected.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-shared", "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(_DIR, p)) > t for p in SOURCES + HEADERS
               if os.path.exists(os.path.join(_DIR, p)))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + \
          [os.path.join(_DIR, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

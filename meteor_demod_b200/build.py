"""In-tree build of liblrpt_b200.so (nvcc for sm_100a + gcc for the C host part).

Cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(os.path.dirname(HERE), "include")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(HERE, "liblrpt_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("CC", "gcc")
CU_SOURCES = ["lrpt_api.cu", "demod_simple.cu", "demod_ws.cu", "demod_spec.cu", "demod_lane.cu", "shard_stitch.cu", "shard_run.cu", "fir_stage.cu", "frontend.cu", "acquire.cu"]
C_SOURCES = ["lrpt_params.c"]
HEADERS = ["demod_core.cuh", "ws_common.cuh", "kernels.h", "lrpt_internal.h", os.path.join(INC, "lrpt_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", INC, "-I", CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    for src in C_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [CC, "-O2", "-ffp-contract=off", "-std=gnu99", "-fPIC", "-Wall", "-Wextra", "-I", INC, "-I", CSRC,
                   "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(SO, objs):
        cmd = [NVCC, "-shared", "-o", SO] + objs + ["-cudart", "static", "-lm", "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    build_host(verbose, force)
    return SO


HOST_SRC = os.path.join(os.path.dirname(HERE), "host", "lrpt_demod.c")
HOST_BIN = os.path.join(os.path.dirname(HERE), "host", "lrpt_demod")


def build_host(verbose=False, force=False):
    """The C host (reference-compatible command line) linked against liblrpt_b200.so."""
    if force or _stale(HOST_BIN, [HOST_SRC, SO, os.path.join(INC, "lrpt_b200.h")]):
        cmd = [CC, "-O2", "-std=gnu99", "-Wall", "-Wextra", "-I", INC, "-o", HOST_BIN, HOST_SRC,
               "-L", HERE, "-llrpt_b200", "-Wl,-rpath,$ORIGIN/../meteor_demod_b200", "-lm", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return HOST_BIN


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))

"""Build libkws.so in-tree with nvcc for sm_100a (B200).  `python -m speech_recognition_b200.build`."""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
PROFILE = os.environ.get("KWS_PROFILE_BUILD", "0") not in ("", "0")
# the profiling build (knockout switches + event traces compiled in) is a SECOND library, libkws_prof.so, selected at
# run time with KWS_LIBKWS=<path>: both travel to the GPU box, so an A/B visit does not spend GPU minutes compiling
OBJ = os.path.join(HERE, "csrc", "_obj_prof" if PROFILE else "_obj")
LIB = os.path.join(HERE, "libkws_prof.so" if PROFILE else "libkws.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]
if PROFILE:                                                          # knockout switches + event trace in the kernels
    NVCC_FLAGS.append("-DKWS_PROFILE_BUILD=1")
if os.environ.get("KWS_FIR_FP16", "0") not in ("", "0"):            # A/B aid: the r01 packed-half depthwise FIR
    NVCC_FLAGS.append("-DKWS_FIR_FP16=1")


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(os.path.dirname(HERE), "include", "kws.h")]
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {s}")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

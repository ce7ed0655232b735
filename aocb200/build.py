"""Builds aocb200/libaocb200.so (in-tree) with nvcc for sm_100a.  `python -m aocb200.build [--force]`."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# tooling builds (e.g. AOCB200_BUILD_TAG=trace AOCB200_NVCC_FLAGS=-DAOC_CONV_TRACE for tools/conv_trace.py) go to their own
# object directory and library name; aocb200.lib loads them when AOCB200_LIB_TAG names the same tag
TAG = os.environ.get("AOCB200_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libaocb200%s.so" % ("_" + TAG if TAG else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("AOCB200_NVCC_FLAGS", "").split()


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "aocb200.h")]
    hdr_m = max(os.path.getmtime(h) for h in hdrs)
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(s, o) or hdr_m > os.path.getmtime(o):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        with open(o[:-2] + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(cc, jobs))
    if jobs or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build librefign_b200.so (sm_100a only) in-tree with nvcc.

    python -m refign_b200.build [--force] [--verbose]

The shared object lands next to this file (refign_b200/librefign_b200.so); it is
git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc
cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "librefign_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-ccbin", "/usr/bin/g++"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "refign_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose, ptxas_v):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [NVCC] + ARCH + FLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if (verbose or ptxas_v) and r.stderr:
        print(r.stderr)
    return obj


def build_library(force=False, verbose=False, ptxas_v=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdr_t = _deps()
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose, ptxas_v), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    lib = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                        ptxas_v="--ptxas" in sys.argv)
    print(lib)

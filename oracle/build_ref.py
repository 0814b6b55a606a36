"""Compile the REFERENCE's own CPU correlation extension into oracle/_ref/.

Test infrastructure only.  The sources are compiled from where they lie under
/root/reference (models/correlation_ops/correlation.cpp and
correlation_sampler_cpu.cpp -- never copied into this repository); only the
resulting shared object is written, to oracle/_ref/correlation_ref_cpu.so,
which is git-ignored but travels to the GPU box with the snapshot.

The reference's own loader (models/correlation_ops/__init__.py:14-30) JIT-builds
into its read-only source directory and links a bare -lgomp, which fails with
this image's /opt/gcc wrapper; so we call g++ directly with torch's include and
library paths (no reference build system involved).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REFIGN_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "correlation_ref_cpu"


def build(verbose=False):
    src_dir = os.path.join(REF, "models", "correlation_ops")
    srcs = [os.path.join(src_dir, f) for f in ("correlation.cpp", "correlation_sampler_cpu.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return None
    import torch
    from torch.utils import cpp_extension
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, NAME + ".so")
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    incs = cpp_extension.include_paths() + [sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp",
           "-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    for i in incs:
        cmd += ["-isystem", i]
    cmd += srcs + ["-o", so, "-L" + libdir, "-Wl,-rpath," + libdir,
                   "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return so


def load():
    """Import the prebuilt reference extension (returns module or None)."""
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    so = build(verbose=True)
    print("built:" if so else "reference sources not found under", so or REF)
    sys.exit(0)

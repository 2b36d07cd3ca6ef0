"""In-tree build of libpylamp_b200.so (sm_100a only).

`python -m pylamp_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU;
the resulting shared object is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libpylamp_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--use_fast_math=false"]
# markers.cu needs IEEE-exact mul/div ordering for bit-identical cell indices: no FMA contraction
PER_FILE = {"markers.cu": ["-fmad=false"]}


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link pylamp_b200/lib/libpylamp_b200.so."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "..", "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "pylamp_b200.h"))
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = ["nvcc"] + ARCH + [c for c in COMMON if not c.startswith("--use_fast_math")] + \
                PER_FILE.get(src, []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out.decode()))
        if verbose and out:
            print(out.decode())
    target = lib_path()
    if force or procs or _stale(target, objs):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", target] + objs + ["-lcudart", "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s" % r.stdout.decode())
    return target


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))

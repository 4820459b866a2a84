"""Builds libmplb.so in-tree with nvcc for sm_100a (no torch dependency in the library).

Every source is compiled to its own object (in parallel, only when it or a header is newer than the object) and the
objects are linked into the shared library."""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("MPLB_OBJ") or os.path.join(HERE, "_obj")
LIB = os.environ.get("MPLB_LIB") or os.path.join(HERE, "libmplb.so")
SOURCES = ["capi.cu", "se3_kernels.cu", "nn_capi.cu", "nn_kernels.cu", "nn_grid.cu", "fetch_capi.cu", "fetch_kernels.cu",
           "planner_capi.cu", "planner_kernels.cu", "bvh_build.cpp", "collada.cpp", "bvh_cache.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-ffp-contract=off,-fopenmp"]
LINK_FLAGS = ["--shared", "-cudart", "shared", "-lgomp"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmplb.so cannot be built")


def _headers():
    inc = os.path.join(HERE, "..", "include")
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h", ".inc"))]
    return hs + [os.path.join(inc, "mplb.h")]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mplb.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    extra = os.environ.get("MPLB_NVCC_EXTRA", "").split()
    # nvcc's host compiler must be the distro g++ (the image also ships a wrapper without libgomp)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    nvcc = _nvcc()
    newest_header = max(os.path.getmtime(h) for h in _headers())
    tag = os.path.join(OBJ, ".flags")
    flags_now = " ".join(NVCC_FLAGS + extra)
    flags_changed = not os.path.exists(tag) or open(tag).read() != flags_now

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src + ".o")
        if (not force and not flags_changed and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(path), newest_header)):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + extra + ccbin + ["-c", "-o", obj, path]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _sources()))
    open(tag, "w").write(flags_now)
    cmd = [nvcc] + NVCC_FLAGS[:2] + LINK_FLAGS + ccbin + ["-o", LIB] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in os.sys.argv, verbose=True))

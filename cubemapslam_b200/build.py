"""Builds cubemapslam_b200/libcubemap_b200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the tree).
Every .cu is compiled to an object on its own (in parallel, only when stale) and the objects are linked into the shared library."""
import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcubemap_b200.so")
OBJ_DIR = os.path.join(_HERE, "build")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(_HERE, "csrc", "*.cu")))


def _headers():
    return glob.glob(os.path.join(_HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(_HERE, "csrc", "*.inc")) + \
        glob.glob(os.path.join(os.path.dirname(_HERE), "include", "*.h"))


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB_PATH, sources() + _headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr = _headers()
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + hdr)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj(src)]
        subprocess.check_call(cmd, cwd=os.path.join(_HERE, "csrc"))

    with ThreadPoolExecutor(max_workers=max(1, min(8, len(todo)))) as ex:
        list(ex.map(compile_one, todo))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + [_obj(s) for s in sources()] + ["-ldl"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

"""Builds cubemapslam_b200/libcubemap_b200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the tree)."""
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcubemap_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(glob.glob(os.path.join(_HERE, "csrc", "*.cu")))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(_HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(_HERE, "csrc", "*.inc")) + \
        glob.glob(os.path.join(os.path.dirname(_HERE), "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources() + ["-ldl"]
    subprocess.check_call(cmd, cwd=os.path.join(_HERE, "csrc"))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))

"""Build libmadrigal_b200.so (sm_100a only) in-tree with nvcc.

The shared library is git-ignored but travels to the GPU box with the working-tree snapshot, so it must be built
here (nvcc cross-compiles without a GPU).  `python -m madrigal_b200.build` or `__graft_entry__.build()`.
"""
import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmadrigal_b200.so")
STAMP_PATH = LIB_PATH + ".srchash"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _sources():
    files = [os.path.join(REPO_ROOT, "include", "madrigal_b200.h")]
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".inl", ".h")):
            files.append(os.path.join(CSRC, name))
    return files


def source_hash() -> str:
    h = hashlib.sha256()
    for f in _sources():
        h.update(os.path.relpath(f, REPO_ROOT).encode())  # relative: the tree is copied to another root on the GPU box
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def library_is_current() -> bool:
    return (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)
            and open(STAMP_PATH).read().strip() == source_hash())


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and library_is_current():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "madrigal_b200.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    with open(STAMP_PATH, "w") as f:
        f.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)

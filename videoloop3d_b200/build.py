"""Builds libvl3d.so (sm_100a) in-tree with nvcc.  `python -m videoloop3d_b200.build [--force] [-v]`.

Every translation unit is compiled to an object file (in parallel, only when it or a header changed) and the
objects are linked into videoloop3d_b200/lib/libvl3d.so.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libvl3d.so")
SOURCES = ["composite.cu", "fused_bwd_adam.cu", "patchnn.cu", "optim.cu", "exchange.cu", "alloc.cu", "terms.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(ROOT, "include", "vl3d.h"), os.path.abspath(__file__)]


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    hdr = _headers()
    return _stale(LIB_PATH, [os.path.join(CSRC, s) for s in SOURCES] + hdr)


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> videoloop3d_b200/lib/libvl3d.so"""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = nvcc_path()
    hdr = _headers()
    common = [nvcc] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-Xcompiler", "-fPIC"]
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(src):
        path, obj = os.path.join(CSRC, src), _obj(src)
        if not force and not _stale(obj, [path] + hdr):
            return src, ""
        r = subprocess.run(common + ["-c", path, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        return src, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(compile_one, SOURCES))
    if verbose:
        for src, log in logs:
            if log:
                print(f"==== {src}\n{log}")
    r = subprocess.run([nvcc] + ARCH + ["-shared", "-o", LIB_PATH] + [_obj(s) for s in SOURCES], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

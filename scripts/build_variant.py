"""Build an experimental variant of libvl3d.so for A/B runs (development tool).

    python scripts/build_variant.py <name> <file.cu> [nvcc flags, e.g. -DVL3D_S8_ROWFREE=1]

-> videoloop3d_b200/lib/libvl3d_<name>.so: the named translation unit recompiled with the flags, the other objects of the
regular build reused.  Select it at run time with VL3D_LIB=<path> (videoloop3d_b200/_lib.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from videoloop3d_b200 import build as B  # noqa: E402

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
obj = os.path.join(B.OBJ_DIR, f"variant_{name}_{os.path.splitext(src)[0]}.o")
common = [B.nvcc_path()] + B.ARCH
subprocess.check_call(common + ["-O3", "-lineinfo", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-Xcompiler", "-fPIC"] + flags +
                      ["-c", os.path.join(B.CSRC, src), "-o", obj])
out = os.path.join(B.LIB_DIR, f"libvl3d_{name}.so")
subprocess.check_call(common + ["-shared", "-o", out] + [B._obj(s) for s in B.SOURCES if s != src] + [obj])
print(out)

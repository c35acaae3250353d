"""Build a variant of the library for A/B measurements: the listed sources are recompiled with extra nvcc flags, the
other objects come from the regular build.  The result is cpflow_b200/lib/libcpflow_b200_NAME.so (use it with
CPF_LIB_PATH=...).   python tools/build_variant.py NAME src1.cu[,src2.cu] -DFLAG=1 ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpflow_b200 import build as B  # noqa: E402

name, srcs, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
B.build()
objs = []
for src in B.SOURCES:
    obj = os.path.join(B.OBJ, os.path.splitext(src)[0] + ".o")
    if src in srcs:
        obj = os.path.join(B.OBJ, os.path.splitext(src)[0] + f"_{name}.o")
        cmd = [B._nvcc()] + B.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-x", "cu", "-c", os.path.join(B.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        for ln in r.stderr.split("\n"):
            if "spill" in ln and " 0 bytes spill stores" not in ln or "Used" in ln and "heis_kernel" in prev:
                print(prev[:150]); print("   ", ln.strip())
            prev = ln
    objs.append(obj)
out = os.path.join(B.LIBDIR, f"libcpflow_b200_{name}.so")
subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, check=True)
print(out)

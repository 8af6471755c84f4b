"""Builds a diagnostic variant of the library: csrc/sparse_conv_umma.cu recompiled with extra -D flags, linked with the
shipped objects into cvpr2023-unidistill_b200/variants/lib_<name>.so (use through UD_B200_LIB).  Development aid.
    python tools/build_variant.py probe -DUD_DG_PROBE
    python tools/build_variant.py nomma -DUD_DG_PROBE -DUD_DG_EXP=2"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cvpr2023-unidistill_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
src_name = "sparse_conv_umma.cu"
for d in list(defs):
    if d.startswith("--src="):
        src_name = d[6:]
        defs.remove(d)
B.build()
out_dir = os.path.join(PKG, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "%s_%s.o" % (src_name[:-3], name))
cmd = [B._nvcc()] + B.NVCC_FLAGS + defs + ["-c", os.path.join(B.CSRC, src_name), "-o", obj]
subprocess.run(cmd, check=True, capture_output=True)
objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f != src_name[:-3] + ".o"]
lib = os.path.join(out_dir, "lib_%s.so" % name)
subprocess.run([B._nvcc(), "-shared", "-o", lib, obj] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart",
                                                                "-Xcompiler", "-fPIC"], check=True)
print(lib)

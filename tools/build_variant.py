"""Build a tuning variant of the library: one source recompiled with extra -D flags, linked
with the objects of the regular build.  usage: build_variant.py NAME SOURCE.cu -DX=1 ...
-> tools/experiments/libsbd_NAME.so (use with SBD_LIB_PATH=...)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sbdart_b200 import _build
name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
_build.build()
obj = f"/tmp/variant_{name}.o"
base = os.path.basename(src)[:-3]
subprocess.run(_build._nvcc() + _build.NVCC_FLAGS + [f'-DSBD_BUILD_ID="{_build.source_id()}"'] + flags +
               ["-c", os.path.join(_build.CSRC, os.path.basename(src)), "-o", obj], check=True)
objs = [os.path.join(_build.OBJ, o) for o in sorted(os.listdir(_build.OBJ)) if o.endswith(".o") and o != base + ".o"]
out = os.path.join(ROOT, "tools", "experiments", f"libsbd_{name}.so")
subprocess.run(_build._nvcc() + ["-shared", "-o", out] + objs + [obj], check=True, stderr=subprocess.DEVNULL)
print(out)

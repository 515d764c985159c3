"""In-tree nvcc build of the C-ABI library (sm_100a only).

Every .cu is compiled to an object file (in parallel), then linked into
libsbdart_b200.so.  The library carries a BUILD ID = SHA-256 over its sources, headers
and compiler flags (`sbd_build_id()`); `needs_build()` and `sbdart_b200.lib()` compare it
with the sources on disk, so a stale binary is rebuilt / refused whatever the file times say.
"""
from __future__ import annotations

import concurrent.futures
import glob
import hashlib
import json
import os
import subprocess
import time

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsbdart_b200.so")
OBJ = os.path.join(PKG, "build")
RECORD = os.path.join(PKG, "build_record.json")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(PKG, "..", "include", "*.h")))


def source_id() -> str:
    """SHA-256 over flags, sources and headers (names and contents)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sources() + headers():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:32]


def built_id():
    """Build id recorded next to the library by the last build (None: no build)."""
    try:
        with open(RECORD) as fh:
            rec = json.load(fh)
        return rec.get("build_id") if os.path.exists(LIB) else None
    except (OSError, ValueError):
        return None


def needs_build() -> bool:
    return built_id() != source_id()


def _nvcc():
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")]
    # nvcc's host compiler: the plain system gcc (the /opt/gcc wrapper lacks specs)
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    return cmd


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link.  force=True recompiles every source even if the build id matches."""
    sid = source_id()
    if not force and built_id() == sid:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    t0 = time.time()
    flags = NVCC_FLAGS + [f'-DSBD_BUILD_ID="{sid}"'] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run(_nvcc() + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        return src, obj, r

    objs, logs = [], []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, obj, r in ex.map(compile_one, sources()):
            logs.append(r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            objs.append(obj)
    if verbose:
        print("\n".join(logs))
    r = subprocess.run(_nvcc() + ["-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    ver = subprocess.run(_nvcc()[:1] + ["--version"], capture_output=True, text=True).stdout.strip().splitlines()
    with open(RECORD, "w") as fh:
        json.dump({"build_id": sid, "forced": bool(force), "seconds": round(time.time() - t0, 1),
                   "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                   "nvcc": ver[-1] if ver else None, "flags": NVCC_FLAGS,
                   "sources": [os.path.basename(s) for s in sources()]}, fh, indent=1)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))

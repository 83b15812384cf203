"""Builds libglasdi_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.environ.get("GLASDI_LIBDIR") or os.path.join(HERE, "lib")   # GLASDI_LIBDIR: a second (instrumented) build beside the product
LIB = os.path.join(LIBDIR, "libglasdi_b200.so")
SOURCES = ["api.cu", "rollout.cu", "decode_var2.cu", "decode_gram.cu", "decode_gram2.cu", "simt_fallback.cu", "refine.cu", "reduce.cu", "calibrate.cu", "gp.cu", "decode_wide.cu", "train_mlp.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def source_hash() -> str:
    """sha1 over the CUDA sources and the public header: compiled into the library (`glasdi_source_hash`) so that a
    library older than the sources is detected whatever the file times say (a snapshot copy resets them)."""
    import hashlib
    h = hashlib.sha1()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(HERE, "..", "include", "glasdi_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def built_hash() -> str:
    side = os.path.join(LIBDIR, "source_hash.txt")
    return open(side).read().strip() if os.path.exists(side) else "none"


def needs_build() -> bool:
    if not os.path.exists(LIB) or built_hash() != source_hash():
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "glasdi_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _stale(src: str, obj: str) -> bool:
    """An object is stale when its source or any header in csrc/ / include/ is newer."""
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if src == "api.cu":                                   # carries the hash of ALL sources
        deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu")]
    deps.append(os.path.join(HERE, "..", "include", "glasdi_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles the stale objects (all with force) in parallel and links the library.  GLASDI_NVCC_EXTRA adds flags
    (an instrumented build: pair it with force, the objects are shared)."""
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    shash = source_hash()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if not force and not _stale(src, obj):
            continue
        extra = [f'-DGLASDI_SOURCE_HASH="{shash}"'] if src == "api.cu" else []
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, *os.environ.get("GLASDI_NVCC_EXTRA", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(LIBDIR, "build.log"), "a" if not force else "w") as f:
        f.write("\n".join(log))
    if verbose or failed:
        print("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see gplasdi_b200/lib/build.log")
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(os.path.join(LIBDIR, "source_hash.txt"), "w") as f:      # what _lib.load compares before it opens the library
        f.write(shash + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

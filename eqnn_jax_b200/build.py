"""In-tree build of libeqnn_b200.so (nvcc, sm_100a only).

``python -m eqnn_jax_b200.build`` or ``__graft_entry__.build()``.  The shared
library lands next to this file so it travels with the repo snapshot to the GPU
box; nothing is written outside the repository.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeqnn_b200.so")
SOURCES = ["knn.cu", "csr.cu", "pairs.cu", "tpconv.cu", "gemm.cu", "elementwise.cu", "tc_mlp.cu", "rmlp.cu", "tc_gemm.cu", "node.cu", "egnn.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


EXTRA_SIGS = os.path.join(CSRC, "gen", "extra_signatures.json")


def extra_signatures():
    """Tensor-product signatures beyond codegen.PREBUILT, added on first use (add_tp_signature): [x irreps, lmax, live]."""
    import json
    if not os.path.exists(EXTRA_SIGS):
        return []
    return [tuple(t) for t in json.load(open(EXTRA_SIGS))]


def add_tp_signature(x_irreps: str, lmax: int, live: str) -> bool:
    """models/nequip.py:124 accepts any node irreps; the interaction kernels are generated per (x irreps, lmax, live
    irreps) signature.  Registers a new one, regenerates csrc/gen/tp_gen.cuh and rebuilds the library (nvcc, well under a
    minute).  Returns True if the library changed; the caller reloads it (_lib.reload)."""
    import json
    sigs = extra_signatures()
    key = (str(x_irreps), int(lmax), str(live))
    if key in sigs:
        return False
    sigs.append(key)
    os.makedirs(os.path.dirname(EXTRA_SIGS), exist_ok=True)
    json.dump([list(t) for t in sigs], open(EXTRA_SIGS, "w"))
    build_library()
    return True


def build_library(force: bool = False, verbose: bool = False) -> str:
    from . import codegen

    gen = os.path.join(CSRC, "gen", "tp_gen.cuh")
    codegen.generate(gen, list(codegen.PREBUILT) + extra_signatures())
    headers = [os.path.join(CSRC, "common.cuh"), gen, os.path.join(HERE, "..", "include", "eqnn_b200.h"),
               os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "rmlp_common.cuh")]
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_xla_shim(verbose: bool = False):
    """Compiles jax_binding/xla_ffi_shim.cc against jaxlib's XLA FFI headers WHEN THEY EXIST (they do not in the
    build image: returns None there).  The shim links against libeqnn_b200.so and lands next to ffi.py."""
    try:
        import jaxlib
    except Exception:
        return None
    inc = os.path.join(os.path.dirname(jaxlib.__file__), "include")
    if not os.path.exists(os.path.join(inc, "xla", "ffi", "api", "ffi.h")):
        return None
    src = os.path.join(HERE, "jax_binding", "xla_ffi_shim.cc")
    out = os.path.join(HERE, "jax_binding", "libeqnn_xla.so")
    if _stale(out, [src, LIB]):
        cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I" + inc, "-I/usr/local/cuda/include",
               "-I" + os.path.join(HERE, "..", "include"), src, "-L" + HERE, "-leqnn_b200", "-Wl,-rpath,$ORIGIN/..", "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"xla_ffi_shim build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print("xla shim:", build_xla_shim() or "skipped (no jaxlib headers)")

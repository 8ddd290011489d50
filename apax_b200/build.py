"""Build libapax_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library is a plain C-ABI shared object that links only cudart)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libapax_b200.so")
SOURCES = ["common.cu", "gm.cu", "nl.cu", "ctx.cu", "tc_gemm.cu", "i8_gemm.cu", "md.cu", "train.cu", "halo.cu", "ensemble.cu"]
HEADERS = ["common.cuh", "tc_gemm.cuh", "tc_ptx.cuh", "i8_gemm.cuh", "gm_internal.cuh", "nl_internal.cuh", "contract_gen.cuh", "gen_contract.py",
           os.path.join("..", "..", "include", "apax_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _stamp() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p):
            with open(p, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of the shared library."""
    stamp_file = LIB + ".stamp"
    # regenerate the contraction code from its single source of truth
    gen = os.path.join(CSRC, "contract_gen.cuh")
    subprocess.check_call([sys.executable, os.path.join(CSRC, "gen_contract.py"), gen])
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file):
        with open(stamp_file) as fh:
            if fh.read().strip() == stamp:
                return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs])  # static cudart (nvcc default)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

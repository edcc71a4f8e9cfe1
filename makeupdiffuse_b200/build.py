"""Builds libmkd_b200.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmkd_b200.so")
SOURCES = ["elementwise.cu", "norm.cu", "conv_generic.cu", "gemm_tcgen05.cu", "gemm_pair.cu", "attention.cu", "attention_tcgen05.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         *(["-DMKD_ENABLE_TRACE"] if os.environ.get("MKD_TRACE") == "1" else []),  # %globaltimer stamps + timing switches (tools/gemm_trace.py, tools/dbg_epilogue.sh)
         *([f"-DMKD_EPI_PIPE={os.environ['MKD_EPI_PIPE']}"] if os.environ.get("MKD_EPI_PIPE") else []),  # A/B builds of the GEMM epilogue
         *([f"-DMKD_MAX_STAGES={os.environ['MKD_MAX_STAGES']}"] if os.environ.get("MKD_MAX_STAGES") else []),  # pipeline-depth experiments
         *([f"-DMKD_GEGLU_PIPE={os.environ['MKD_GEGLU_PIPE']}"] if os.environ.get("MKD_GEGLU_PIPE") else []),  # A/B: GEGLU epilogue
         *os.environ.get("MKD_EXTRA_NVCC_FLAGS", "").split(),  # ablation builds (tools/), never set for the shipped library
         "-Xptxas", "-v"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mkd_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    if not os.path.exists(NVCC):
        raise RuntimeError(f"libmkd_b200.so is missing/stale and nvcc was not found at {NVCC}")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        r = subprocess.run([NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        results = list(ex.map(cc, SOURCES))
    log = []
    for src, obj, r in results:
        log.append(f"==== {src}\n{r.stderr}")
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    r = subprocess.run([NVCC, "-shared", "-o", OUT, *[o for _, o, _ in results], "-gencode",
                        "arch=compute_100a,code=sm_100a", "-cudart", "static"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

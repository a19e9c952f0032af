"""Build libspb200.so (sm_100a only) in-tree with nvcc.  ``python -m pytorch_superpoint_b200.build``."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libspb200.so")
SOURCES = ["warp.cu", "heat.cu", "peer.cu", "ha_fast.cu", "nms.cu", "desc.cu", "match.cu", "sampler.cu", "ingest.cu", "conv_simt.cu", "bn.cu", "conv_tc.cu", "conv_tc2.cu", "conv_s3.cu", "net.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr",
              "-Xptxas", "-v" if os.environ.get("SPB200_VERBOSE") else "-warn-spills"]
if os.environ.get("SPB200_DEFINES"):  # e.g. "-DSPB_C1A_TAP=6" for A/B experiments
    NVCC_FLAGS += os.environ["SPB200_DEFINES"].split()


def _newer(a: str, b: str) -> bool:
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "spb200.h")]
    if not force and os.path.exists(LIB) and not any(_newer(d, LIB) for d in deps):
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- nvcc {src} (rc={p.returncode})\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

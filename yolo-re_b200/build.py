"""Builds libyre.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python yolo-re_b200/build.py [--force]

Cross-compiles without a GPU.  The .so stays next to the python package (git-ignored, but it
travels to the GPU box with the working tree)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "yolo_b200" / "libyre.so"
OBJ = HERE / "build"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
# YRE_TUNING=1 at BUILD time compiles the YRE_TC_* / YRE_STEM_* environment knobs and the trace reader in
# (kernel tuning sessions only); the product build reads no environment variables.
if os.environ.get("YRE_TUNING", "0") not in ("", "0"):
    FLAGS.append("-DYRE_TUNING")
for _d in os.environ.get("YRE_DEFINES", "").split():       # experiment switches, e.g. YRE_DEFINES="YRE_SILU_F16X2"
    FLAGS.append("-D" + _d)
OUT = Path(os.environ.get("YRE_OUT", str(OUT)))              # build an experimental variant next to the product library


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "yre.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = OBJ / "digest.txt"
    dig = _digest()
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == dig:
        return OUT
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))

    def cc(src: Path):
        obj = OBJ / (src.stem + ".o")
        r = subprocess.run([NVCC, *FLAGS, "-c", str(src), "-o", str(obj)], capture_output=True, text=True)
        (OBJ / (src.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(cc, srcs))
    r = subprocess.run([NVCC, "-shared", "-cudart", "static", "-o", str(OUT), *map(str, objs)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)

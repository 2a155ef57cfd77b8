"""Build libelph_b200.so (sm_100a only) in-tree with nvcc.

The shared library is the product: hand-written CUDA kernels behind the C ABI of
``include/elph_b200.h``.  It links only against the CUDA runtime.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libelph_b200.so"
SOURCES = ["api.cu", "matvec.cu", "mtm_square.cu", "ssh_square.cu","cg.cu", "cg_persistent.cu", "cg_p2p.cu", "cg_pipe.cu", "pcg_fused.cu","fft.cu", "kpm.cu", "kpm_square.cu", "kpm_shard.cu", "force.cu", "dynamics.cu", "hmc.cu", "greens.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libelph_b200.so cannot be built")
    return cand


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES if (CSRC / s).exists()] + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "elph_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    srcs = [s for s in SOURCES if (CSRC / s).exists()]
    procs = []
    for s in srcs:
        obj = objdir / (s + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out:
            print(out)
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lpthread",
           "-Xlinker", f"--version-script={CSRC / 'exports.map'}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

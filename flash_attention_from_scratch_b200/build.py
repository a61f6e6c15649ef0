"""Build recipe for libfa_sm100.so (plain nvcc, no torch headers -> seconds, not minutes).

Replaces the reference's torch CUDAExtension build (/root/reference/setup.py:14-75), which
compiles 85 template instantiations for sm_80 in ~7 minutes.  The library is built IN-TREE
(flash_attention_from_scratch_b200/csrc/libfa_sm100.so) so it travels with the source snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libfa_sm100.so"
SOURCES = [CSRC / "fa_api.cu"]
HEADERS = [CSRC / "fa_fwd_sm100.cuh", CSRC / "fa_fwd_pp_sm100.cuh", CSRC / "ptx_sm100.cuh", CSRC / "softmax_sm100.cuh",
           CSRC.parent.parent / "include" / "fa_sm100.h"]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xptxas", "-v,-warn-spills",
    "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: Path = LIB_PATH) -> Path:
    """Compile the shared library if missing or older than its sources; return its path."""
    if not force and out == LIB_PATH and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", str(out), *map(str, SOURCES)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    (CSRC / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if verbose:
        print(log)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{log[-4000:]}")
    return out


def build_guard() -> Path:
    """Bring-up variant: bounded mbarrier spins that trap with a diagnostic instead of hanging."""
    return build(force=True, extra_flags=("-DFA_HANG_GUARD=1",), out=CSRC / "libfa_sm100_guard.so")


if __name__ == "__main__":
    print(build(force=True, verbose=True))

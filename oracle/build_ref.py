#!/usr/bin/env python
"""TEST / COMPARATOR INFRASTRUCTURE -- never imported by the product package.

Compiles the reference's own CUDA extension (`/root/reference/src/flash_attention.cu`, the pybind module
`flash_attention_kernels`, /root/reference/setup.py:14-63) for sm_100 from the sources where they lie.
Nothing is copied into the repo: the only output is `oracle/_ref/flash_attention_kernels.so` (git-ignored,
travels to the GPU box).  The recipe is setup.py's nvcc flag list with two changes: `-gencode
arch=compute_100,code=sm_100` instead of sm_80 (the shipped wheel would not load on a B200) and no `--keep`
(it would write intermediates next to the read-only sources).  One translation unit, 85 kernel
instantiations: ~7 minutes.

The reference kernel has no CPU implementation, so this is not a CPU oracle; it is (1) the same-box GPU
comparator SURVEY.md 8(d) asks for (kernel 16, legacy mma.sync path recompiled) and (2) a second parity
anchor on the GPU: tests/test_reference_kernel_gpu.py compares this repo's kernel with the reference's own
kernel on the same inputs (n_heads must be 16, static_kernel_configuration.cuh:146).
"""
import os
import subprocess
import sys
import sysconfig
from pathlib import Path

REF = Path(os.environ.get("FA_REFERENCE_ROOT", "/root/reference"))
OUT_DIR = Path(__file__).resolve().parent / "_ref"
OUT = OUT_DIR / "flash_attention_kernels.so"


def main() -> int:
    if not (REF / "src" / "flash_attention.cu").exists():
        print(f"[build_ref] {REF} not present: keeping any prebuilt {OUT}")
        return 0
    if OUT.exists() and "--force" not in sys.argv:
        print(f"[build_ref] {OUT} exists (use --force to rebuild)")
        return 0
    from torch.utils import cpp_extension as ce

    OUT_DIR.mkdir(exist_ok=True)
    inc = [str(REF / "src" / "include"), sysconfig.get_paths()["include"], *ce.include_paths("cuda")]
    libdirs = ce.library_paths("cuda")
    cmd = ["nvcc", "-std=c++20", '-Xcudafe=--diag_suppress=3189', "--use_fast_math", "--generate-line-info",
           "--expt-relaxed-constexpr", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
           "-U__CUDA_NO_HALF2_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
           "--ftemplate-backtrace-limit=0", "-O3", "-gencode", "arch=compute_100,code=sm_100",
           "-DTORCH_EXTENSION_NAME=flash_attention_kernels", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=1", "-shared", "-Xcompiler", "-fPIC,-O3",
           *[f"-I{p}" for p in inc], str(REF / "src" / "flash_attention.cu"), "-o", str(OUT),
           *[f"-L{p}" for p in libdirs], *[f"-Xlinker=-rpath={p}" for p in libdirs],
           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart", "-lcuda"]
    print("[build_ref]", " ".join(cmd), flush=True)
    return subprocess.call(cmd, cwd="/tmp")


if __name__ == "__main__":
    sys.exit(main())

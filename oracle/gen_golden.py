#!/usr/bin/env python
"""Generates tests/golden/*.pt by running the REFERENCE's own Python code on seeded inputs.

Runs only in the build container (needs /root/reference); the fixtures travel, this import does
not.  The reference modules import CUDA-only packages at module scope
(/root/reference/py/flash_helpers/test/utils.py:6-7, tools/debug/debug.py:7-12), so inert stubs
are registered for those names first -- none of the stubbed symbols is reached by the two
functions used here:

  * py_flash_attention(q, k, v, upcast)   utils.py:137-162   (the reference's test oracle)
  * block_flash_attention(...)            debug.py:40-153    (block-wise kernel emulation)
"""
import importlib.util
import io
import os
import sys
import types

import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, seed, dtype, (B, N, H, D)
    ("cfg1_bf16_1x128x2", 0, torch.bfloat16, (1, 128, 2, 128)),   # BASELINE.json configs[0]
    ("bf16_2x256x3", 1, torch.bfloat16, (2, 256, 3, 128)),
    ("fp16_2x256x3", 2, torch.float16, (2, 256, 3, 128)),
    ("bf16_1x512x1", 3, torch.bfloat16, (1, 512, 1, 128)),
    # long enough for AUTO to pick the CTA-pair kernel (seq_len > 1024): 10 KV blocks, 3 work tiles of 512 rows,
    # the last one half empty
    ("bf16_1x1280x1", 4, torch.bfloat16, (1, 1280, 1, 128)),
]


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    for name in ("flash_attn_2_cuda", "flash_attn_3_cuda", "flash_attention", "wurlitzer",
                 "flash_attn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["flash_attn"].flash_attn_func = None
    sys.modules["wurlitzer"].pipes = None
    sys.path.insert(0, os.path.join(REF, "py"))
    ref_utils = _load(os.path.join(REF, "py/flash_helpers/test/utils.py"), "ref_test_utils")
    ref_debug = _load(os.path.join(REF, "tools/debug/debug.py"), "ref_debug")
    os.makedirs(OUT, exist_ok=True)
    for name, seed, dtype, shape in CASES:
        g = torch.Generator().manual_seed(seed)
        q = torch.randn(shape, generator=g).to(dtype)
        k = torch.randn(shape, generator=g).to(dtype)
        v = torch.randn(shape, generator=g).to(dtype)
        fix = {
            "seed": seed, "shape": shape, "dtype": str(dtype), "q": q, "k": k, "v": v,
            "ref16": ref_utils.py_flash_attention(q, k, v, upcast=False),
            "ref32": ref_utils.py_flash_attention(q, k, v, upcast=True),
        }
        # block-wise emulation of one warp's 32 rows (debug.py picks warp_rank 2 of 4, B_r=128),
        # run in fp32 on the 16-bit inputs of (batch 0, head 0)
        B_r, B_c = 128, 64
        o_blk = ref_debug.block_flash_attention(
            128, q[0, :, 0].float(), k[0, :, 0].float(), v[0, :, 0].float(), B_r, B_c, io.StringIO())
        fix["blockwise_rows"] = (2 * B_r // 4, 3 * B_r // 4)
        fix["blockwise_fp32"] = o_blk.contiguous()
        path = os.path.join(OUT, name + ".pt")
        torch.save(fix, path)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

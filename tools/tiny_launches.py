#!/usr/bin/env python
"""Launches whose GPU duration is (almost) all fixed cost, to be run under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/tiny_launches.py
In each machine mapping: the debug instantiation at level 1 (set-up and tear-down only: barrier init, tensor-memory
allocation, cluster sync, de-allocation), then the production kernel on problems of one work tile with 1 / 2 / 4 KV
blocks and on 74 tiles of one block.  Development aid."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flash_attention_from_scratch_b200 as fa  # noqa: E402
from flash_attention_from_scratch_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    dump = torch.zeros(2 * 128 * 128 + 512 + 2048, device="cuda", dtype=torch.float32)
    diag = torch.zeros(256, device="cuda", dtype=torch.int32)
    for mode_name, mode in (("single", 1), ("pair", 2), ("pp", 3)):
        with _lib.thread_kernel_mode(mode):
            for (B, N, H) in ((1, 128, 1), (1, 256, 1), (1, 512, 1), (37, 128, 4)):
                q, k, v = (torch.randn(B, N, H, 128, device="cuda", dtype=torch.bfloat16) for _ in range(3))
                o = torch.empty_like(q)
                sb, sn, sh, _ = q.stride()
                if (B, N, H) == (1, 128, 1):
                    knobs = (C.c_uint32 * 8)(0, 0, 0, 0, 0, 0, 0, 1)
                    for _ in range(3):
                        rc = lib.fa_fwd_debug(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, N, H, 128, sb,
                                              sn, sh, 15, dump.data_ptr(), knobs, diag.data_ptr())
                        assert rc == 0, _lib.last_error()
                for _ in range(3):
                    fa.forward(None, q, k, v, o)
                torch.cuda.synchronize()
                print(mode_name, (B, N, H), "done", flush=True)


if __name__ == "__main__":
    main()

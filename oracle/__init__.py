"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's attention arithmetic.

Nothing under `flash_attention_from_scratch_b200/` may import this package: only `tests/`,
`__graft_entry__.smoke()`, the `flash_helpers.test` shim and `bench.py`'s CPU-baseline /
reference arm use it, and only as the checker or the timed CPU baseline.

Parity status: PINNED -- `tests/golden/*.pt` were produced by importing the reference's own
Python oracle (`py_flash_attention`, /root/reference/py/flash_helpers/test/utils.py:137-162) and
its block-wise emulation (`block_flash_attention`, /root/reference/tools/debug/debug.py:40-153) in
the build container with `oracle/gen_golden.py`; `tests/test_oracle.py` checks this restatement
against them bit for bit (py_flash_attention) / to fp32 round-off (block-wise).
"""
from .attention_ref import (  # noqa: F401
    blockwise_kernel_ref,
    ex2_emulated_ref,
    py_flash_attention,
    reference_pass_criterion,
    sdpa_ref,
)

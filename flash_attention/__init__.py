"""Drop-in module name of the reference operator package
(/root/reference/flash_attention/__init__.py:7-17): `flash_attention.forward(kernel_cfg, q, k, v,
o=None)` and `forward_timed(...)`, now backed by the B200 library."""
from flash_attention_from_scratch_b200.op import forward, forward_host, forward_timed  # noqa: F401

"""`flash_helpers.kernel_configs` as the reference's scripts import it
(/root/reference/py/flash_helpers/kernel_configs.py); re-exports the B200 package's module."""
from flash_attention_from_scratch_b200.kernel_configs import *  # noqa: F401,F403
from flash_attention_from_scratch_b200.kernel_configs import (  # noqa: F401
    DType,
    FlashForwardKernelConfig,
    calc_self_attn_flop,
    calc_total_flop,
    get_kernel_configs,
    get_kernels_to_build,
    parse_flash_forward_kernel_config,
    parse_kernel_name_into_config,
)

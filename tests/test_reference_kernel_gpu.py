"""Parity against the reference's OWN kernel, run on the same GPU.

`oracle/_ref/flash_attention_kernels.so` is the reference's CUDA extension (/root/reference/src/flash_attention.cu,
kernel 16: mma.sync / ldmatrix / cp.async) compiled for sm_100 by oracle/build_ref.py from the sources where they
lie; it travels to the GPU box as a built artefact.  The reference kernel hard-codes n_heads = 16
(/root/reference/src/include/static_kernel_configuration.cuh:146), so every shape here has 16 heads.  Both kernels
round P and O to the same 16-bit type at the same points but walk the KV blocks in
opposite order and the reference rescales O every block: equality is not promised, agreement within a couple of
16-bit ulps of O (|O| < 1 here) is, and both must satisfy the reference's criterion against the fp32 oracle.
Checker infrastructure: nothing under oracle/ is imported by the product path."""
import importlib.util
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

import flash_attention
from oracle import py_flash_attention, reference_pass_criterion

pytestmark = pytest.mark.gpu
REF_SO = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "flash_attention_kernels.so"


@pytest.fixture(scope="module")
def ref_kernels():
    if not REF_SO.exists():
        pytest.skip("oracle/_ref is not built (python oracle/build_ref.py, in the build container)")
    spec = importlib.util.spec_from_file_location("flash_attention_kernels", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_cfg(dtype):
    # the A100-best configuration of kernel 16 (kernel_sass/16_A100.asm header): (128, 64), 4 warps, load_2_2_0, buffer
    return SimpleNamespace(dtype=SimpleNamespace(to_torch_dtype=lambda: dtype), d_head=128, B_r=128, B_c=64, n_warps=4,
                           async_copy=True, eager_load_blocks=True, swizzled=True, Q_mma_load_K_tiles=2,
                           K_mma_load_K_tiles=2, V_mma_load_K_tiles=0, mma_double_buffer_loads=True,
                           optimized_softmax=True)


@pytest.mark.parametrize("mode", ["single", "pair", "pingpong"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,N", [(2, 512), (1, 2048), (1, 4096)])
def test_same_inputs_same_outputs_as_the_reference_kernel(ref_kernels, mode, dtype, B, N):
    from flash_attention_from_scratch_b200 import _lib

    g = torch.Generator(device="cuda:0").manual_seed(N + B)
    q, k, v = (torch.randn(B, N, 16, 128, device="cuda:0", dtype=dtype, generator=g) for _ in range(3))
    o_ref, _ = ref_kernels.forward(ref_cfg(dtype), q, k, v, None, False)
    prev = _lib.set_kernel_mode({"single": 1, "pair": 2, "pingpong": 3}[mode])
    try:
        o = flash_attention.forward(None, q, k, v)
        torch.cuda.synchronize()
    finally:
        _lib.set_kernel_mode(prev)
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11  # spacing of 16-bit values in [0.5, 1)
    diff = (o.float() - o_ref.float()).abs().max().item()
    assert diff <= 2 * ulp, (diff, ulp)
    ref16, ref32 = py_flash_attention(q, k, v, False), py_flash_attention(q, k, v, True)
    for name, out in (("ours", o), ("reference kernel", o_ref)):
        ok, d_out, d_ref = reference_pass_criterion(out, ref16, ref32)
        assert ok, (name, d_out, d_ref)

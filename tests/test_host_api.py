"""CPU tests of the host side: the C-ABI library builds for sm_100a, loads without a GPU, exports
every symbol include/fa_sm100.h declares; the operator's argument checks mirror the reference
launcher's error behaviour (/root/reference/src/flash_attention.cu:38-98).  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fa_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from flash_attention_from_scratch_b200 import _lib

    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_library_is_sm100a_with_tcgen05_and_tma(lib):
    from flash_attention_from_scratch_b200 import _lib

    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.lib_path())], capture_output=True,
                          text=True).stdout
    assert "EF_CUDA_SM100" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM"):
        assert mnemonic in sass, mnemonic
    assert "UTCHMMA.2CTA" in sass  # the CTA-pair kernel (tcgen05 cta_group::2)
    assert re.search(r"(?<![A-Z])HMMA\.", sass) is None  # no legacy mma.sync path


PRODUCTION_KERNELS = {  # bf16, aligned shapes: the instantiations bench.py times
    "pair": "_ZN2fa18fa_fwd_kernel_pairILb1ELb0ELb0EEEv14CUtensorMap_stS1_S1_S1_NS_9FwdParamsENS_8FwdDebugE",
    "single": "_ZN2fa13fa_fwd_kernelILb1ELb0ELb0EEEv14CUtensorMap_stS1_S1_S1_NS_9FwdParamsENS_8FwdDebugE",
    "pingpong": "_ZN2fa2pp16fa_fwd_kernel_ppILb1ELb0ELb0EEEv14CUtensorMap_stS2_S2_S2_NS_9FwdParamsENS_8FwdDebugE",
}


@pytest.mark.parametrize("kernel,n_mma,max_r2ur,max_r2ur_before_mma", [("pair", 32, 100, 16), ("single", 32, 64, 12),
                                                                       ("pingpong", 16, 44, 12)])
def test_sass_properties_the_speed_depends_on(lib, kernel, n_mma, max_r2ur, max_r2ur_before_mma):
    """What a toolkit bump could silently undo: the issuing warps' operands must live in UNIFORM
    registers.  Generation 9 found ~26 R2UR (vector -> uniform moves) between every barrier wait and the first
    UTCHMMA of a group to be the pacing item of the whole kernel; generation 14 found the same pattern coming back
    in the CTA-pair kernel when ptxas ran out of uniform registers (fixed with one asm statement per MMA group).
    Pinned here, on the SASS of the in-tree library: no spills, the number of tcgen05.mma issue sites, the total
    R2UR count and the R2UR count between the last SYNCS...TRYWAIT and the first UTCHMMA of every issue group."""
    from flash_attention_from_scratch_b200 import _lib

    sass = subprocess.run(["cuobjdump", "-sass", "-fun", PRODUCTION_KERNELS[kernel], str(_lib.lib_path())],
                          capture_output=True, text=True).stdout
    ops = [re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().split(";")[0].strip()
           for ln in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln)]
    assert len(ops) > 1000, "kernel not found in the library"
    assert not any(o.startswith(("STL", "LDL")) for o in ops), "register spills"
    assert sum("UTCHMMA" in o for o in ops) == n_mma
    assert sum(o.startswith("R2UR") for o in ops) <= max_r2ur
    worst, since, in_group = 0, None, False
    for o in ops:
        if "SYNCS.PHASECHK" in o:
            since, in_group = 0, False
        elif "UTCHMMA" in o:
            if not in_group and since is not None:
                worst = max(worst, since)
            in_group = True
        elif since is not None and not in_group and o.startswith("R2UR"):
            since += 1
    assert worst <= max_r2ur_before_mma, worst
    for mnemonic in ("UTMALDG", "UTMASTG", "LDTM", "STTM", "MUFU.EX2", "FFMA2"):
        assert any(mnemonic in o for o in ops), mnemonic


def test_kernel_info(lib):
    from flash_attention_from_scratch_b200 import _lib

    info = _lib.kernel_info()
    assert info["rows_per_cta"] == 256 and info["tmem_cols"] == 512
    assert info["smem_bytes"] <= 227 * 1024
    assert _lib.launch_count() == 0 or torch.cuda.is_available()


def test_c_abi_argument_validation_without_gpu(lib):
    from flash_attention_from_scratch_b200 import _lib

    buf = C.create_string_buffer(4096 + 64)
    p = (C.addressof(buf) + 63) & ~63
    args = lambda **kw: (p, p, p, p, kw.get("B", 1), kw.get("N", 128), kw.get("H", 1),  # noqa: E731
                         kw.get("D", 128), kw.get("sb", 128 * 128), kw.get("sn", 128), kw.get("sh", 128),
                         kw.get("dtype", 15), None)
    assert lib.fa_fwd(*args(dtype=6)) == 1 and "Only fp16 and bf16" in _lib.last_error()
    assert lib.fa_fwd(*args(D=64)) == 2 and "Kernel configuration was not found" in _lib.last_error()
    assert lib.fa_fwd(*args(N=(1 << 24) + 1)) == 3 and "out of range" in _lib.last_error()
    assert lib.fa_fwd(*args(B=0)) == 4
    assert lib.fa_fwd(*args(sn=100)) == 4 and "strides" in _lib.last_error()
    a = list(args())
    a[0] = None
    assert lib.fa_fwd(*a) == 4
    if not torch.cuda.is_available():
        # valid arguments but no device: must fail loudly, never fall back
        rc = lib.fa_fwd(*args())
        assert rc in (5, 7), rc
        assert _lib.last_error() != ""


def test_operator_checks_mirror_reference_messages():
    import flash_attention
    from flash_helpers.kernel_configs import DType, FlashForwardKernelConfig

    cfg = FlashForwardKernelConfig(dtype=DType.BF16)
    q = torch.zeros(1, 128, 2, 128, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        flash_attention.forward(cfg, q, q, q)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback_exists():
    import flash_attention_from_scratch_b200 as pkg

    src = open(os.path.join(os.path.dirname(pkg.__file__), "op.py")).read()
    assert "oracle" not in src and "scaled_dot_product_attention" not in src


def test_kernel_configs_surface():
    from flash_helpers.kernel_configs import (DType, FlashForwardKernelConfig, calc_self_attn_flop,
                                              get_kernel_configs, get_kernels_to_build,
                                              parse_kernel_name_into_config)

    assert DType.FP16 == 5 and DType.BF16 == 15
    assert DType.BF16.to_torch_dtype() == torch.bfloat16
    assert DType.from_string("bf16") == DType.BF16 and DType.from_string("5") == DType.FP16
    cfgs = get_kernels_to_build()
    assert {c.dtype for c in cfgs} == {DType.FP16, DType.BF16}
    for c in cfgs:
        assert parse_kernel_name_into_config(str(c)) == c
    assert get_kernel_configs("128,128") == cfgs and get_kernel_configs("64,64") == []
    # the Blackwell tuning grid (KERNELS=tune): the three machine mappings per dtype, round-tripping names
    tune = get_kernel_configs("tune")
    assert sorted((c.dtype, c.cta_group) for c in tune) == sorted(
        (dt, cg) for dt in (DType.FP16, DType.BF16) for cg in (1, 2, 3))
    for c in tune:
        assert f"cta{c.cta_group}" in str(c) and parse_kernel_name_into_config(str(c)) == c
    assert tune[0].kernel_name() in ("fa_fwd_kernel", "fa_fwd_kernel_pair", "fa_fwd_kernel_pp")
    assert {c.kernel_name() for c in tune} == {"fa_fwd_kernel", "fa_fwd_kernel_pair", "fa_fwd_kernel_pp"}
    assert all(c.cta_group == 0 and "cta" not in str(c) for c in cfgs)
    # FLOP model of the reference README (kernel_configs.py:102-103)
    assert calc_self_attn_flop(4, 32, 4096, 128) == 4 * 32 * (4 * 4096**2 * 128 + 6 * 4096**2)
    assert FlashForwardKernelConfig(dtype=DType.BF16).total_flop(4, 32, 4096) == 4 * 4 * 32 * 4096**2 * 128


def test_forward_host_rejects_a_bad_output_buffer():
    """fa_fwd_host copies batch*seq*heads*128 elements into `o`; anything but a contiguous host tensor of
    q's shape and dtype would be written out of bounds, so the operator refuses it before the library is called."""
    import torch

    from flash_attention_from_scratch_b200 import op

    q = torch.zeros(1, 128, 2, 128, dtype=torch.bfloat16)
    for bad in (torch.zeros(1, 64, 2, 128, dtype=torch.bfloat16),              # too small
                torch.zeros(1, 128, 2, 128, dtype=torch.float16),              # wrong dtype
                torch.zeros(1, 128, 4, 128, dtype=torch.bfloat16)[:, :, ::2],  # not contiguous
                ):
        with pytest.raises(RuntimeError, match="contiguous host tensor"):
            op.forward_host(q, q, q, bad)
    with pytest.raises(RuntimeError, match="was not found"):
        op.forward_host(q[..., :64].contiguous(), q[..., :64].contiguous(), q[..., :64].contiguous())


def test_auto_kernel_choice_is_grid_aware(lib):
    """AUTO (csrc/fa_api.cu: pick_kernel) without a GPU: ping-pong up to seq_len 1024, CTA pairs above -- except where
    the pair kernel's 512-row work tiles leave the machine under-filled and half-size tiles need fewer waves."""
    from flash_attention_from_scratch_b200 import _lib

    pk = lambda B, N, H: _lib.pick_kernel(N, B, H, 148)  # noqa: E731
    # BASELINE's shapes: the reference's sweep, the headline, the 8-way shard of config 5
    assert [pk(16, n, 16) for n in (512, 1024)] == ["fa_fwd_kernel_pp"] * 2
    assert [pk(b, n, 16) for b, n in ((16, 2048), (16, 4096), (8, 8192), (4, 16384))] == ["fa_fwd_kernel_pair"] * 4
    assert pk(4, 4096, 32) == "fa_fwd_kernel_pair" and pk(1, 16384, 32) == "fa_fwd_kernel_pair"
    # 16 pair tiles on 74 CTA pairs: one wave either way, the half-size tile finishes in half the time
    assert pk(1, 2048, 4) == "fa_fwd_kernel_pp"
    # 400 pair tiles = 6 waves of 2 units; 800 half tiles = 11 waves of 1.05
    assert pk(2, 4096, 25) == "fa_fwd_kernel_pp"
    # 192 pair tiles = 3 waves (6 units); 384 half tiles = 6 waves (6.3 units)
    assert pk(1, 8192, 12) == "fa_fwd_kernel_pair"
    # an explicit mode wins over the rule
    with _lib.thread_kernel_mode(1):
        assert pk(1, 2048, 4) == "fa_fwd_kernel"

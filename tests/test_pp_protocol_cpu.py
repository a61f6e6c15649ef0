"""Deadlock check of the ping-pong kernel's barrier protocol (tools/pp_protocol_sim.py) and of the shim's
demangled-name parser.  CPU only."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import pp_protocol_sim  # noqa: E402


def test_no_deadlock_for_any_block_or_tile_count():
    for n in range(1, 10):
        for tiles in range(1, 6):
            for ks, vs in ((4, 4), (2, 2), (1, 4), (4, 1)):
                assert pp_protocol_sim.simulate(n, tiles, ks, vs), (n, tiles, ks, vs)


def test_a_missing_signal_is_detected():
    """The checker must be able to fail: an epilogue that never frees O for the next tile hangs every problem
    with more than one tile."""
    src = Path(pp_protocol_sim.__file__).read_text()
    assert src.count('            sig("o_free")\n') == 1
    ns = {}
    exec(compile(src.replace('            sig("o_free")\n', ""), "pp_protocol_sim_sabotaged", "exec"), ns)
    assert ns["simulate"](4, 1) and not ns["simulate"](4, 2)


def test_demangled_kernel_names_parse():
    from flash_helpers.kernel_configs import DType, parse_flash_forward_kernel_config

    c = parse_flash_forward_kernel_config("void fa::pp::fa_fwd_kernel_pp<(bool)1, (bool)0, (bool)0>(CUtensorMap_st)")
    assert c.dtype == DType.BF16 and c.cta_group == 3 and c.kernel_name() == "fa_fwd_kernel_pp"
    c = parse_flash_forward_kernel_config("void fa::fa_fwd_kernel_pair<(bool)0, (bool)0, (bool)1>(x)")
    assert c.dtype == DType.FP16 and c.cta_group == 2
    c = parse_flash_forward_kernel_config("void fa::fa_fwd_kernel<(bool)1, (bool)0, (bool)0>(x)")
    assert c.cta_group == 1 and c.kernel_name() == "fa_fwd_kernel"

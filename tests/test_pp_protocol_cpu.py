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


def test_shared_s_protocol_has_no_deadlock_and_keeps_parity_discipline():
    """Generation 15 (epilogue warpgroup) and generation 14 protocols of csrc/fa_fwd_sm100.cuh, every block / tile
    count, both ring depths (pair: 4 + 4 slots, single: 2 + 2), with and without the lazy-rescale wait, and with
    each agent in turn as the slow one."""
    for n in range(1, 8):
        for tiles in range(1, 5):
            for ring in (2, 4):
                for rescale in (False, True):
                    for epi in (True, False):
                        lazies = (None, "producer", "mma_s", "mma_pv", "wg0", "wg1") + (("epilogue",) if epi else ())
                        for lazy in lazies:
                            assert pp_protocol_sim.simulate_shared_s(n, tiles, ring, rescale, epi, lazy=lazy), \
                                (n, tiles, ring, rescale, epi, lazy)


def test_shared_s_checker_catches_the_two_mistakes_generation_15_could_have_made():
    src = Path(pp_protocol_sim.__file__).read_text()
    # (a) an epilogue warpgroup that never hands O_s back: the next tile's first PV_s waits for ever
    bad = src.replace('                sig("o_free%d" % s)\n\n    agents', '\n    agents')
    assert bad != src
    ns = {}
    exec(compile(bad, "sim_no_o_free", "exec"), ns)
    assert ns["simulate_shared_s"](4, 1) and not ns["simulate_shared_s"](4, 2)
    # (b) the first P of a tile not waiting for the previous tile's last PV (the softmax warpgroup's own epilogue
    # used to guarantee it): no deadlock, but P_s runs ahead of the PV that still reads it
    bad = src.replace("if (g > 0) if epi_wg else (j > 0):", "if j > 0:")
    assert bad != src
    ns = {}
    exec(compile(bad, "sim_no_p_wait", "exec"), ns)
    assert ns["simulate_shared_s"](4, 1)
    assert not ns["simulate_shared_s"](1, 3, lazy="mma_pv")


def test_demangled_kernel_names_parse():
    from flash_helpers.kernel_configs import DType, parse_flash_forward_kernel_config

    c = parse_flash_forward_kernel_config("void fa::pp::fa_fwd_kernel_pp<(bool)1, (bool)0, (bool)0>(CUtensorMap_st)")
    assert c.dtype == DType.BF16 and c.cta_group == 3 and c.kernel_name() == "fa_fwd_kernel_pp"
    c = parse_flash_forward_kernel_config("void fa::fa_fwd_kernel_pair<(bool)0, (bool)0, (bool)1>(x)")
    assert c.dtype == DType.FP16 and c.cta_group == 2
    c = parse_flash_forward_kernel_config("void fa::fa_fwd_kernel<(bool)1, (bool)0, (bool)0>(x)")
    assert c.cta_group == 1 and c.kernel_name() == "fa_fwd_kernel"

"""CPU tests of the benchmark tooling that mirrors the reference's tools/benchmark."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


NCU_LOG = '''==PROF== Connected to process 1
==PROF== Disconnected from process 1
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","1","python","h","void at::native::vectorized_elementwise_kernel<4>(int)","1","7","(128, 1, 1)","(8, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","3.2"
"1","1","python","h","void fa::fa_fwd_kernel_pair<(bool)1, (bool)0, (bool)0>(CUtensorMap_st, fa::FwdParams)","1","7","(384, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","800"
"1","1","python","h","void fa::fa_fwd_kernel_pair<(bool)1, (bool)0, (bool)0>(CUtensorMap_st, fa::FwdParams)","1","7","(384, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Mbyte","403.0"
"1","1","python","h","void fa::fa_fwd_kernel_pair<(bool)1, (bool)0, (bool)0>(CUtensorMap_st, fa::FwdParams)","1","7","(384, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed","%","75.1"
"2","1","python","h","void fa::fa_fwd_kernel_pair<(bool)1, (bool)0, (bool)0>(CUtensorMap_st, fa::FwdParams)","1","7","(384, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ms","0.9"
'''


def test_ncu_bench_parses_and_summarises_an_ncu_log():
    nb = _load("tools/benchmark/ncu_bench.py", "ncu_bench")
    per_kernel = nb.parse_ncu_csv(NCU_LOG)
    assert list(per_kernel) == ["fa::fa_fwd_kernel_pair<"] or len(per_kernel) == 1
    (name, metrics), = per_kernel.items()
    assert "fa_fwd_kernel_pair" in name
    assert metrics["gpu__time_duration.sum"] == [800e3, 900e3]  # normalised to ns
    assert metrics["dram__bytes_read.sum"] == [403e6]            # normalised to bytes
    rows = nb.summarise(per_kernel, seq_len=4096, batch=16)
    assert len(rows) == 1 and rows[0]["launches"] == 2
    assert abs(rows[0]["duration"] - 0.85) < 1e-9                # ms
    assert abs(rows[0]["dram_rd"] - 403.0) < 1e-9                # MB
    assert abs(rows[0]["tensor"] - 75.1) < 1e-9
    # the reference's FLOP model B*H*(4 N^2 d + 6 N^2) at (16, 4096, 16, 128)
    flop = 16 * 16 * (4 * 4096 * 4096 * 128 + 6 * 4096 * 4096)
    assert abs(rows[0]["tflops"] - flop / 0.85e-3 / 1e12) < 1e-6
    table = nb.format_table(rows)
    assert "tensor %" in table and "75.1" in table
    assert nb.format_table(rows, as_csv=True).count("\n") == 1


def test_ncu_bench_ignores_logs_without_a_csv_header():
    nb = _load("tools/benchmark/ncu_bench.py", "ncu_bench")
    assert nb.parse_ncu_csv("==PROF== nothing profiled\n") == {}


def test_sanity_check_tool_imports_and_lists_its_flags():
    # tools/debug/sanity_check.py mirrors the reference's flags (--small --diff --kernel)
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "debug", "sanity_check.py"), "--help"],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-500:]
    for flag in ("--small", "--diff", "--kernel"):
        assert flag in p.stdout


def test_pipeline_model_ranks_the_measured_variants_correctly():
    # tools/pipeline_model.py: with this round's measured latencies the dataflow model must put the
    # production schedule (shared S accumulator) ahead of generation 4b, as measured (1455 vs 1399
    # TFLOP/s), land within 10 % of the measured loop period (~2690 clk), and never beat the MMA bound
    pm = _load("tools/pipeline_model.py", "pipeline_model")
    gen9, g4b, psmem = (pm.run(v) for v in ("gen9", "g4b", "psmem"))
    assert gen9["period"] < g4b["period"]
    assert abs(gen9["period"] - 2690) / 2690 < 0.10
    assert psmem["period"] >= psmem["mma_bound"] - 1e-6 and psmem["period"] <= gen9["period"]

#!/usr/bin/env python
"""Nsight Compute sweep of the operator: one ncu pass per sequence length, one table.

A working equivalent of the reference's /root/reference/tools/benchmark/ncu_bench.py (its imports are
stale): same flags (`--d_heads --seq_lens --runs --csv --no_sort`), same idea
(wrap `run_kernels.py` in `ncu --csv`, average the runs, print duration / cycles / registers / L2 hit
rate and TFLOP/s per kernel), plus the Blackwell counters the roofline argument needs (tensor pipe and
MUFU utilisation, DRAM bytes).  Batch size and head count per sequence length are the reference's
benchmark shapes (/root/reference/py/flash_helpers/test/utils.py:9-17).

    python tools/benchmark/ncu_bench.py --seq_lens 1024,4096 --runs 2
    python tools/benchmark/ncu_bench.py --from_csv profiles/r01_g9_launches.csv      # parse only

ncu serialises and replays kernels: use the durations for SHARES and counters, never as bench values.
"""
import argparse
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from flash_helpers.kernel_configs import calc_self_attn_flop  # noqa: E402

# the reference's benchmark shapes: seq_len -> batch, with 16 heads
BATCH_FOR_SEQ_LEN = {512: 16, 1024: 16, 2048: 16, 4096: 16, 8192: 8, 16384: 4}
N_HEADS = 16

# logical name -> (ncu metric, column title, scale, format)
METRICS = {
    "duration": ("gpu__time_duration.sum", "dur (ms)", 1e-6, "{:.4f}"),
    "cycles": ("sm__cycles_elapsed.max", "cycles", 1.0, "{:.0f}"),
    "regs": ("launch__registers_per_thread", "regs", 1.0, "{:.0f}"),
    "l2_hit": ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0, "{:.1f}"),
    "tensor": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %", 1.0, "{:.1f}"),
    "mufu": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "MUFU %", 1.0, "{:.1f}"),
    "dram_rd": ("dram__bytes_read.sum", "DRAM rd (MB)", None, "{:.1f}"),
    "dram_wr": ("dram__bytes_write.sum", "DRAM wr (MB)", None, "{:.1f}"),
}
_UNIT_TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
_UNIT_TO_NS = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def parse_ncu_csv(text, kernel_regex="fa_fwd"):
    """ncu `--csv` (details page) -> {kernel name: {metric name: [value per launch, ...]}}.
    Durations are normalised to ns and byte counts to bytes, whatever unit ncu chose."""
    import re

    lines = text.splitlines()
    start = next((i for i, line in enumerate(lines) if line.startswith('"ID"')), None)
    if start is None:
        return {}
    pat = re.compile(kernel_regex)
    out = defaultdict(lambda: defaultdict(list))
    for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        name = row.get("Kernel Name", "")
        if not pat.search(name):
            continue
        try:
            val = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "")
        metric = row["Metric Name"]
        if metric.startswith("gpu__time_duration"):
            val *= _UNIT_TO_NS.get(unit, 1.0)
        elif "bytes" in metric:
            val *= _UNIT_TO_BYTES.get(unit, 1.0)
        short = name.split("(")[0].replace("void ", "").strip()
        out[short][metric].append(val)
    return {k: dict(v) for k, v in out.items()}


def summarise(per_kernel, seq_len=None, batch=None, n_heads=N_HEADS):
    """Mean of every tracked metric per kernel (+ TFLOP/s when the shape is known)."""
    rows = []
    for kernel, metrics in per_kernel.items():
        row = {"kernel": kernel, "launches": max((len(v) for v in metrics.values()), default=0)}
        for key, (name, _title, scale, _fmt) in METRICS.items():
            vals = metrics.get(name)
            if not vals:
                row[key] = None
                continue
            mean = sum(vals) / len(vals)
            row[key] = mean * (1e-6 if scale is None else scale)
        if seq_len and batch and row.get("duration"):
            row["tflops"] = calc_self_attn_flop(batch, n_heads, seq_len, 128) / (row["duration"] * 1e-3) / 1e12
            row["seq_len"] = seq_len
        rows.append(row)
    return rows


def format_table(rows, as_csv=False):
    cols = ["kernel", "seq_len", "launches", "tflops"] + list(METRICS)
    titles = {"kernel": "kernel", "seq_len": "seq_len", "launches": "n", "tflops": "TFLOP/s (ref. FLOP model)"}
    titles.update({k: v[1] for k, v in METRICS.items()})
    fmts = {"tflops": "{:.1f}"}
    fmts.update({k: v[3] for k, v in METRICS.items()})

    def cell(row, c):
        v = row.get(c)
        if v is None:
            return "-"
        return fmts[c].format(v) if c in fmts else str(v)

    table = [[titles[c] for c in cols]] + [[cell(r, c) for c in cols] for r in rows]
    if as_csv:
        return "\n".join(",".join(line) for line in table)
    widths = [max(len(line[i]) for line in table) for i in range(len(cols))]
    sep = "+".join("-" * (w + 2) for w in widths)
    lines = [" | ".join(x.ljust(w) for x, w in zip(line, widths)) for line in table]
    return "\n".join([lines[0], sep] + lines[1:])


def run_ncu(seq_len, batch, runs, dtype):
    target = [sys.executable, str(ROOT / "tools" / "benchmark" / "run_kernels.py"), "--seq_len", str(seq_len),
              "--batch", str(batch), "--n_heads", str(N_HEADS), "--dtype", dtype, "--n_runs", str(runs + 1)]
    cmd = ["ncu", "--csv", "--clock-control", "none", "-k", "regex:fa_fwd", "-s", "1", "-c", str(runs),
           "--metrics", ",".join(m[0] for m in METRICS.values())] + target
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"ncu failed ({res.returncode}): {res.stderr[-2000:]}")
    return res.stdout


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--d_heads", default="128", help="comma list; only 128 exists (as in the reference's final kernel)")
    ap.add_argument("--seq_lens", default="1024", help="comma list of sequence lengths")
    ap.add_argument("--runs", type=int, default=1, help="profiled launches per configuration (averaged)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--csv", action="store_true", help="print CSV instead of a table")
    ap.add_argument("--no_sort", action="store_true", help="keep the input order instead of sorting by duration")
    ap.add_argument("--from_csv", default="", help="parse an existing ncu --csv log instead of running ncu")
    a = ap.parse_args()
    if any(int(d) != 128 for d in a.d_heads.split(",")):
        sys.exit("Kernel configuration was not found: d_head must be 128")
    rows = []
    if a.from_csv:
        rows = summarise(parse_ncu_csv(Path(a.from_csv).read_text()))
    else:
        for seq_len in (int(x) for x in a.seq_lens.split(",")):
            batch = BATCH_FOR_SEQ_LEN.get(seq_len, max(1, 65536 // seq_len))
            rows += summarise(parse_ncu_csv(run_ncu(seq_len, batch, a.runs, a.dtype)), seq_len, batch)
    if not a.no_sort:
        rows.sort(key=lambda r: (r.get("seq_len") or 0, r.get("duration") or 0.0))
    print(format_table(rows, a.csv))
    out_dir = ROOT / "gpurun_out"
    if out_dir.is_dir() and os.access(out_dir, os.W_OK) and not a.from_csv:
        (out_dir / "ncu_bench.csv").write_text(format_table(rows, True) + "\n")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Extracts the metrics we track from an .ncu-rep (read on the CPU box with `ncu -i`) into a small
JSON + text summary under profiles/.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep r01_v2 [--headline]
(--headline also rewrites profiles/ncu_summary.json, which bench.py reads for roofline.traffic)"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "sm__cycles_elapsed.max",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__grid_size",
    "launch__block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, vals = rows[0], rows[1], rows[2:]
    kernels = []
    for v in vals:
        d = dict(zip(header, v))
        name = d.get("Kernel Name", "?")
        picked = {"kernel": name}
        for k in header:
            if any(k == key or k.startswith(key) for key in KEYS) or "tensor" in k and "pct" in k \
                    or k.startswith("smsp__average_warp") or "pipe_xu" in k or "utcmma" in k:
                picked[k] = d[k]
        kernels.append(picked)
    summ = {"report": rep, "kernels": kernels}
    if kernels:
        k0 = kernels[0]
        try:
            rd = float(k0["dram__bytes_read.sum"].replace(",", ""))
            wr = float(k0["dram__bytes_write.sum"].replace(",", ""))
            ui = header.index("dram__bytes_read.sum")
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[ui], 1)
            summ["dram_bytes_per_launch"] = (rd + wr) * mult
        except Exception as e:  # noqa: BLE001
            summ["dram_bytes_err"] = str(e)
    (ROOT / "profiles").mkdir(exist_ok=True)
    (ROOT / "profiles" / f"{tag}_ncu_summary.json").write_text(json.dumps(summ, indent=1))
    if "--headline" in sys.argv:  # only the headline capture feeds bench.py's roofline.traffic
        (ROOT / "profiles" / "ncu_summary.json").write_text(json.dumps(
            {"tag": tag, "dram_bytes_per_launch": summ.get("dram_bytes_per_launch")}, indent=1))
    for k in kernels:
        for kk, vv in k.items():
            print(f"{kk:90s} {vv}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the attention-forward hot path on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is ONE pass of the hot path (one launch of the fused attention forward) over one batch of
synthetic Q/K/V already resident in HBM.  Workloads (BASELINE.json `configs`):

  headline  bf16 (B=4, N=4096, H=32, d=128) per GPU            -- configs[1], the default
  sweep     bf16 H=16, N in {512..16384} with the reference's batch sizes, harmonic mean -- configs[2]
  dtype16k  fp16|bf16 (8, 8192, 16, 128) per GPU               -- configs[3]  (--dtype fp16)
  shard16k  bf16 global (8, 16384, 32, 128) split over the ranks -- configs[4] (strong scaling)

With N > 1 (torchrun, one rank per GPU) every rank owns its own (batch x heads) shard and launches
independently: there is NO collective on the data path; torch.distributed is used only for the
barrier and the max-over-ranks of the device time.  Default scaling is WEAK: each rank runs the
per-GPU headline problem (global batch = 4 N).

Prints ONE JSON line (rank 0) with the contract keys plus `roofline`, `cpu_baseline`, `e2e`, `clocks`,
`gpu_launches`, `back_to_back` (the same K launches with nothing between them) and `sustained` (>= 2 s of them: the
power-capped regime, with >= 100 NVML samples of clock and power).  `value` follows BASELINE.md section 2: an L2
flush before every launch, one cudaEvent pair around each single launch, mean of the K pairs.  `--impl reference`
times the reference-side CPU implementation of the path (torch CPU attention = the reference's own test oracle
family, see oracle/) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "attention TFLOPs @ seq_len=4096 d_head=128; % of B200 bf16 tensor-core peak"
UNIT = "TFLOP/s"

# the reference's benchmark shapes: BATCH_SIZE_FOR_SEQ_LEN and BENCHMARK_N_HEADS = 16
# (/root/reference/py/flash_helpers/test/utils.py:9-17)
SWEEP_SHAPES = [(16, 512, 16, 128), (16, 1024, 16, 128), (16, 2048, 16, 128), (16, 4096, 16, 128),
                (8, 8192, 16, 128), (4, 16384, 16, 128)]

WORKLOADS = {
    # name: (B, N, H, d, default dtype, scaling)
    "headline": (4, 4096, 32, 128, "bf16", "weak"),
    "sweep": (16, 4096, 16, 128, "bf16", "weak"),  # placeholder shape: see SWEEP_SHAPES
    "dtype16k": (8, 8192, 16, 128, "bf16", "weak"),
    "shard16k": (8, 16384, 32, 128, "bf16", "strong"),
}


def matmul_flops(B, N, H, D):
    return 4.0 * B * H * N * N * D  # QK^T + PV, 2 flop per MAC


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    return 1590.0, 1400.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under
    NCCL_DEBUG=VERSION, for one), so fd 1 is pointed at stderr for the rest of the process and the result
    line goes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def kernel_name():
    """The kernel the library launched last on this thread (asked from the C ABI: fa_last_kernel)."""
    from flash_attention_from_scratch_b200 import _lib

    return {"fa_fwd_kernel": "fa::fa_fwd_kernel", "fa_fwd_kernel_pair": "fa::fa_fwd_kernel_pair (2-CTA clusters)",
            "fa_fwd_kernel_pp": "fa::pp::fa_fwd_kernel_pp (2-CTA clusters, KV ping-pong)"}.get(_lib.last_kernel(), "?")


def load_ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (or None)."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index, period=0.001):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples = []
        self.power_mw = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        n = 0
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                # an NVML call costs about a millisecond and the power reading is a ~1 s average: the fast sampler
                # (short timed regions) reads clocks and throttle reasons only
                if self.period >= 0.005:
                    self.power_mw.append(nv.nvmlDeviceGetPowerUsage(self.h))
                n += 1
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(s)}
        if self.power_mw and self.period >= 0.005:
            # (NVML's power reading is a ~1 s average: only the >= 2 s `sustained` region reports it)
            out["power_w_max"] = max(self.power_mw) / 1000.0
            out["power_w_mean"] = sum(self.power_mw) / len(self.power_mw) / 1000.0
        return out


def bind_to_gpu_numa_node(index):
    """Pins this process to the CPUs NVML reports as local to GPU `index`, BEFORE any pinned host buffer is
    allocated (first touch then places the pages on that NUMA node).  Returns a short description."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"cpus {cpus[0]}-{cpus[-1]} ({len(cpus)})"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"


def respawn_under_torchrun(n):
    """`python bench.py --gpus N` without a launcher: start the N ranks ourselves, the way the driver does."""
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    os.execv(sys.executable, cmd)


def cpu_attention_baseline(shape, dtype_name, reps, warmup=1):
    """Times torch CPU SDPA (all host threads) on `shape`; returns (tflops, cores, seconds/rep)."""
    import torch

    from oracle import sdpa_ref

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dt = torch.bfloat16 if dtype_name == "bf16" else torch.float16
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(shape, generator=g).to(dt) for _ in range(3))
    for _ in range(warmup):
        sdpa_ref(q, k, v)
    t0 = time.perf_counter()
    for _ in range(reps):
        sdpa_ref(q, k, v)
    dt_s = (time.perf_counter() - t0) / reps
    return matmul_flops(*shape) / dt_s / 1e12, cores, dt_s


def cpu_sample_shape(B, N, H, D):
    """Bounded sample of the workload for the CPU arm: one batch entry, at most 32 heads, capped so
    a rep stays within a few seconds (about 0.3 TFLOP of work)."""
    h = H
    while h > 1 and matmul_flops(1, N, h, D) > 0.3e12:
        h //= 2
    return (1, N, h, D)


def run_reference(args, rank, world):
    if rank != 0:
        return
    B, N, H, D, _, scaling = WORKLOADS[args.workload]
    shape = cpu_sample_shape(B, N, H, D)
    tflops, cores, sec = cpu_attention_baseline(shape, args.dtype, reps=max(1, args.steps),
                                                warmup=max(1, min(args.warmup, 2)))
    sample = f"torch CPU SDPA {args.dtype} on (B,N,H,d)={shape} of the {args.workload} workload, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": tflops, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{args.workload}: {args.dtype} (B,N,H,d)=({B},{N},{H},{D}); each step = "
                               f"a bounded sample {shape}", "timing": "host wall clock"},
        "cpu_baseline": {"value": tflops, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # 10 warm-ups, 50 reps, L2 flush before each, one event pair per launch, mean.
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default=None, choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    if args.dtype is None:
        args.dtype = WORKLOADS[args.workload][4]
    args.warmup = max(args.warmup, 3)

    if "WORLD_SIZE" not in os.environ and args.gpus > 1:
        respawn_under_torchrun(args.gpus)  # (before stdout is claimed: the ranks inherit the descriptors)
    claim_stdout()
    # the NVML sampler thread must get the interpreter often enough to see a 15 ms timed region more than once
    # (default switch interval: 5 ms)
    sys.setswitchinterval(2e-4)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and args.impl != "reference":
        raise SystemExit(f"bench.py: --gpus {args.gpus} but the launcher started WORLD_SIZE={world} ranks")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist

    import flash_attention_from_scratch_b200 as fa
    from flash_attention_from_scratch_b200 import _lib
    from flash_attention_from_scratch_b200.shard import shard_for_rank

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    numa = bind_to_gpu_numa_node(local_rank)  # before any pinned buffer exists (e2e leg)
    B, N, H, D, _, scaling = WORKLOADS[args.workload]
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def local_shape(B, N, H, D):
        if scaling == "weak":
            return (B, N, H, D), (B * world, N, H, D)
        sh = shard_for_rank(B, H, world, rank)
        return (sh.batch, N, sh.heads, D), (B, N, H, D)

    def make_sets(shape):
        """Rotating input sets, >= 512 MiB in total (L2 is 126 MB), so no step starts with its inputs cached by
        an earlier one.  The headline set (Q, K, V, O) is 512 MiB by itself: two sets."""
        set_bytes = 4 * shape[0] * shape[1] * shape[2] * shape[3] * 2
        n_sets = max(2, -(-(512 << 20) // set_bytes))
        gen = torch.Generator(device=dev).manual_seed(1000 + rank)
        sets = []
        for _ in range(n_sets):
            q, k, v = (torch.randn(shape, device=dev, dtype=dt, generator=gen) for _ in range(3))
            sets.append((q, k, v, torch.empty_like(q)))
        return sets

    # "L2 flush (>= 256 MB zero-fill) before each rep, one cudaEvent pair around the single
    # kernel launch, TFLOP/s from the mean".  The flush sits between two timed launches, outside their event pairs.
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def measure(sets, steps, warmup, flush=True):
        """W warm-ups, then exactly `steps` launches bracketed by barrier + synchronize, device time from cudaEvents on
        the launch stream, max over ranks.  flush=True: every launch has its own event pair and an L2 flush in front
        of it (ms/step = mean of the pairs); flush=False: back-to-back launches, ms/step = whole region / steps.
        Returns (ms/step, mean kernel ms, launches, clocks, region ms)."""
        def step(i):
            q, k, v, o = sets[i % len(sets)]
            fa.forward(None, q, k, v, o)

        for i in range(warmup):
            if flush:
                flush_buf.zero_()
            step(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        n0 = _lib.launch_count()
        e_a = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        e_b = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        for i in range(steps):
            if flush:
                flush_buf.zero_()
            e_a[i].record(stream)
            step(i)
            e_b[i].record(stream)
        e_a[steps].record(stream)
        barrier()
        launches = _lib.launch_count() - n0
        sampler.stop_flag.set()
        sampler.join()
        region_ms = e_a[0].elapsed_time(e_a[steps])
        per_launch = [e_a[i].elapsed_time(e_b[i]) for i in range(steps)]
        kern_ms = sum(per_launch) / len(per_launch)
        t = torch.tensor([kern_ms if flush else region_ms / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), kern_ms, int(launches), sampler.result(), region_ms

    def measure_sustained(sets, flops_per_launch, seconds=2.2, chunk=25):
        """The same launches back to back for >= `seconds`: the power-capped regime a long-running caller sees.
        NVML clock / power every 10 ms (>= 100 samples)."""
        def step(i):
            q, k, v, o = sets[i % len(sets)]
            fa.forward(None, q, k, v, o)

        barrier()
        sampler = ClockSampler(local_rank, period=0.01)
        sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(chunk):
                step(n)
                n += 1
            if n % (4 * chunk) == 0:
                torch.cuda.current_stream().synchronize()  # keep the launch queue bounded, never empty for long
        e1.record(stream)
        barrier()
        sampler.stop_flag.set()
        sampler.join()
        ms = e0.elapsed_time(e1)
        clk = sampler.result()
        return {"value": flops_per_launch * n / (ms * 1e-3) / 1e12, "unit": UNIT, "launches": n,
                "region_s": ms * 1e-3, "sm_mhz_median": clk.get("sm_mhz"), "power_w_mean": clk.get("power_w_mean"),
                "power_w_max": clk.get("power_w_max"), "nvml_samples": clk.get("samples"),
                "reasons": clk.get("reasons")}

    sweep_rows = None
    if args.workload == "sweep":
        # BASELINE config 3: one figure per sequence length, value = harmonic mean over the six
        sweep_rows = []
        launches = 0
        clocks = None
        for (sB, sN, sH, sD) in SWEEP_SHAPES:
            shape, gshape = local_shape(sB, sN, sH, sD)
            sets = make_sets(shape)
            ms_step, k_ms, n_l, clk, _ = measure(sets, args.steps, args.warmup)
            launches += n_l
            clocks = clk if sN == 4096 else clocks
            sweep_rows.append({"seq_len": sN, "batch": gshape[0], "n_heads": sH, "ms_per_step": ms_step,
                               "tflops": matmul_flops(*gshape) / (ms_step * 1e-3) / 1e12,
                               "kernel_tflops": matmul_flops(*shape) / (k_ms * 1e-3) / 1e12, "kernel": kernel_name(),
                               "sm_mhz": clk.get("sm_mhz"), "reasons": clk.get("reasons")})
            del sets
            torch.cuda.empty_cache()
        value = len(sweep_rows) / sum(1.0 / r["tflops"] for r in sweep_rows)
        ms_per_step = sum(r["ms_per_step"] for r in sweep_rows)  # one pass over the six shapes
        B, N, H, D = SWEEP_SHAPES[3]
    shape, gshape = local_shape(B, N, H, D)
    lB, lH = shape[0], shape[2]
    gB, gH = gshape[0], gshape[2]
    local_flops = matmul_flops(*shape)
    global_flops = matmul_flops(*gshape)
    sets = make_sets(shape)
    back_to_back = None
    if sweep_rows is None:
        ms_per_step, kern_ms, launches, clocks, region_ms = measure(sets, args.steps, args.warmup)
        value = global_flops / (ms_per_step * 1e-3) / 1e12
        # the same K launches with nothing between them (round 1's `value`): the board's power management pulls the
        # clock down within milliseconds under this kernel, so this figure falls with the length of the region
        b_ms, _, b_launches, b_clk, b_region = measure(sets, args.steps, 3, flush=False)
        back_to_back = {"value": global_flops / (b_ms * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": b_ms,
                        "steps": args.steps, "gpu_launches": b_launches, "region_ms": b_region, "sm_mhz": b_clk.get("sm_mhz"),
                        "reasons": b_clk.get("reasons")}
    else:
        kern_ms = local_flops / (sweep_rows[3]["kernel_tflops"] * 1e12) * 1e3
        region_ms = None
    # host cost of one call through the Python operator (argument checks, cached TMA descriptors, launch):
    # asynchronous calls, wall clock / calls -- the launch queue never fills, so this is pure host time
    torch.cuda.synchronize()
    n_host = 200 if kern_ms < 1.0 else 20  # (long kernels: keep the queued GPU time well under a second)
    t_h0 = time.perf_counter()
    for i in range(n_host):
        fa.forward(None, *sets[i % len(sets)])
    host_us = (time.perf_counter() - t_h0) / n_host * 1e6
    torch.cuda.synchronize()
    sustained_rec = None
    if not args.no_sustained:
        sustained_rec = measure_sustained(sets, global_flops)
        if world > 1:
            sustained_rec["note"] = "rank 0's clock; flops of all ranks over rank 0's device time"

    # ---------------------------------------------------------------- e2e: host buffers through the C ABI
    e2e = None
    if not args.no_e2e:
        hq, hk, hv = (torch.empty(shape, dtype=dt).pin_memory() for _ in range(3))
        for src, dst in zip(sets[0][:3], (hq, hk, hv)):
            dst.copy_(src)
        ho = torch.empty(shape, dtype=dt).pin_memory()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(3):
            fa.forward_host(hq, hk, hv, ho, device=local_rank)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fa.forward_host(hq, hk, hv, ho, device=local_rank)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        nbytes = hq.numel() * hq.element_size()
        e2e = {"value": global_flops / te.item() / 1e12, "unit": UNIT,
               "h2d_bytes_per_step": 3 * nbytes * world, "d2h_bytes_per_step": nbytes * world,
               "ms_per_step": te.item() * 1e3, "steps": e2e_steps,
               "host_numa_binding": numa,
               "path": "fa_fwd_host (C ABI): pinned host Q,K,V -> HBM, kernel, O -> pinned host; "
                       "batch-pipelined copies inside the timed region"}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N == 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample: one batch entry per rep, about 10-20 s of CPU work in total
        cshape = cpu_sample_shape(B, N, H, D)
        _, _, probe = cpu_attention_baseline(cshape, args.dtype, reps=1, warmup=1)
        reps = int(max(3, min(40, 12.0 / max(probe, 1e-3))))
        tf, cores, sec = cpu_attention_baseline(cshape, args.dtype, reps=reps, warmup=0)
        cpu = {"value": tf, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"torch CPU SDPA {args.dtype} on (B,N,H,d)={cshape} (1 of {B} batch entries of the "
                         f"workload per rep), 2 warm-ups + {reps} reps, {sec:.3f} s/rep"}

    if rank == 0:
        burst, sustained, how = load_peaks()
        achieved = local_flops / (kern_ms * 1e-3) / 1e12
        wl = (f"{args.workload}: {args.dtype} fused attention forward, per-GPU (B,N,H,d)=({lB},{N},{lH},{D}), "
              f"global ({gB},{N},{gH},{D}), non-causal, randn inputs")
        if sweep_rows is not None:
            wl = (f"sweep: {args.dtype} fused attention forward, H=16, d=128, (batch, seq_len) = "
                  "(16,512) (16,1024) (16,2048) (16,4096) (8,8192) (4,16384) per GPU "
                  "[the reference's benchmark shapes]; value = harmonic mean of the six TFLOP/s; roofline = the "
                  "seq_len 4096 point")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": wl,
                "flops_per_step": global_flops, "flop_model": "4*B*H*N^2*d",
                "l2_policy": "L2 flushed (256 MiB zero-fill) before every timed launch; inputs also rotate over "
                             ">= 512 MiB of 4-tensor sets (two at the headline), so no step re-reads a cached input",
                "parallelism": f"dp{world} over (batch x heads), no collective on the data path",
                "timing": "BASELINE.md section 2: one cudaEvent pair around each single launch on the launch stream, "
                          "ms_per_step = mean of the K pairs, barrier+synchronize on both sides of the K steps, max "
                          "over ranks; `back_to_back` = the same K launches with nothing between them, whole region "
                          "/ K; `sustained` = >= 2 s of them",
                "regime": (f"{args.steps} launches, each behind an L2 flush; the K steps span {region_ms:.1f} ms "
                           "including the flushes" if region_ms is not None else "per sequence length as above"),
            },
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": UNIT,
                         "frac": achieved / burst, "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({how}, "
                         f"cuBLAS burst; sustained {sustained})",
                         "frac_of_nominal_2250": achieved / 2250.0, "kernel": kernel_name(),
                         "kernel_ms": kern_ms, "traffic": load_ncu_traffic()},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "host_us_per_call": host_us,
            "tensor_map_cache": _lib.tensor_map_cache_stats(),
        }
        if back_to_back is not None:
            line["back_to_back"] = back_to_back
        if sustained_rec is not None:
            sustained_rec["peak_sustained"] = sustained
            sustained_rec["frac_of_cublas_sustained"] = sustained_rec["value"] / sustained
            line["sustained"] = sustained_rec
        if sweep_rows is not None:
            line["sweep"] = sweep_rows
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

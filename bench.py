#!/usr/bin/env python
"""Benchmark of the batched word-crop recognition forward pass (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 4096]

A "step" is one pass of the whole hot path - TPS rectification, ResNet, 2x BiLSTM, attention
greedy decode (all MAX_TEXT_LENGTH+1 steps, no early exit, every logit the reference computes) and
the semantic-vector embedding - over one batch of synthetic 1x32x100 crops with seeded random-init
weights.  One process per GPU; each rank owns ``--batch`` crops (weak scaling) and, when N > 1, the
only collective is one NCCL allgather of the decoded ids.

One JSON line on stdout (rank 0):
  value         crops/s, inputs resident in HBM, timed on the device with CUDA events, max over ranks
  e2e           same metric through Model.recognize with HOST buffers: every step copies its crops and
                semantic ids from pinned host memory (pipeline.HostBatchPrefetcher: on a copy stream, one step
                ahead) and reads the decoded ids / lengths back; all copies inside the timed region
  roofline      dominant kernel (tcgen05 implicit-GEMM convolutions): algorithmic FLOPs / per-launch
                CUDA-event time, against the measured cuBLAS bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the CPU port of the reference forward (oracle/torch_port.py: ATen CPU kernels, all
                host threads) on a bounded sample of the same workload (rank 0, N=1)
``--impl reference`` times only that CPU port and prints the same line shape with impl=reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "crops/sec TPS-ResNet-BiLSTM-Attn 32x100 bs4096"
UNIT = "crops/s"
WEIGHT_SEED, DATA_SEED = 999, 123
CONV_FLOP_PER_CROP = 10.443e9   # all convolutions, SURVEY.md 8d (ResNet 10.090 + LocNet 0.353)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0))),
                "bf16_burst": float(p.get("bf16_tflops", 1590.0)), "hbm": float(p.get("hbm_gbs", 6650.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # the first samples precede the load; the median over the busy part is what matters
        busy = [s for s, p in zip(sm, power) if p > 300.0] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU port of the reference (cpu_baseline and --impl reference)
# --------------------------------------------------------------------------------------------
def cpu_port_setup():
    from multimodal_scene_text_recognition_b200 import synth
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = tp.to_torch(synth.make_state_dict(WEIGHT_SEED))
    return tp, sd, synth, cores


def cpu_port_step(tp, sd, crops, sem):
    t0 = time.perf_counter()
    tp.forward(sd, crops, *sem)
    return time.perf_counter() - t0


def cpu_sample_inputs(synth, n):
    x = torch.from_numpy(synth.make_crops(DATA_SEED, n))
    sem = [torch.from_numpy(a) for a in synth.make_semantics(DATA_SEED + 1, n)]
    return x, sem


def cpu_baseline(budget_s: float = 15.0):
    """Port timed on the host cores on a bounded sample (~budget_s of CPU work)."""
    tp, sd, synth, cores = cpu_port_setup()
    x, sem = cpu_sample_inputs(synth, 32)
    cpu_port_step(tp, sd, x, sem)                      # warm-up (thread pool, oneDNN primitives)
    t = cpu_port_step(tp, sd, x, sem)
    n = int(min(1024, max(32, (budget_s * 0.6) * 32 / t)))
    n -= n % 32
    x, sem = cpu_sample_inputs(synth, n)
    best = min(cpu_port_step(tp, sd, x, sem) for _ in range(2))
    return {"value": round(n / best, 2), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of the 4096 crops of one step, best of 2 passes, torch %s CPU (oracle/torch_port.py), "
                      "fp32, all 26 decode steps" % (n, torch.__version__)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tp, sd, synth, cores = cpu_port_setup()
    x, sem = cpu_sample_inputs(synth, 32)
    cpu_port_step(tp, sd, x, sem)
    t = cpu_port_step(tp, sd, x, sem)
    total_steps = args.steps + args.warmup
    per_step_budget = max(0.5, 120.0 / total_steps)     # whole run ~2 minutes
    n = int(min(4096, max(32, per_step_budget * 32 / t)))
    n -= n % 32
    x, sem = cpu_sample_inputs(synth, n)
    for _ in range(args.warmup):
        cpu_port_step(tp, sd, x, sem)
    times = [cpu_port_step(tp, sd, x, sem) for _ in range(args.steps)]
    total = sum(times)
    value = n * args.steps / total
    sample = "%d-crop sample of the 4096-crop step per timed step, torch %s CPU port of the reference forward" % (
        n, torch.__version__)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "TPS-ResNet-BiLSTM-Attn greedy forward, 1x32x100 crops, 26 decode steps, "
                               "seeded random-init weights", "batch_per_gpu": args.batch, "sample_crops_per_step": n},
        "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from multimodal_scene_text_recognition_b200 import _cabi, config, model as M, ops, parallel, synth
    from multimodal_scene_text_recognition_b200.pipeline import HostBatchPrefetcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the recognition path has no CPU implementation "
                         "(use --impl reference for the CPU port)")
    _cabi.lib()   # fail loudly if the extension is missing
    rank, local_rank, world = parallel.init_from_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun (--nproc-per-node %d)" %
                             (args.gpus, args.gpus))
        raise SystemExit("WORLD_SIZE=%d does not match --gpus %d" % (world, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = _peaks()

    config.PRECISION = args.precision
    config.DECODER_EARLY_EXIT = False      # every step and logit the reference computes
    config.TPS_GRID = "reference"
    config.SEMANTIC_VECTOR = "overlap"
    B = args.batch
    sd = synth.make_state_dict(WEIGHT_SEED)
    model = M.load_numpy_state_dict(M.Model().eval(), sd).to(dev)

    # per-rank shard of the global batch (weak scaling): seeded by rank, two alternating host batches
    host_sets = []
    for k in range(2):
        x = torch.from_numpy(synth.make_crops(DATA_SEED + 17 * rank + 1000 * k, B)).pin_memory()
        sem = [torch.from_numpy(a).pin_memory() for a in synth.make_semantics(DATA_SEED + 1 + 17 * rank + 1000 * k, B)]
        host_sets.append((x, sem))
    dev_sets = [(x.to(dev), [s.to(dev) for s in sem]) for x, sem in host_sets]
    steps_out = config.MAX_TEXT_LENGTH + 1
    ids_host = torch.empty((B * world, steps_out), dtype=torch.int32).pin_memory()
    len_host = torch.empty((B * world,), dtype=torch.int32).pin_memory()

    def step_device(k):
        x, sem = dev_sets[k % 2]
        with torch.no_grad():
            out = model.recognize(x, *sem)
        if world > 1:
            return parallel.allgather_ids(out["ids"], out["lengths"], B * world)
        return out["ids"], out["lengths"]

    prefetcher = HostBatchPrefetcher(dev, depth=2)
    staged = {}

    def stage(k):
        hx, hsem = host_sets[k % 2]
        staged[k] = prefetcher.stage([hx] + hsem)

    def step_e2e(k):
        # every step copies its crops + semantic ids from pinned host memory (on the copy stream, enqueued
        # one step ahead so that it overlaps the previous step's compute) and reads the decoded ids back
        if k not in staged:
            stage(k)
        t0 = time.perf_counter()
        tensors, ready, sid = staged.pop(k)
        torch.cuda.current_stream().wait_event(ready)
        with torch.no_grad():
            out = model.recognize(tensors[0], *tensors[1:])
        prefetcher.release(sid)
        t1 = time.perf_counter()
        stage(k + 1)                                    # H2D of the next step's inputs, overlapped
        t2 = time.perf_counter()
        ids, lengths = out["ids"], out["lengths"]
        if world > 1:
            ids, lengths = parallel.allgather_ids(ids, lengths, B * world)
        ids_host.copy_(ids, non_blocking=True)
        len_host.copy_(lengths, non_blocking=True)
        t3 = time.perf_counter()
        torch.cuda.current_stream().synchronize()      # the caller needs the decoded ids of this step
        if os.environ.get("MSTR_BENCH_TRACE"):
            print("rank %d e2e step %d: launch %.1f stage %.1f gather+d2h %.1f sync %.1f ms" % (
                rank, k, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (time.perf_counter() - t3)),
                file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, trace=None):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for k in range(steps):
            t0 = time.perf_counter()
            fn(k)
            if trace is not None:
                trace.append(round(1e3 * (time.perf_counter() - t0), 2))
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for k in range(max(args.warmup, 3)):
        step_device(k)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _cabi.launch_count()
    ops.TIMER = ops.KernelTimer() if rank == 0 else None
    total_ms = timed(step_device, args.steps)
    timer, ops.TIMER = ops.TIMER, None
    launches = _cabi.launch_count() - launches0
    for k in range(2):                     # the host-buffer path has its own allocation pattern: warm it up too
        step_e2e(k)
    torch.cuda.synchronize()
    staged.clear()                         # the timed region starts with nothing staged: step 0 pays its copy in full
    e2e_trace = []
    e2e_ms = timed(step_e2e, args.steps, e2e_trace)
    if rank == 0:
        print("e2e per-step host wall ms: %s" % e2e_trace, file=sys.stderr)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = B * world * args.steps / (total_ms * 1e-3)
    e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
    kern = timer.summary()
    conv = kern.get({"bf16": "conv_gemm_bf16_2cta", "tf32": "conv_gemm_tf32x3"}.get(args.precision, "conv_gemm_f32"),
                    {"launches": 0, "ms": 0.0, "work": 0.0})
    conv_tflops = conv["work"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    peak = peaks["bf16_sustained"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    stage_ms = {k: round(v["ms"] / args.steps, 3) for k, v in sorted(kern.items())}
    if args.kernel_table:
        with open(args.kernel_table, "w") as f:
            json.dump(timer.table(args.steps), f, indent=1)
    h2d = sum(t.numel() * t.element_size() for t in [host_sets[0][0]] + host_sets[0][1])
    d2h = ids_host.numel() * 4 + len_host.numel() * 4
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "tf32": "tf32x3"}.get(args.precision, "f32"), "data": "synthetic",
        "config": {
            "workload": "TPS-ResNet-BiLSTM-Attn greedy forward (BASELINE.json configs[2]): 1x32x100 crops, "
                        "26 decode steps without early exit, 97 classes, overlap semantic vectors, seeded random-init weights",
            "batch_per_gpu": B, "global_batch": B * world, "precision": args.precision,
            "parallelism": "batch-sharded, 1 process/GPU, allgather(ids) only" if world > 1 else "single GPU",
            "l2": "two alternating input batches; per-step activation traffic (>1 GB) exceeds the 126 MB L2",
        },
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(e2e_ms / args.steps, 3)},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "tensor", "kernel": "conv_gemm_bf16_2cta_kernel<256> (tcgen05 cta_group::2 implicit GEMM; %d launches per step, "
                      "%.2f TFLOP and %.3f ms per launch on average)" % (
                          conv["launches"] // max(args.steps, 1), conv["work"] / max(conv["launches"], 1) / 1e12,
                          conv["ms"] / max(conv["launches"], 1)),
            "achieved": round(conv_tflops, 1), "peak": peak, "unit": "TFLOP/s",
            "frac": round(conv_tflops / peak, 4) if peak else None, "traffic": traffic,
            "peak_source": peaks["source"] + ", sustained cuBLAS bf16",
            "whole_step_frac": round(value / world * CONV_FLOP_PER_CROP / 1e12 / peak, 4),
            "all_conv_gemm_tflops": round(sum(v["work"] for k, v in kern.items() if k.startswith("conv_gemm_bf16")) /
                                          max(sum(v["ms"] for k, v in kern.items() if k.startswith("conv_gemm_bf16")), 1e-9) / 1e9, 1),
        },
        "stage_ms_per_step": stage_ms,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=4096, help="crops per GPU per step")
    ap.add_argument("--precision", choices=["bf16", "tf32", "fp32"], default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-table", default=None, help="write the per-call-site kernel table (JSON) to this path")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the EEG preprocessing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the fused DSP chain (FIR -> STFT log-power -> z-score)
over one batch of synthetic trials per GPU: BASELINE config 2,
256 trials x 64 channels x 2048 samples, n_fft 256, hop 64.  Ranks are
independent (trials shard by batch, no data-path collective): weak scaling.

Prints ONE JSON line (rank 0).  `value` is trials/s with inputs resident in HBM
(CUDA events, max over ranks); `e2e` is trials/s through the public API from
pinned HOST buffers with the H2D copy of the raw trials and the D2H read of the
features inside the timed region; `roofline` is the DSP kernel's algorithmic
bytes / measured launch time against the measured HBM peak; `cpu_baseline` is
the oracle's torch-CPU port of the same chain on the host cores (bounded sample).
`--impl reference` times that CPU port alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

B_PER_GPU, C, T = 256, 64, 2048
N_FFT, HOP = 256, 64
F, NF = N_FFT // 2 + 1, 1 + T // HOP
BYTES_PER_TRIAL = 4 * C * T + 4 * C * F * NF          # 1,614,080
METRIC = "EEG trials/sec (preprocess: FIR+STFT log-spectrogram+z-score, cfg2 256x64x2048)"
UNIT = "trials/s"


def workload_config(extra=None):
    cfg = {"workload": "BASELINE configs[1]: preprocessing-only, batch 256 x 64 ch x 2048 samples, "
                       "STFT n_fft=256 hop=64, FIR 65 taps 8-30 Hz, per GPU",
           "batch_per_gpu": B_PER_GPU, "channels": C, "samples": T, "n_fft": N_FFT, "hop": HOP,
           "l2": "inputs+outputs (413 MB/step, rotating 2 buffer sets) exceed the 126 MB L2"}
    if extra:
        cfg.update(extra)
    return cfg


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([p.strip() for p in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_chain(x, h):
    from oracle import preprocess_oracle as po   # bench.py's cpu_baseline / reference leg only
    return po.dsp_torch_cpu_f32(x, h, n_fft=N_FFT, hop=HOP)


def time_cpu(sample_b, budget_s, steps=None, warmup=1):
    """Oracle port (torch CPU fp32, all host threads) on a bounded sample."""
    from imagined_speech_translation_b200.preprocess import design_bandpass_fir
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    x = 20.0 * torch.randn(sample_b, C, T, generator=g)
    h = torch.from_numpy(design_bandpass_fir(65, (8.0, 30.0), 256.0))
    for _ in range(warmup):
        cpu_chain(x, h)
    times = []
    t_end = time.perf_counter() + budget_s
    while True:
        t0 = time.perf_counter()
        cpu_chain(x, h)
        times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and time.perf_counter() > t_end and len(times) >= 2:
            break
    return times


def run_reference(args, rank):
    if rank != 0:
        return
    sample_b = 32
    times = time_cpu(sample_b, budget_s=0.0, steps=args.steps, warmup=max(args.warmup, 1))
    ms = 1000.0 * float(np.mean(times))
    val = sample_b / (ms / 1000.0)
    cores = torch.get_num_threads()
    sample = (f"{sample_b} trials x {C} ch x {T} samples per step (1/8 of the per-GPU batch), "
              f"torch CPU fp32 F.conv1d + torch.stft, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config({"sample_batch": sample_b}),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import imagined_speech_translation_b200 as pkg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    fe = pkg.SpectrogramFrontEnd(C, T, {"n_fft": N_FFT, "hop": HOP})
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    nbuf = 2
    xs = [20.0 * torch.randn(B_PER_GPU, C, T, generator=g, device=dev) for _ in range(nbuf)]
    outs = [torch.empty(fe.out_shape(B_PER_GPU), device=dev) for _ in range(nbuf)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput + per-launch kernel time ----------------
    for i in range(args.warmup):
        fe(xs[i % nbuf], out=outs[i % nbuf])
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        fe(xs[i % nbuf], out=outs[i % nbuf])
        evs[i + 1].record()
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    # keep the clocks sampler alive for a minimum window so short runs still get samples
    if sampler:
        t_hold = time.perf_counter() + 0.6
        i = 0
        while time.perf_counter() < t_hold:
            fe(xs[i % nbuf], out=outs[i % nbuf]); i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        clocks = sampler.stop()
    else:
        clocks = None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = B_PER_GPU * world / (ms_per_step / 1000.0)

    # ---------------- end to end from pinned host memory through the public API ----------------
    h_in = [torch.empty(B_PER_GPU, C, T, pin_memory=True).normal_(0, 20.0) for _ in range(2)]
    h_out = [torch.empty(fe.out_shape(B_PER_GPU), pin_memory=True) for _ in range(2)]
    d_in = [torch.empty(B_PER_GPU, C, T, device=dev) for _ in range(2)]
    d_out = [torch.empty(fe.out_shape(B_PER_GPU), device=dev) for _ in range(2)]
    s_h2d, s_cmp, s_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    ev_free_in = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(n):
        # 3-stage software pipeline over two buffer sets: H2D(i+1) | DSP(i) | D2H(i-1)
        for i in range(n):
            k = i % 2
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(ev_free_in[k])      # d_in[k] consumed by the kernel of step i-2
                d_in[k].copy_(h_in[k], non_blocking=True)
                ev_in[k].record(s_h2d)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[k])
                s_cmp.wait_event(ev_out[k])          # d_out[k] drained by the D2H of step i-2
                fe(d_in[k], out=d_out[k])
                ev_cmp[k].record(s_cmp)
                ev_free_in[k].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_cmp[k])
                h_out[k].copy_(d_out[k], non_blocking=True)
                ev_out[k].record(s_d2h)

    e2e_steps(max(args.warmup, 2))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_steps(args.steps)
    for s in (s_h2d, s_cmp, s_d2h):
        torch.cuda.current_stream().wait_stream(s)
    t1.record()
    barrier()
    e = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    e2e_ms = float(e.item()) / args.steps
    e2e_val = B_PER_GPU * world / (e2e_ms / 1000.0)
    checksum = float(h_out[(args.steps - 1) % 2][0, 0].abs().sum())

    if rank == 0:
        peak, peak_src = measured_peak()
        kern_ms = float(np.mean(per_launch_ms))
        achieved = BYTES_PER_TRIAL * B_PER_GPU / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dsp_traffic.json")) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        except Exception:
            pass
        cpu_times = time_cpu(32, budget_s=12.0)
        cpu_val = 32 / float(np.mean(cpu_times))
        cores = torch.get_num_threads()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config({"kernel": fe.kernel_name}),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 4 * B_PER_GPU * C * T,
                    "d2h_bytes_per_step": 4 * B_PER_GPU * C * F * NF,
                    "path": "pinned host -> H2D -> SpectrogramFrontEnd (C ABI) -> D2H pinned host, "
                            "3-stream pipeline", "checksum": checksum},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": f"dsp_{fe.kernel_name}_kernel", "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": BYTES_PER_TRIAL * B_PER_GPU,
                         "frac_of_nominal_8TBps": achieved / 8000.0},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"32 trials x {C} ch x {T} samples x {len(cpu_times)} repeats "
                                       f"({sum(cpu_times):.1f} s), torch CPU fp32 F.conv1d + torch.stft"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

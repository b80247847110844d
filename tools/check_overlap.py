"""Data-parallel check of the all-reduce that overlaps backward (run under torchrun, >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_overlap.py [--batch 16] [--time]

From identical replicas and per-rank batches, one train step is run three ways -- one all-reduce after backward,
slices all-reduced from inside backward (eager), and the same captured in a CUDA graph -- and the flat gradient
buffers after the reduction are compared BIT FOR BIT (a two-rank sum does not depend on how the buffer is cut).
With more than two ranks the comparison is to 1e-6 relative (NCCL may pick different algorithms per message size).
--time also reports ms/step for both modes (graph replay, optimizer step included).
"""
import argparse
import os
os.environ.setdefault("EEGX_BART_RANDOM_INIT", "1")   # synthetic benchmark: reference architecture, random weights (no HF cache here)
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (shapes and the synthetic token generator)
import imagined_speech_translation_b200 as pkg  # noqa: E402
from imagined_speech_translation_b200 import trainer as tr  # noqa: E402
from imagined_speech_translation_b200.model import EEGDecodingModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fe = pkg.SpectrogramFrontEnd(bench.C, bench.T, {"n_fft": bench.N_FFT, "hop": bench.HOP})
    torch.manual_seed(0)
    model = EEGDecodingModel(n_timepoints=fe.n_frames,
                             region_channel_counts={k: v * fe.n_freqs for k, v in bench.COUNTS.items()})
    tr.initialize_custom_weights(model)
    model = model.to(dev).train()
    cfg = dict(tr.CONFIG, accumulation_steps=1)
    opt = tr.build_optimizer(model, cfg)
    sched = tr.cosine_schedule_with_warmup(opt, cfg["warmup_steps"], 100000)
    trainer = tr.EEGTrainer(model, None, None, None, opt, sched, cfg, front_end=fe, region_channel_counts=bench.COUNTS)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ids, labels = bench.synth_tokens(a.batch, g, dev)
    batch = {"raw": 20.0 * torch.randn(a.batch, bench.C, bench.T, generator=g, device=dev),
             "decoder_input_ids": ids, "labels": labels}

    from imagined_speech_translation_b200 import fused

    def reduced_grads(overlap: bool, graph: bool):
        """Gradients after the all-reduce of one step taken at a FIXED dropout state and fixed weights."""
        trainer.overlap_allreduce = overlap
        trainer._graph = None
        opt.zero_grad()
        fused.set_seed(77)
        if graph:
            trainer.capture(batch, warmup=0)
            fused.set_seed(77)
        trainer.train_step(batch)
        if not trainer._grads_reduced:
            from imagined_speech_translation_b200 import distributed as dp
            dp.allreduce_sum_(opt.flat_grads(), trainer.process_group)
            assert not overlap, "overlap requested but the step did not reduce its gradients"
        else:
            assert overlap
        trainer._grads_reduced = False
        torch.cuda.synchronize()
        out = opt._all_grads.clone()
        trainer._graph = None
        return out

    trainer.overlap_allreduce = False
    trainer.train_step(batch)                     # first step: builds the flat buffers (lr = 0 at step 0: weights unchanged)
    trainer.optimizer_step(step_scheduler=False)
    base = reduced_grads(False, False)
    plan = trainer._overlap_plan()
    n_total = opt._all_grads.numel()
    early = sum(hi - lo for k, runs in plan.items() if k != 'rest' for lo, hi in runs)
    results = {}
    for name, (ov, gr) in {"eager_overlap": (True, False), "graph_plain": (False, True),
                           "graph_overlap": (True, True)}.items():
        got = reduced_grads(ov, gr)
        if world == 2:
            ok = torch.equal(got, base)
        else:
            ok = bool(((got - base).abs().max() / base.abs().max()) <= 1e-6)
        results[name] = ok
    # plan / path mismatch: the plan was laid out for the lock-step region path, the batch takes the per-region one (as
    # happens for channel counts the stacked kernels do not serve).  The slices whose boundaries never fire must still
    # be reduced exactly once.
    enc = model.brain_encoder
    if any(isinstance(k, tuple) and k[1] == 'group' for k in plan):
        saved = enc._lock_step
        enc._lock_step = lambda eeg_data: None
        try:
            base_pr = reduced_grads(False, False)
            got_pr = reduced_grads(True, False)
        finally:
            enc._lock_step = saved
        results["per_region_path_under_lock_step_plan"] = (torch.equal(got_pr, base_pr) if world == 2 else
                                                           bool(((got_pr - base_pr).abs().max() / base_pr.abs().max()) <= 1e-6))
    flags = torch.tensor([int(v) for v in results.values()], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    timing = {}
    if a.time:
        for ov in (False, True):
            trainer.overlap_allreduce = ov
            trainer._graph = None
            opt.zero_grad()
            trainer.capture(batch, warmup=0)
            for _ in range(5):
                trainer.train_step(batch); trainer.optimizer_step(step_scheduler=True)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                trainer.train_step(batch); trainer.optimizer_step(step_scheduler=True)
            e1.record()
            dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            timing["overlap" if ov else "plain"] = round(float(t), 3)
    if rank == 0:
        print({"world": world, "batch_per_gpu": a.batch, "bitwise" if world == 2 else "rel1e-6": results,
               "all_ranks_ok": bool(flags.min().item()), "early_fraction": round(early / n_total, 4),
               "slices": {str(k[0]) if isinstance(k, tuple) else k: len(v) for k, v in plan.items()},
               "ms_per_step": timing}, flush=True)
    ok = bool(flags.min().item())
    from imagined_speech_translation_b200 import distributed as dp
    dp.shutdown(trainer)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

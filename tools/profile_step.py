"""Profile one preprocess + train step at BASELINE config 3 (B per GPU from argv, default 256)."""
import os
os.environ.setdefault("EEGX_BART_RANDOM_INIT", "1")   # synthetic benchmark: reference architecture, random weights (no HF cache here)
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imagined_speech_translation_b200 as pkg
from imagined_speech_translation_b200 import trainer as tr
from imagined_speech_translation_b200.model import EEGDecodingModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C, T = 64, 2048
counts = {'frontal': 16, 'temporal': 16, 'central': 16, 'parietal': 16}
fe = pkg.SpectrogramFrontEnd(C, T)
enc_counts = {k: v * fe.n_freqs for k, v in counts.items()}
torch.manual_seed(0)
model = EEGDecodingModel(n_timepoints=fe.n_frames, region_channel_counts=enc_counts).cuda()
cfg = dict(tr.CONFIG, accumulation_steps=1)
opt = tr.build_optimizer(model, cfg)
sched = tr.cosine_schedule_with_warmup(opt, 2, 1000)
t = tr.EEGTrainer(model, None, None, None, opt, sched, cfg, front_end=fe, region_channel_counts=counts)
model.train()
if os.environ.get('EEGX_SERIAL', '0') == '1':
    model.brain_encoder.parallel_regions = False
g = torch.Generator(device="cuda").manual_seed(1)
raw = 20 * torch.randn(B, C, T, device="cuda", generator=g)
labels = torch.randint(1, 51271, (B, 16), device="cuda", generator=g); labels[:, 12:] = -100
ids = torch.cat([torch.full((B, 1), 101, device="cuda"), labels[:, :-1].clamp_min(0)], 1)
batch = {'raw': raw, 'decoder_input_ids': ids, 'labels': labels}

def step():
    loss = t.train_step(batch)
    t.optimizer_step(True)
    return loss

for _ in range(3): step()
torch.cuda.synchronize()
if os.environ.get("EEGX_GRAPH", "0") == "1":
    t.capture(batch)
    for _ in range(2): step()
    torch.cuda.synchronize()
    print("captured CUDA graph")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n): loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"B={B}: {ms:.2f} ms/step  {B/ms*1e3:.0f} trials/s  loss {loss.item():.3f}  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
if os.environ.get("EEGX_TRACE", "0") == "1":
    prof.export_chrome_trace(os.path.join("gpurun_out", "step_trace.json"))
    import json
    ev = [e for e in json.load(open(os.path.join("gpurun_out", "step_trace.json")))["traceEvents"]
          if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
    def union(evs):
        tot, cur_s, cur_e = 0.0, None, None
        for e in sorted(evs, key=lambda e: e["ts"]):
            s_, e_ = e["ts"], e["ts"] + e["dur"]
            if cur_e is None or s_ > cur_e:
                if cur_e is not None: tot += cur_e - cur_s
                cur_s, cur_e = s_, e_
            else:
                cur_e = max(cur_e, e_)
        return tot + (cur_e - cur_s if cur_e is not None else 0.0)
    gemm = [e for e in ev if "gemm" in e["name"]]
    attn = [e for e in ev if "attn_" in e["name"]]
    lines = [f"kernels {len(ev)}  wall {(t1 - t0) / 1e3:.2f} ms  busy(union) {union(ev) / 1e3:.2f} ms  sum {sum(e['dur'] for e in ev) / 1e3:.2f} ms",
             f"gemm: n {len(gemm)} union {union(gemm) / 1e3:.2f} ms sum {sum(e['dur'] for e in gemm) / 1e3:.2f} ms",
             f"attn: n {len(attn)} union {union(attn) / 1e3:.2f} ms sum {sum(e['dur'] for e in attn) / 1e3:.2f} ms",
             f"non-gemm union {union([e for e in ev if 'gemm' not in e['name']]) / 1e3:.2f} ms"]
    # coarse timeline: 1 ms buckets, fraction busy and fraction with a gemm running
    nb = int((t1 - t0) / 1000) + 1
    for b in range(nb):
        lo, hi = t0 + b * 1000, t0 + (b + 1) * 1000
        def cover(evs):
            c = [dict(ts=max(e["ts"], lo), dur=min(e["ts"] + e["dur"], hi) - max(e["ts"], lo)) for e in evs
                 if e["ts"] < hi and e["ts"] + e["dur"] > lo]
            return union(c) / 1000 if c else 0.0
        n_k = sum(1 for e in ev if lo <= e["ts"] < hi)
        lines.append(f"  ms {b:3d}: busy {cover(ev):.2f} gemm {cover(gemm):.2f} kernels {n_k}")
    open(os.path.join("gpurun_out", "step_timeline.txt"), "w").write("\n".join(lines))
    print("\n".join(lines[:4]))
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=200)
os.makedirs("gpurun_out", exist_ok=True)
open(os.path.join("gpurun_out", f"step_profile_B{B}.txt"), "w").write(tab)
print(tab[:3000])

"""Per-shape GEMM time inside one single-stream eager train step (CUDA events, GPU held behind the CPU)."""
import os
os.environ.setdefault("EEGX_BART_RANDOM_INIT", "1")   # synthetic benchmark: reference architecture, random weights (no HF cache here)
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imagined_speech_translation_b200 as pkg
from imagined_speech_translation_b200 import ops, trainer as tr
from imagined_speech_translation_b200.model import EEGDecodingModel
B, C, T = 256, 64, 2048
counts = {'frontal': 16, 'temporal': 16, 'central': 16, 'parietal': 16}
fe = pkg.SpectrogramFrontEnd(C, T)
torch.manual_seed(0)
model = EEGDecodingModel(n_timepoints=fe.n_frames, region_channel_counts={k: v * fe.n_freqs for k, v in counts.items()}).cuda().train()
model.brain_encoder.parallel_regions = False
from imagined_speech_translation_b200 import fused
fused.DEFER_WGRAD = False      # one stream: an event pair brackets exactly one GEMM
cfg = dict(tr.CONFIG, accumulation_steps=1)
opt = tr.build_optimizer(model, cfg)
t = tr.EEGTrainer(model, None, None, None, opt, tr.cosine_schedule_with_warmup(opt, 2, 1000), cfg, front_end=fe, region_channel_counts=counts)
g = torch.Generator(device="cuda").manual_seed(1)
labels = torch.randint(1, 51271, (B, 16), device="cuda", generator=g); labels[:, 12:] = -100
ids = torch.cat([torch.full((B, 1), 101, device="cuda"), labels[:, :-1].clamp_min(0)], 1)
batch = {'raw': 20 * torch.randn(B, C, T, device="cuda", generator=g), 'decoder_input_ids': ids, 'labels': labels}
for _ in range(3):
    t.train_step(batch); t.optimizer_step(True)
torch.cuda.synchronize()
torch.cuda._sleep(int(0.2 * 1.9e9))
ops.GEMM_TIMING = []
t.train_step(batch); t.optimizer_step(True)
torch.cuda.synchronize()
rec, ops.GEMM_TIMING = ops.GEMM_TIMING, None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e0, e1, fl, shp in rec:
    a = agg[shp]; a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl
tot = sum(a[1] for a in agg.values())
print(f"{len(rec)} GEMM launches, {tot:.2f} ms (raw event time), {sum(a[2] for a in agg.values()) / tot / 1e9:.0f} TF/s")
print(f"{'batch,M,N,K,aT,bT,out bytes':>34} {'n':>4} {'ms':>7} {'us/call':>8} {'TF/s':>7}")
for shp, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{str(shp):>34} {n:4d} {ms:7.3f} {1e3 * ms / n:8.1f} {fl / ms / 1e9:7.0f}")

"""Profile one preprocess + train step at BASELINE config 3 (B per GPU from argv, default 256)."""
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
    t._optimizer_step(True)
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
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=200)
os.makedirs("gpurun_out", exist_ok=True)
open(os.path.join("gpurun_out", f"step_profile_B{B}.txt"), "w").write(tab)
print(tab[:3000])

"""GPU tests of the train-step pieces: fused clip + AdamW against torch's own, and the
EEGTrainer.train_epoch step semantics of the reference (trainer.py:69-151)."""
import math

import pytest
import torch

import imagined_speech_translation_b200 as pkg
from imagined_speech_translation_b200 import trainer as tr
from imagined_speech_translation_b200.model import EEGDecodingModel

pytestmark = pytest.mark.gpu


def _toy(seed):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.GELU(), torch.nn.Linear(53, 11),
                            torch.nn.Linear(11, 11))      # last layer unused: its grad stays None
    return m.cuda()


def test_fused_clip_adamw_matches_torch():
    a, b = _toy(1), _toy(1)
    groups = lambda m: [{'params': list(m[0].parameters()), 'lr': 3e-3},
                        {'params': list(m[2].parameters()) + list(m[3].parameters()), 'lr': 1e-3}]
    ours = pkg.FlatAdamW(groups(a), eps=1e-8, betas=(0.9, 0.999), weight_decay=0.01)
    ref = torch.optim.AdamW(groups(b), eps=1e-8, betas=(0.9, 0.999), weight_decay=0.01)
    unused_before = a[3].weight.detach().clone()
    g = torch.Generator(device="cuda").manual_seed(3)
    for step in range(6):
        x = torch.randn(16, 37, device="cuda", generator=g) * (10.0 if step % 2 else 0.1)   # clip on/off
        for m in (a, b):
            m[2](m[1](m[0](x))).pow(2).sum().backward()
        ref_norm = torch.nn.utils.clip_grad_norm_(b.parameters(), 1.0)
        ref.step(); ref.zero_grad()
        ours.step(max_grad_norm=1.0)
        assert ours.grad_norm().item() == pytest.approx(ref_norm.item(), rel=1e-5)
        ours.zero_grad()
        for pa, pb in zip(a.parameters(), b.parameters()):
            torch.testing.assert_close(pa, pb, rtol=2e-6, atol=2e-7)
    assert torch.equal(a[3].weight, unused_before)            # grad None => untouched (no decay)
    assert a[0].weight.grad is not None and float(a[0].weight.grad.abs().sum()) == 0.0


def _batches(n, B, counts, T, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        labels = torch.randint(1, 51271, (B, 16), generator=g)
        labels[:, 12:] = -100
        ids = torch.cat([torch.full((B, 1), 101), labels[:, :-1].clamp_min(0)], dim=1)
        out.append({'eeg': [torch.randn(B, counts[k], T, generator=g) for k in
                            ('frontal', 'temporal', 'central', 'parietal')],
                    'decoder_input_ids': ids, 'labels': labels})
    return out


def test_train_epoch_semantics():
    counts = {'frontal': 16, 'temporal': 16, 'central': 16, 'parietal': 16}
    torch.manual_seed(0)
    model = EEGDecodingModel(n_timepoints=33, region_channel_counts=counts).cuda()
    cfg = dict(tr.CONFIG, accumulation_steps=2, warmup_steps=2)
    opt = tr.build_optimizer(model, cfg)
    sched = tr.cosine_schedule_with_warmup(opt, cfg['warmup_steps'], 10)
    loader = _batches(5, 2, counts, 33, seed=1)               # 5 micro-batches, accumulation 2
    t = tr.EEGTrainer(model, None, loader, None, opt, sched, cfg)
    w0 = model.brain_encoder.feature_enhancer[0].weight.detach().clone()
    enc0 = model.bart_decoder.bart.model.encoder.layers[0].fc1.weight.detach().clone()
    loss = t.train_epoch(0)
    # random-init decoder: loss ~ ln(vocab)
    assert abs(loss - math.log(51271)) < 1.0
    # 2 full accumulation windows stepped the scheduler; the flush (5th batch) did not
    assert t.global_step == 2 and sched.last_epoch == 2
    assert opt._step == 3
    # lr of the first optimizer step is 0 (warm-up lambda(0) = 0) but later steps move the weights
    assert not torch.equal(model.brain_encoder.feature_enhancer[0].weight, w0)
    # the BART encoder never runs: its parameters get no gradient and are never touched
    assert torch.equal(model.bart_decoder.bart.model.encoder.layers[0].fc1.weight, enc0)
    assert model.bart_decoder.bart.model.encoder.layers[0].fc1.weight.grad is None
    n_grad = sum(p.numel() for p in model.parameters() if p.grad is not None)
    n_all = sum(p.numel() for p in model.parameters())
    assert n_all - n_grad == 43_316_736


def test_first_step_has_zero_lr():
    counts = {'frontal': 16, 'temporal': 16, 'central': 16, 'parietal': 16}
    torch.manual_seed(0)
    model = EEGDecodingModel(n_timepoints=33, region_channel_counts=counts).cuda()
    cfg = dict(tr.CONFIG, accumulation_steps=1, warmup_steps=500)
    opt = tr.build_optimizer(model, cfg)
    sched = tr.cosine_schedule_with_warmup(opt, 500, 1000)
    t = tr.EEGTrainer(model, None, _batches(1, 2, counts, 33, seed=2), None, opt, sched, cfg)
    w0 = model.brain_encoder.region_encoders['frontal'].conv2.weight.detach().clone()
    t.train_epoch(0)
    assert torch.equal(model.brain_encoder.region_encoders['frontal'].conv2.weight, w0)

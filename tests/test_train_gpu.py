"""GPU tests of the train-step pieces: fused clip + AdamW against torch's own, and the
EEGTrainer.train_epoch step semantics of the reference (trainer.py:69-151)."""
import math
import os

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


def test_flat_adamw_state_dict_interchanges_with_torch_adamw():
    """a checkpoint must carry the moments.  3 steps with ours -> state_dict -> torch.optim.AdamW
    continues for 3 steps; the mirror run stays on torch.optim.AdamW throughout; then the other direction
    (torch state -> FlatAdamW, before AND after its flat buffers exist)."""
    groups = lambda m: [{'params': list(m[0].parameters()), 'lr': 3e-3},
                        {'params': list(m[2].parameters()) + list(m[3].parameters()), 'lr': 1e-3}]
    kw = dict(eps=1e-8, betas=(0.9, 0.999), weight_decay=0.01)

    def run(m, opt, steps, seed, ours):
        g = torch.Generator(device="cuda").manual_seed(seed)
        for _ in range(steps):
            x = torch.randn(16, 37, device="cuda", generator=g)
            m[2](m[1](m[0](x))).pow(2).sum().backward()
            opt.step()
            opt.zero_grad() if ours else opt.zero_grad(set_to_none=False)

    a, b = _toy(5), _toy(5)
    oa, ob = pkg.FlatAdamW(groups(a), **kw), torch.optim.AdamW(groups(b), **kw)
    run(a, oa, 3, 11, True); run(b, ob, 3, 11, False)
    sd = oa.state_dict()
    assert set(sd['state']) == set(ob.state_dict()['state'])              # the unused layer has no entry in either
    for k, st in ob.state_dict()['state'].items():
        torch.testing.assert_close(sd['state'][k]['exp_avg'], st['exp_avg'], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(sd['state'][k]['exp_avg_sq'], st['exp_avg_sq'], rtol=1e-4, atol=1e-7)
        assert float(sd['state'][k]['step']) == float(st['step']) == 3.0
    # ours -> torch
    c = _toy(5)
    c.load_state_dict(a.state_dict())
    oc = torch.optim.AdamW(groups(c), **kw)
    oc.load_state_dict(sd)
    run(c, oc, 3, 12, False); run(b, ob, 3, 12, False)
    for pc, pb in zip(c.parameters(), b.parameters()):
        torch.testing.assert_close(pc, pb, rtol=5e-6, atol=5e-7)
    # torch -> ours, loaded before the flat buffers exist (applied at the first step) ...
    d = _toy(5)
    d.load_state_dict(b.state_dict())
    od = pkg.FlatAdamW(groups(d), **kw)
    od.load_state_dict(ob.state_dict())
    assert od.state_dict()['state'].keys() == ob.state_dict()['state'].keys()
    # ... and after
    e = _toy(5)
    e.load_state_dict(b.state_dict())
    oe = pkg.FlatAdamW(groups(e), **kw)
    run(e, oe, 1, 99, True)
    with torch.no_grad():
        for pe, pb in zip(e.parameters(), b.parameters()):
            pe.copy_(pb)                                                     # in place: the views stay bound
    oe.load_state_dict(ob.state_dict())
    run(d, od, 2, 13, True); run(e, oe, 2, 13, True); run(b, ob, 2, 13, False)
    for pd, pe, pb in zip(d.parameters(), e.parameters(), b.parameters()):
        torch.testing.assert_close(pd, pb, rtol=5e-6, atol=5e-7)
        torch.testing.assert_close(pe, pb, rtol=5e-6, atol=5e-7)


def test_flat_adamw_fails_loudly_when_a_parameter_escapes_the_flat_buffers():
    m = _toy(2)
    opt = pkg.FlatAdamW([{'params': list(m.parameters()), 'lr': 1e-3}])
    x = torch.randn(4, 37, device="cuda")
    m[2](m[1](m[0](x))).sum().backward()
    opt.step(); opt.zero_grad()
    m(x).sum().backward()                                   # the last layer gets its FIRST gradient only now
    with pytest.raises(pkg.EegxError, match="first gradient"):
        opt.step()
    m2 = _toy(2)
    opt2 = pkg.FlatAdamW([{'params': list(m2.parameters()), 'lr': 1e-3}])
    m2(x).sum().backward()
    opt2.step(); opt2.zero_grad()
    m2[0].weight.data = m2[0].weight.data.clone()           # rebinding .data detaches the parameter from the buffers
    m2(x).sum().backward()
    with pytest.raises(pkg.EegxError, match="no longer points"):
        opt2.step()


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


def _zero_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
        if isinstance(m, torch.nn.TransformerEncoderLayer):
            m.dropout.p = m.dropout1.p = m.dropout2.p = 0.0
    dec = model.bart_decoder.bart.model.decoder
    dec.dropout = 0.0
    for layer in dec.layers:
        layer.dropout = 0.0


def test_high_density_long_window_step():
    """BASELINE configs[3] shape at a small batch: 128 ch x 4096 samples, STFT n_fft=1024 hop=256
    (long-window DSP kernel) -> 4 regions of 32 x 513 feature channels x 17 frames -> one train step."""
    C, T, B = 128, 4096, 2
    counts = {'frontal': 32, 'temporal': 32, 'central': 32, 'parietal': 32}
    fe = pkg.SpectrogramFrontEnd(C, T, pkg.DSP_CONFIG_LONG)
    assert (fe.n_freqs, fe.n_frames) == (513, 17) and fe.kernel_name == "long"
    torch.manual_seed(0)
    model = EEGDecodingModel(n_timepoints=fe.n_frames,
                             region_channel_counts={k: v * fe.n_freqs for k, v in counts.items()}).cuda().train()
    cfg = dict(tr.CONFIG, accumulation_steps=1, warmup_steps=1)
    opt = tr.build_optimizer(model, cfg)
    sched = tr.cosine_schedule_with_warmup(opt, 1, 10)
    t = tr.EEGTrainer(model, None, None, None, opt, sched, cfg, front_end=fe, region_channel_counts=counts)
    g = torch.Generator().manual_seed(5)
    labels = torch.randint(1, 51271, (B, 16), generator=g)
    labels[:, 10:] = -100
    ids = torch.cat([torch.full((B, 1), 101), labels[:, :-1].clamp_min(0)], dim=1)
    batch = {'raw': 20.0 * torch.randn(B, C, T, generator=g), 'decoder_input_ids': ids, 'labels': labels}
    for _ in range(2):
        loss = t.train_step(batch)
        t._optimizer_step(step_scheduler=True)
    assert torch.isfinite(loss) and abs(loss.item() - math.log(51271)) < 1.0
    assert float(opt.grad_norm()) > 0.0


def test_cuda_graph_replay_matches_eager():
    """Same weights, same batches, dropout off: a model stepped eagerly and a model stepped through the
    captured graph follow the same loss trajectory while the weights move (so the graph must be reading the
    live weights, not copies frozen at capture time)."""
    counts = {'frontal': 16, 'temporal': 16, 'central': 16, 'parietal': 16}

    class _T(tr.EEGTrainer):
        def _regions(self, batch):
            return [batch['eeg0'], batch['eeg1'], batch['eeg2'], batch['eeg3']]

    def make():
        torch.manual_seed(0)
        model = EEGDecodingModel(n_timepoints=33, region_channel_counts=counts).cuda().train()
        _zero_dropout(model)
        cfg = dict(tr.CONFIG, accumulation_steps=1)
        opt = tr.build_optimizer(model, cfg)
        sched = tr.cosine_schedule_with_warmup(opt, 1, 100)
        return _T(model, None, None, None, opt, sched, cfg), opt

    flat = [{'eeg0': b['eeg'][0].cuda(), 'eeg1': b['eeg'][1].cuda(), 'eeg2': b['eeg'][2].cuda(), 'eeg3': b['eeg'][3].cuda(),
             'decoder_input_ids': b['decoder_input_ids'].cuda(), 'labels': b['labels'].cuda()}
            for b in _batches(1, 4, counts, 33, seed=9)] * 6     # one batch, repeated: the loss falls quickly
    t_e, opt_e = make()
    eager = []
    for b in flat:
        eager.append(t_e.train_step(b).item())
        t_e._optimizer_step(step_scheduler=True)
    t_g, opt_g = make()
    got = []
    for i, b in enumerate(flat):
        if i == 2:
            t_g.capture(flat[0])
        got.append(t_g.train_step(b).item())
        t_g._optimizer_step(step_scheduler=True)
    assert eager[0] - eager[-1] > 0.3                     # the weights really moved
    for a, e in zip(got, eager):
        assert abs(a - e) <= 5e-3 * abs(e), (got, eager)
    assert abs(float(opt_g.grad_norm()) - float(opt_e.grad_norm())) <= 3e-2 * float(opt_e.grad_norm())


def test_step_is_bit_reproducible():
    """Two identical runs (same seed, same batches, dropout ON) give bit-identical losses and gradients:
    every reduction in libeegx has a fixed summation order and the dropout masks are counter based."""
    from imagined_speech_translation_b200 import fused
    counts = {'frontal': 16, 'temporal': 16, 'central': 16, 'parietal': 16}
    batches = _batches(3, 4, counts, 33, seed=21)

    def run():
        torch.manual_seed(0)
        fused.set_seed(1234)
        model = EEGDecodingModel(n_timepoints=33, region_channel_counts=counts).cuda().train()
        cfg = dict(tr.CONFIG, accumulation_steps=1, warmup_steps=1)
        opt = tr.build_optimizer(model, cfg)
        sched = tr.cosine_schedule_with_warmup(opt, 1, 100)
        t = tr.EEGTrainer(model, None, None, None, opt, sched, cfg)
        losses = []
        for b in batches:
            losses.append(t.train_step(b).item())
            t._optimizer_step(step_scheduler=True)
        t.train_step(batches[0])
        names = ["brain_encoder.region_encoders.frontal.conv1.weight", "brain_encoder.region_encoders.parietal.bn3.weight",
                 "brain_encoder.region_encoders.central.attn_layers.1.attn.in_proj_weight",
                 "brain_encoder.region_encoders.temporal.attn_layers.0.ffn.gate.bias",
                 "brain_encoder.feature_enhancer.0.weight", "bart_decoder.bart.model.decoder.layers.3.fc1.weight",
                 "bart_decoder.bart.model.shared.weight", "bart_decoder.eeg_to_bart.1.weight"]
        params = dict(model.named_parameters())
        return losses, {n: params[n].grad.detach().clone() for n in names}

    l1, g1 = run()
    l2, g2 = run()
    assert l1 == l2
    for n in g1:
        assert torch.equal(g1[n], g2[n]), n


@pytest.mark.gpu
def test_overlapped_allreduce_is_bitwise_equal_to_single_allreduce_two_gpus():
    """Two ranks, identical replicas, per-rank batches: the gradients after the slice-wise all-reduce issued from
    inside backward (eager and CUDA graph) equal the single post-backward all-reduce bit for bit
    (tools/check_overlap.py).  Needs two GPUs; the single-GPU box of the round-end run skips it."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tools", "check_overlap.py"), "--batch", "16"],
                         capture_output=True, text=True, timeout=400)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "'all_ranks_ok': True" in res.stdout


@pytest.mark.gpu
def test_evaluate_reports_loss_and_generated_text():
    """EEGTrainer.evaluate (trainer.py:153-212): eval-mode loss + beam-3 generation per validation batch, decoded
    by the tokenizer; generation goes through generation.generate and equals model.generate on the same batch."""
    from transformers import BertTokenizer
    fix = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset")
    tok = BertTokenizer(os.path.join(fix, "vocab.txt"), bos_token="[CLS]", eos_token="[SEP]")
    torch.manual_seed(0)
    counts = {"frontal": 16, "temporal": 16, "central": 16, "parietal": 16}
    model = EEGDecodingModel(n_timepoints=33, region_channel_counts=counts).cuda()
    tr.initialize_custom_weights(model)
    g = torch.Generator().manual_seed(3)
    batches = []
    for _ in range(2):
        labels = torch.randint(1, len(tok), (4, 16), generator=g)
        labels[:, 12:] = -100
        ids = torch.cat([torch.full((4, 1), 101), labels[:, :-1].clamp(min=0)], dim=1)
        batches.append({"eeg": [torch.randn(4, 16, 33, generator=g) for _ in range(4)],
                        "decoder_input_ids": ids, "labels": labels})
    cfg = dict(tr.CONFIG, accumulation_steps=1)
    trainer = tr.EEGTrainer(model, tok, None, batches, None, None, cfg)

    class Count:
        def compute_all_metrics(self, preds, targets):
            return {"n": len(preds), "n_targets": len(targets)}

    model.train()
    metrics = trainer.evaluate(Count())
    assert model.training                                           # mode restored
    assert metrics["n"] == metrics["n_targets"] == 8 and math.isfinite(metrics["val_loss"])
    assert all(isinstance(t, str) for t in trainer.last_predictions)
    model.eval()
    with torch.no_grad():
        direct = model.generate(eeg_data=[r.cuda() for r in batches[0]["eeg"]], **cfg["generation"]["eval"]).cpu()
    want = [tok.decode(direct[i], skip_special_tokens=True, clean_up_tokenization_spaces=True).strip() for i in range(4)]
    assert trainer.last_predictions[:4] == want


def test_raw_batches_take_the_reference_normalisation_path():
    """Row a2 reachable from the step: a batch carrying 'raw' trials and no spectrogram front-end goes through the
    RegionNormalizer (dataset.py:172-225 on the GPU) and yields the same loss as the pre-normalised 'eeg' list."""
    from imagined_speech_translation_b200.preprocess import RegionNormalizer, REGION_ORDER
    g = torch.Generator().manual_seed(3)
    C_in, T, B = 24, 37, 2
    region_indices = {n: list(range(1 + 5 * i, 1 + 5 * i + 4)) for i, n in enumerate(REGION_ORDER)}
    centers = {n: torch.randn(4, generator=g).numpy() for n in REGION_ORDER[:3]}        # 4th region: z-score fallback
    scales = {n: (1.0 + torch.rand(4, generator=g)).numpy() for n in REGION_ORDER[:3]}
    norm = RegionNormalizer(region_indices, centers, scales)
    counts = {n: 4 for n in REGION_ORDER}
    torch.manual_seed(0)
    model = EEGDecodingModel(n_timepoints=T, region_channel_counts=counts).cuda().eval()
    cfg = dict(tr.CONFIG, accumulation_steps=1)
    opt = tr.build_optimizer(model, cfg)
    t = tr.EEGTrainer(model, None, None, None, opt, tr.cosine_schedule_with_warmup(opt, 1, 10), cfg, normalizer=norm)
    raw = 30.0 * torch.randn(B, C_in, T, generator=g)
    raw[0, 2, 5] = float('nan')
    labels = torch.randint(1, 51271, (B, 8), generator=g)
    ids = torch.cat([torch.full((B, 1), 101), labels[:, :-1]], 1)
    batch = {'raw': raw, 'decoder_input_ids': ids, 'labels': labels}
    regions = t._regions(batch)
    assert [tuple(r.shape) for r in regions] == [(B, 4, T)] * 4 and all(torch.isfinite(r).all() for r in regions)
    with torch.no_grad():
        a = t.forward_pass(regions, ids.cuda(), labels.cuda()).loss
        b = t.forward_pass(t._regions({'eeg': [r.cpu() for r in regions]}), ids.cuda(), labels.cuda()).loss
    assert torch.equal(a, b)
    # no front-end, no normalizer, no dataset to take one from: a loud error, not a silent pass-through
    t2 = tr.EEGTrainer(model, None, None, None, opt, tr.cosine_schedule_with_warmup(opt, 1, 10), cfg)
    with pytest.raises(ValueError):
        t2._regions(batch)

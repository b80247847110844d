"""Training loop step on the B200 path.

Mirror of the working part of the reference trainer -- ``EEGTrainer.forward_pass`` and
``EEGTrainer.train_epoch`` (``main_model/src/training/trainer.py:40-151``) -- and of its wiring
(``config/training_config.py:5-77``, ``scripts/train.py:199-241``): same constructor, same
step semantics (loss / accumulation_steps, clip to grad_clip_norm, optimizer step, zero_grad,
scheduler step, global_step; the end-of-epoch flush steps the optimizer but NOT the scheduler),
three learning-rate groups routed by parameter-name substring.  Differences, all deliberate:
errors are raised instead of swallowed, the loss is accumulated on the device and read back once
per epoch instead of ``.item()`` every micro-batch, clipping + AdamW are one fused step, and under
data parallelism the flat gradients are all-reduced once per optimizer step.
"""
from __future__ import annotations

import math
import os

import torch

from . import distributed as dp
from . import fused
from .optim import FlatAdamW

CONFIG = {
    # values of the reference config (training_config.py:5-52) that define the step
    'hidden_dim': 768, 'n_timepoints': 1651, 'max_length': 16,
    'epochs': 100, 'batch_size': 4, 'accumulation_steps': 8, 'grad_clip_norm': 1.0,
    'brain_encoder_lr': 3e-4, 'bart_decoder_lr': 3e-5, 'projection_lr': 1e-4,
    'warmup_steps': 500, 'weight_decay': 0.01, 'log_interval': 20, 'seed': 42,
    'generation': {'eval': {'max_length': 16, 'min_length': 4, 'num_beams': 3, 'early_stopping': True}},
    # DSP front-end keys (SURVEY.md section 5 "Config / flags")
    'fs': 256.0, 'band': (8.0, 30.0), 'numtaps': 65, 'n_fft': 256, 'hop': 64, 'log_eps': 1.0,
}


def get_optimizer_groups(model, config=CONFIG):
    """Three groups by name substring, as training_config.py:55-77."""
    enc, proj, bart = [], [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if 'brain_encoder' in name:
            enc.append(p)
        elif 'eeg_to_bart' in name:
            proj.append(p)
        elif 'bart' in name:
            bart.append(p)
    return [{'params': enc, 'lr': config['brain_encoder_lr']},
            {'params': proj, 'lr': config['projection_lr']},
            {'params': bart, 'lr': config['bart_decoder_lr']}]


def build_optimizer(model, config=CONFIG):
    """scripts/train.py:210-215: AdamW(eps 1e-8, betas .9/.999, weight_decay from the config)."""
    enc = getattr(model, 'brain_encoder', None)
    stacks = enc.parameter_stacks() if hasattr(enc, 'parameter_stacks') else None     # lock-step region path (grouped.py)
    return FlatAdamW(get_optimizer_groups(model, config), eps=1e-8, betas=(0.9, 0.999),
                     weight_decay=config['weight_decay'], stacks=stacks)


def cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, num_cycles=0.5):
    """Same multiplier as transformers.get_cosine_schedule_with_warmup (scripts/train.py:227-231);
    note lambda(0) = 0: the first optimizer step runs with lr 0."""
    def lr_lambda(step):
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda)


def initialize_custom_weights(model):
    """Initialisation of the non-BART parameters by name substring, as
    ``scripts/train.py:108-126``: '*weight*' with 'norm' -> 1, with 'embedding' -> N(0, 0.02^2),
    other weights with >= 2 dims -> xavier_uniform(gain 0.02); '*bias*' -> 0.  Tokens, positional
    embeddings, region_importance and 1-D BatchNorm / in-Sequential LayerNorm gains keep their
    constructor values (no 'weight'/'bias' in the name, or 1-D without 'norm')."""
    for name, p in model.named_parameters():
        if 'bart' in name.lower():
            continue
        if 'weight' in name:
            if 'norm' in name:
                torch.nn.init.ones_(p)
            elif 'embedding' in name:
                torch.nn.init.normal_(p, std=0.02)
            elif p.dim() >= 2:
                torch.nn.init.xavier_uniform_(p, gain=0.02)
        elif 'bias' in name:
            torch.nn.init.zeros_(p)
    return model


class EEGTrainer:
    def __init__(self, model, tokenizer, train_loader, val_loader, optimizer, scheduler, config,
                 front_end=None, region_channel_counts=None, process_group=None, normalizer=None):
        self.model = model
        self.tokenizer = tokenizer
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.config = config
        self.device = next(model.parameters()).device
        self.front_end = front_end                      # SpectrogramFrontEnd: batch['raw'] -> regions
        # RegionNormalizer (the reference's own preprocessing, dataset.py:172-225: gather the four regions,
        # nan_to_num, RobustScaler / z-score fallback) for batch['raw'] when no spectrogram front-end is given;
        # taken from the train loader's dataset when that is a data.EEGDataset
        self.normalizer = normalizer
        self.region_channel_counts = region_channel_counts
        self.process_group = process_group
        self.world_size = torch.distributed.get_world_size(process_group) \
            if torch.distributed.is_initialized() else 1
        # data-parallel: all-reduce slices of the flat gradient buffer as backward finishes them (EEGX_OVERLAP_ALLREDUCE=0
        # falls back to one all-reduce after backward)
        self.overlap_allreduce = os.environ.get('EEGX_OVERLAP_ALLREDUCE', '1') != '0'
        self._plan = None
        self._plan_lock_step = None
        self._works = []
        self._fired = set()
        self._grads_reduced = False
        self._graph_reduces = False
        self.best_bleu4 = 0.0
        self.patience_counter = 0
        self.global_step = 0
        self.epoch = 0

    # ------------------------------------------------------------------
    def _regions(self, batch):
        if 'raw' in batch and self.front_end is not None:
            raw = batch['raw'].to(self.device, non_blocking=True)
            return self.front_end.split_regions(self.front_end(raw), self.region_channel_counts)
        if 'raw' in batch:
            # reference-actual path (rows a1/a2): the arithmetic the reference does per item in __getitem__ runs
            # here, once per batch, on the GPU
            norm = self.normalizer
            if norm is None:
                ds = getattr(self.train_loader, 'dataset', None)
                if ds is None or not hasattr(ds, 'normalizer'):
                    raise ValueError("batch carries 'raw' trials but the trainer has neither a front_end nor a "
                                     "normalizer (pass normalizer=RegionNormalizer(...) or a data.EEGDataset loader)")
                norm = self.normalizer = ds.normalizer()
            return norm(batch['raw'].to(self.device, non_blocking=True).float())
        return [r.to(self.device, non_blocking=True) for r in batch['eeg']]

    def forward_pass(self, eeg, decoder_input_ids, labels):
        return self.model(eeg, decoder_input_ids=decoder_input_ids, labels=labels)

    # ------------------------------------------------------------------ all-reduce overlapped with backward
    def _overlap_plan(self):
        """Which slices of the flat gradient buffer are final at which point of backward (SURVEY.md 8(e):
        "bucketed and overlapped with backward").  Backward runs decoder -> fusion stage -> the four region
        encoders (attention stack -- split in two --, then CNN); a `fused.grad_boundary` sits at each of those transitions and
        its backward starts the all-reduce of the slice that has just become final.  Only what is left
        (region CNN stacks, tokens, positions: ~8 % of the bytes) is reduced after backward."""
        opt, model = self.optimizer, self.model
        enc, dec = model.brain_encoder, model.bart_decoder
        from . import grouped
        lock_step = getattr(enc, 'lock_step_regions', True) and grouped.available(list(enc.region_encoders.values()))
        if self._plan is not None and self._plan_lock_step == lock_step:
            return self._plan
        self._plan_lock_step = lock_step
        plan = {('decoder', id(dec)): opt.grad_runs(list(dec.parameters())),
                ('fusion', id(enc)): opt.grad_runs([p for n, p in enc.named_parameters()
                                                    if not n.startswith('region_encoders.')])}
        if lock_step:
            # the regions run as one stacked pass: one boundary pair for all of them, and the stacked layout of the
            # flat buffers makes the union of the regions' slices contiguous again
            mods = list(enc.region_encoders.values())
            mid = len(mods[0].attn_layers) // 2
            upper = [p for m in mods for sub in (m.attn_layers[mid:], m.multi_scale_proj, m.projection, m.diversity_head)
                     for p in sub.parameters()]
            lower = [p for m in mods for sub in (m.attn_layers[:mid], m.cross_scale_attn) for p in sub.parameters()]
            if mid > 0:
                plan[('region_mid', 'group')] = opt.grad_runs(upper)
                plan[('region', 'group')] = opt.grad_runs(lower)
            else:
                plan[('region', 'group')] = opt.grad_runs(upper + lower)
        for m in ([] if lock_step else enc.region_encoders.values()):
            if getattr(m, 'cnn_only', False):
                continue
            # two points per region: halfway through the attention stack (upper layers + heads) and at its input
            # (lower layers + cross_scale_attn, which every layer but the first uses)
            mid = len(m.attn_layers) // 2
            upper = [p for sub in (m.attn_layers[mid:], m.multi_scale_proj, m.projection, m.diversity_head)
                     for p in sub.parameters()]
            lower = [p for sub in (m.attn_layers[:mid], m.cross_scale_attn) for p in sub.parameters()]
            if mid > 0:
                plan[('region_mid', id(m))] = opt.grad_runs(upper)
                plan[('region', id(m))] = opt.grad_runs(lower)
            else:
                plan[('region', id(m))] = opt.grad_runs(upper + lower)
        covered = sorted(r for runs in plan.values() for r in runs)
        rest, pos, total = [], 0, opt._all_grads.numel()
        for lo, hi in covered:
            if lo < pos:
                raise RuntimeError("overlap plan: gradient slices overlap")
            if lo > pos:
                rest.append((pos, lo))
            pos = hi
        if pos < total:
            rest.append((pos, total))
        plan['rest'] = rest
        self._plan = plan
        return plan

    def _overlap_active(self):
        return (self.world_size > 1 and self.overlap_allreduce and self.config['accumulation_steps'] == 1
                and isinstance(self.optimizer, FlatAdamW) and self.optimizer._flat is not None)

    def _reduce_runs(self, runs):
        g = self.optimizer._all_grads
        for lo, hi in runs:
            self._works.append(torch.distributed.all_reduce(g[lo:hi], group=self.process_group, async_op=True))

    def _on_boundary(self, key):
        runs = self._overlap_plan().get(key)
        if runs and key not in self._fired:
            self._fired.add(key)
            fused.join_side()          # weight gradients deferred to the side stream belong to the slice as well
            self._reduce_runs(runs)

    def optimizer_step(self, step_scheduler: bool = True):
        """Public form of the step `train_epoch` takes every `accumulation_steps` micro-batches (reference
        trainer.py:101-113): gradient all-reduce if still pending, clip, AdamW, zero_grad, scheduler."""
        return self._optimizer_step(step_scheduler)

    def _optimizer_step(self, step_scheduler: bool):
        clip = self.config.get('grad_clip_norm', 1.0)
        if isinstance(self.optimizer, FlatAdamW):
            if self.world_size > 1:
                if not self._grads_reduced:
                    dp.allreduce_sum_(self.optimizer.flat_grads(), self.process_group)
                self._grads_reduced = False
                self.optimizer.grad_scale = 1.0 / self.world_size
            self.optimizer.step(max_grad_norm=clip)
        else:
            if self.world_size > 1:
                dp.allreduce_gradients(list(self.model.parameters()), self.process_group)
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), clip)
            self.optimizer.step()
        self.optimizer.zero_grad()
        if step_scheduler:
            self.scheduler.step()
            self.global_step += 1

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, example_batch, warmup: int = 2):
        """Capture preprocess + forward + backward of one micro-batch (fixed shapes) into a CUDA
        graph; later train_step() calls with same-shaped batches copy into the static inputs and
        replay.  The optimizer step stays outside (its learning rates change every step).
        Needs the flat gradient buffers to exist, i.e. at least one optimizer step before."""
        if not isinstance(self.optimizer, FlatAdamW) or self.optimizer._flat is None:
            raise RuntimeError("run one eager train_step + optimizer step before capture()")
        self._static = {k: v.to(self.device).clone() for k, v in example_batch.items() if torch.is_tensor(v)}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step(self._static)
            self.optimizer.zero_grad()
            self._grads_reduced = False
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import nn_ops
        nn_ops.clear_pack_cache()      # derived weight packs must be (re)built INSIDE the graph, from live weights
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._eager_step(self._static)
        self._graph_reduces = self._grads_reduced        # the captured step contains its own all-reduces
        self._grads_reduced = False
        self.optimizer.zero_grad()
        return self

    def release_graph(self):
        """Drop the captured step.  Must run before ``torch.distributed.destroy_process_group()`` when the graph
        holds NCCL all-reduces: the communicator cannot be torn down while a graph still references it."""
        g = getattr(self, "_graph", None)
        self._graph = None
        self._graph_reduces = False
        if g is not None:
            torch.cuda.synchronize()
            g.reset()
            del g
            torch.cuda.synchronize()

    def _eager_step(self, batch):
        fused.begin_step()
        fused.advance_rng(self.device)           # in-graph increment: every replay draws new dropout masks
        overlap = self._overlap_active()
        if not overlap:
            out = self._forward_loss(batch)
            (out.loss / self.config['accumulation_steps']).backward()
            return out.loss.detach()
        self._works = []
        self._fired = set()
        fused.set_grad_boundary_callback(self._on_boundary)          # before forward: the boundaries are tape nodes
        try:
            out = self._forward_loss(batch)
            (out.loss / self.config['accumulation_steps']).backward()
        finally:
            fused.set_grad_boundary_callback(None)
        # slices whose boundary did not fire in this backward pass (the plan assumed the lock-step region path but the
        # batch took the per-region one, or the other way round; a sequence length the fused attention does not
        # serve; ...) are reduced now, with the rest: every gradient is reduced exactly once whatever path ran
        for key, runs in self._overlap_plan().items():
            if key != 'rest' and key not in self._fired:
                self._reduce_runs(runs)
        self._reduce_runs(self._overlap_plan()['rest'])
        for w in self._works:
            w.wait()                             # the compute stream waits for the NCCL stream (no host sync)
        self._works = []
        self._grads_reduced = True
        return out.loss.detach()

    def _forward_loss(self, batch):
        eeg = self._regions(batch)
        ids = batch['decoder_input_ids'].to(self.device, non_blocking=True)
        labels = batch['labels'].to(self.device, non_blocking=True)
        out = self.forward_pass(eeg, ids, labels)
        if out.loss is None:
            raise RuntimeError("model returned no loss")
        return out

    def train_step(self, batch):
        """One micro-batch: preprocess (if raw) + forward + backward of loss / accumulation_steps.
        Returns the un-scaled loss as a device scalar (no host sync)."""
        if getattr(self, "_graph", None) is not None:
            for k, dst in self._static.items():
                dst.copy_(batch[k], non_blocking=True)
            self._graph.replay()
            self._grads_reduced = self._graph_reduces
            return self._static_loss
        return self._eager_step(batch)

    @torch.no_grad()
    def evaluate(self, evaluator=None):
        """``EEGTrainer.evaluate`` (trainer.py:153-212): validation loss and beam-search generation per batch,
        decoded with the tokenizer.  Returns ``{'val_loss': ...}`` plus whatever ``evaluator.compute_all_metrics(
        predictions, targets)`` adds (the reference's BLEU / ROUGE evaluator is host-side text processing and out of
        scope; any object with that method can be passed).  The texts are kept in ``last_predictions`` /
        ``last_targets``.  Unlike the reference, a failing batch raises instead of being skipped."""
        was_training = self.model.training
        self.model.eval()
        gen_config = dict(self.config['generation']['eval'])
        loss_sum = torch.zeros((), device=self.device, dtype=torch.float64)
        n_samples = 0
        predictions, targets = [], []
        try:
            for batch in self.val_loader:
                eeg = self._regions(batch)
                ids = batch['decoder_input_ids'].to(self.device, non_blocking=True)
                labels = batch['labels'].to(self.device, non_blocking=True)
                out = self.forward_pass(eeg, ids, labels)
                if out.loss is not None:
                    loss_sum += out.loss.double() * len(labels)
                    n_samples += len(labels)
                generated = self.model.generate(eeg_data=eeg, **gen_config).cpu()
                labels_cpu = labels.cpu()
                for i in range(len(generated)):
                    if self.tokenizer is None:
                        predictions.append(generated[i].tolist())
                        targets.append(labels_cpu[i][labels_cpu[i] != -100].tolist())
                        continue
                    predictions.append(self.tokenizer.decode(generated[i], skip_special_tokens=True,
                                                             clean_up_tokenization_spaces=True).strip())
                    targets.append(self.tokenizer.decode(labels_cpu[i][labels_cpu[i] != -100], skip_special_tokens=True,
                                                         clean_up_tokenization_spaces=True).strip())
        finally:
            self.model.train(was_training)
        self.last_predictions, self.last_targets = predictions, targets
        metrics = {'val_loss': float(loss_sum / n_samples) if n_samples else float('inf')}
        if evaluator is not None:
            metrics.update(evaluator.compute_all_metrics(predictions, targets))
        return metrics

    def train_epoch(self, epoch):
        self.model.train()
        self.epoch = epoch
        loss_sum = torch.zeros((), device=self.device, dtype=torch.float64)
        n_samples = 0
        pending = 0
        for batch in self.train_loader:
            loss = self.train_step(batch)
            pending += 1
            if pending >= self.config['accumulation_steps']:
                self._optimizer_step(step_scheduler=True)
                pending = 0
            n = len(batch['labels'])
            loss_sum += loss.double() * n
            n_samples += n
        if pending > 0:                      # trainer.py:139-145: flush without scheduler.step()
            self._optimizer_step(step_scheduler=False)
        return (loss_sum / n_samples).item() if n_samples else float('inf')

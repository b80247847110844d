"""Composed-model / train-step golden, produced by the REFERENCE itself,
imported from /root/reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_train_golden.py

(1) ``train_step_<shape>.pt``: the reference ``EEGDecodingModel`` (eeg_model.py:11-41; BART built from its
    config because the checkpoint is not cached, SURVEY 8(c) shim 1) stepped by the reference's OWN
    ``EEGTrainer.train_epoch`` (trainer.py:69-151) with the reference's ``get_optimizer_groups``
    (training_config.py:55-77), torch.optim.AdamW (the removed transformers.AdamW's successor) and the
    transformers cosine schedule (train.py:227-231) -- 3 optimizer steps of 2 micro-batches, dropout 0,
    BatchNorm in train mode.  Recorded: the loss of every micro-batch, the pre-clip gradient norm per LR group at
    every optimizer step, the per-parameter update norm ||w_after - w_init|| of every parameter and strided
    samples of the update of a few of them.  Weights come from the shared seeded recipe (param_recipe.py) and the
    batches from seeded generators, so neither is stored.
(2) ``init_routing.json``: which initialisation the reference's ``initialize_custom_weights``
    (scripts/train.py:108-126; the function's source is compiled on its own -- importing the script would start
    wandb / logging set-up) applies to every parameter NAME of the reference model.
"""
import ast
import json
import logging
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
os.environ["WANDB_MODE"] = "disabled"
REF = "/root/reference/main_model"
sys.path.insert(0, REF)
sys.modules.setdefault("jieba", types.ModuleType("jieba"))          # evaluator.py imports it; never called here

from param_recipe import (BART_SHAPE, classify_init, fill_params, fill_sentinel, train_batches,  # noqa: E402
                          zero_bart_dropout, zero_dropout)

import transformers  # noqa: E402
from transformers import BartConfig, BartForConditionalGeneration, get_cosine_schedule_with_warmup  # noqa: E402

BartForConditionalGeneration.from_pretrained = classmethod(
    lambda cls, *a, **k: cls(BartConfig(**BART_SHAPE)))                 # SURVEY 8(c) shim (1)

from config.training_config import CONFIG, get_optimizer_groups  # noqa: E402  (the reference's)
from src.models.eeg_model import EEGDecodingModel  # noqa: E402
from src.training.trainer import EEGTrainer  # noqa: E402

torch.set_num_threads(8)
REGIONS = ["frontal", "temporal", "central", "parietal"]


def reference_init_function():
    src = open(os.path.join(REF, "scripts", "train.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "initialize_custom_weights")
    ns = {"torch": torch, "logger": logging.getLogger("ref")}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "scripts/train.py", "exec"), ns)
    return ns["initialize_custom_weights"]


def init_routing():
    torch.manual_seed(0)
    model = EEGDecodingModel(n_timepoints=33, region_channel_counts={r: 16 for r in REGIONS}, hidden_dim=768)
    fill_sentinel(model)
    reference_init_function()(model)
    table = {name: classify_init(p) for name, p in model.named_parameters()}
    with open(os.path.join(HERE, "init_routing.json"), "w") as fh:
        json.dump(table, fh, indent=0, sort_keys=True)
    kinds = {}
    for v in table.values():
        kinds[v["kind"]] = kinds.get(v["kind"], 0) + 1
    print("init routing:", len(table), "parameters", kinds)


def train_golden(tag, channels, T, B, accum=2, opt_steps=3):
    torch.manual_seed(0)
    counts = {r: channels for r in REGIONS}
    model = EEGDecodingModel(n_timepoints=T, region_channel_counts=counts, hidden_dim=768)
    fill_params(model, seed=13)
    zero_dropout(model)
    zero_bart_dropout(model.bart_decoder.bart)
    w0 = {n: p.detach().clone() for n, p in model.named_parameters()}
    cfg = dict(CONFIG, accumulation_steps=accum)
    opt = torch.optim.AdamW(get_optimizer_groups(model), eps=1e-8, betas=(0.9, 0.999), weight_decay=cfg["weight_decay"])
    sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=2, num_training_steps=50)
    batches = train_batches(accum * opt_steps, B, counts, T, seed=77)

    rec = {"cfg": dict(channels=channels, T=T, B=B, accum=accum, opt_steps=opt_steps, warmup=2, total=50),
           "loss": [], "group_grad_norm": [], "param_grad_norm": [], "total_grad_norm": [], "lr": []}
    # observers on the reference's own step (no change of behaviour): loss per micro-batch, gradient norms at clip time
    real_forward = EEGTrainer.forward_pass

    def forward_pass(self, eeg, ids, labels):
        out = real_forward(self, eeg, ids, labels)
        rec["loss"].append(float(out.loss.detach()))
        return out

    real_clip = torch.nn.utils.clip_grad_norm_

    def clip(params, max_norm, *a, **k):
        groups = {"brain_encoder": [], "eeg_to_bart": [], "bart": []}
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            key = "brain_encoder" if "brain_encoder" in n else ("eeg_to_bart" if "eeg_to_bart" in n else "bart")
            groups[key].append(p.grad.double().pow(2).sum())
        rec["group_grad_norm"].append({k_: float(torch.stack(v).sum().sqrt()) for k_, v in groups.items()})
        rec["param_grad_norm"].append({n: float(p.grad.double().norm()) for n, p in model.named_parameters()
                                       if p.grad is not None})
        rec["lr"].append([g["lr"] for g in opt.param_groups])
        total = real_clip(params, max_norm, *a, **k)
        rec["total_grad_norm"].append(float(total))
        return total

    EEGTrainer.forward_pass = forward_pass
    torch.nn.utils.clip_grad_norm_ = clip
    try:
        trainer = EEGTrainer(model, None, batches, batches, opt, sched, cfg)
        rec["epoch_loss"] = float(trainer.train_epoch(0))
        rec["global_step"] = trainer.global_step
    finally:
        EEGTrainer.forward_pass = real_forward
        torch.nn.utils.clip_grad_norm_ = real_clip
    assert len(rec["loss"]) == accum * opt_steps and len(rec["total_grad_norm"]) == opt_steps, rec
    upd, samples = {}, {}
    for n, p in model.named_parameters():
        d = (p.detach() - w0[n]).flatten()
        upd[n] = float(d.double().norm())
        if n in rec["param_grad_norm"][-1] and any(k in n for k in SAMPLED):
            if True:
                stride = max(1, d.numel() // 4096)
                samples[n] = d[::stride][:4096].clone()
    rec["update_norm"] = upd
    rec["update_sample"] = samples
    rec["no_grad_params"] = sorted(n for n, _ in model.named_parameters() if n not in rec["param_grad_norm"][-1])
    path = os.path.join(HERE, f"train_step_{tag}.pt")
    torch.save(rec, path)
    print(tag, "losses", [round(x, 4) for x in rec["loss"]], "grad norms", [round(x, 3) for x in rec["total_grad_norm"]],
          "lr", rec["lr"], os.path.getsize(path), "bytes")


# parameters whose update is stored as a strided sample (direction check): one of every kind along the path
SAMPLED = ("region_encoders.frontal.conv1.weight", "region_encoders.temporal.conv3.weight",
           "region_encoders.central.attn_layers.0.attn.in_proj_weight", "region_encoders.parietal.attn_layers.2.ffn.linear2.weight",
           "region_encoders.frontal.cross_scale_attn.out_proj.weight", "region_encoders.temporal.pos_emb",
           "region_encoders.central.bn2.weight", "region_encoders.parietal.projection.0.weight",
           "brain_encoder.temporal_scales.1.weight", "brain_encoder.fusion_transformer.layers.0.linear1.weight",
           "brain_encoder.region_importance", "brain_encoder.feature_enhancer.3.weight",
           "eeg_to_bart.0.weight", "decoder.layers.0.self_attn.q_proj.weight", "decoder.layers.5.fc2.weight",
           "decoder.layers.3.encoder_attn.k_proj.weight", "model.shared.weight", "decoder.layernorm_embedding.weight")

if __name__ == "__main__":
    print("transformers", transformers.__version__, "torch", torch.__version__)
    init_routing()
    train_golden("stft", channels=16 * 129, T=33, B=4)
    train_golden("long", channels=32 * 513, T=17, B=2)

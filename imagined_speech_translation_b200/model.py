"""EEG-to-text model on the B200 path: our encoder + the BART decoder.

Drop-in for the reference ``EEGDecodingModel`` / ``BARTDecoder``
(``main_model/src/models/eeg_model.py:11-41``, ``bart_decoder.py:13-79``): same constructor
arguments, attribute / parameter names (``brain_encoder.*``, ``bart_decoder.eeg_to_bart.*``,
``bart_decoder.bart.*`` -- ``get_optimizer_groups`` routes learning rates by these substrings)
and call signatures.  The BART decoder itself is the third-party ``transformers``
implementation the reference also calls (fusing it is "next").
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .brain_encoder import BrainRegionEncoder
from .layers import run_sequential

BART_BASE_CHINESE = dict(vocab_size=51271, d_model=768, encoder_layers=6, decoder_layers=6,
                         encoder_attention_heads=12, decoder_attention_heads=12, encoder_ffn_dim=3072,
                         decoder_ffn_dim=3072, max_position_embeddings=1024, pad_token_id=0,
                         bos_token_id=101, eos_token_id=102, decoder_start_token_id=101)


def _load_bart(pretrained: str):
    """``from_pretrained`` when the checkpoint is in the local HF cache, otherwise the same
    architecture initialised from its config (no network in this environment)."""
    from transformers import BartConfig, BartForConditionalGeneration
    try:
        return BartForConditionalGeneration.from_pretrained(pretrained, local_files_only=True)
    except Exception:
        return BartForConditionalGeneration(BartConfig(**BART_BASE_CHINESE))


class BARTDecoder(nn.Module):
    def __init__(self, hidden_dim, pretrained="fnlp/bart-base-chinese", autocast_dtype=torch.bfloat16):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.bart = _load_bart(pretrained)
        self.bart_dim = self.bart.config.d_model
        self.eeg_to_bart = nn.Sequential(nn.Linear(hidden_dim, self.bart_dim), nn.LayerNorm(self.bart_dim))
        self.autocast_dtype = autocast_dtype

    def create_encoder_sequence(self, eeg_feat):
        B = eeg_feat.shape[0]
        proj = run_sequential(self.eeg_to_bart, eeg_feat.to(torch.bfloat16)).float()
        n = self.bart.config.encoder_layers        # the feature is repeated 6x as the "encoder output"
        return proj.unsqueeze(1).expand(-1, n, -1), torch.ones(B, n, device=eeg_feat.device)

    def forward(self, eeg_feat, decoder_input_ids=None, labels=None, **kwargs):
        from transformers.modeling_outputs import BaseModelOutput
        enc, mask = self.create_encoder_sequence(eeg_feat)
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            return self.bart(input_ids=None, attention_mask=mask,
                             encoder_outputs=BaseModelOutput(last_hidden_state=enc),
                             decoder_input_ids=decoder_input_ids, labels=labels, return_dict=True)

    def generate_from_eeg(self, eeg_feat, max_length=32, **kwargs):
        from transformers.modeling_outputs import BaseModelOutput
        enc, mask = self.create_encoder_sequence(eeg_feat)
        cfg = dict(max_length=max_length, num_beams=3, early_stopping=True,
                   decoder_start_token_id=self.bart.config.decoder_start_token_id)
        cfg.update(kwargs)
        return self.bart.generate(encoder_outputs=BaseModelOutput(last_hidden_state=enc.contiguous()),
                                  attention_mask=mask, **cfg)


class EEGDecodingModel(nn.Module):
    def __init__(self, n_timepoints, region_channel_counts, hidden_dim=768, disable_cross_region_attn=False,
                 uniform_region_weight=False, cnn_only=False):
        super().__init__()
        self.brain_encoder = BrainRegionEncoder(
            n_timepoints=n_timepoints, region_channel_counts=region_channel_counts, hidden_dim=hidden_dim,
            disable_cross_region_attn=disable_cross_region_attn, uniform_region_weight=uniform_region_weight,
            cnn_only=cnn_only)
        self.bart_decoder = BARTDecoder(hidden_dim=hidden_dim)

    def forward(self, eeg_data, decoder_input_ids=None, labels=None, **kwargs):
        return self.bart_decoder(eeg_feat=self.brain_encoder(eeg_data), decoder_input_ids=decoder_input_ids,
                                 labels=labels, **kwargs)

    def generate(self, eeg_data, **kwargs):
        return self.bart_decoder.generate_from_eeg(self.brain_encoder(eeg_data), **kwargs)

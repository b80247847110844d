"""Batched beam-search decoding for ``BARTDecoder.generate_from_eeg``.

The reference evaluates with ``self.bart.generate(encoder_outputs=..., num_beams=3, early_stopping=True,
max_length=16, min_length=4)`` (``main_model/src/models/bart_decoder.py:59-79``, ``config/training_config.py:32-39``,
called per validation batch from ``trainer.py:153-212``), i.e. the algorithm lives in the third-party ``transformers``
package (reference pinned 4.46.2; installed here and used as the oracle: 5.5.0, ``generation/utils.py``
``GenerationMixin._beam_search`` and its helpers, ``generation/logits_process.py`` MinLength / ForcedEOS processors,
``generation/stopping_criteria.py`` MaxLength / EosToken criteria).  This module restates that published algorithm:

* ``beam_search`` -- the bookkeeping, device-resident and static-shaped, generic over a ``step_logits`` callable
  (so it is pinned exactly: driven by the stock model's own logits it returns ``generate``'s sequences bit for bit,
  ``tests/test_generation.py``);
* ``decoder_step_logits`` -- the model side on our kernels: the decoder over the current prefixes (tcgen05 GEMMs,
  fused attention / LayerNorm kernels, eval mode) and the LM head on the LAST position only, fp32 logits.

* ``CachedDecoder`` -- the same decoder one token at a time with a key/value cache: per layer the self-attention
  keys/values of the prefix are kept as one (rows, t, 2d) bf16 tensor that is gathered by the parent-beam index and
  extended by the new token each step (so it is always compact and the fused attention kernel reads it as an
  ordinary (B*Sk, 2d) key/value matrix); the cross-attention keys/values of the 6-vector EEG memory are projected
  once.  This is what ``generate`` uses; the re-encoding ``decoder_step_logits`` stays as its cross-check.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

NEG = -1.0e9          # the "cannot be chosen" score of the algorithm (not -inf: scores stay finite and sortable)


def _take(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Select beams along dim 1 (``idx``: (B, k))."""
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


@torch.no_grad()
def beam_search(step_logits: Callable[[torch.Tensor], torch.Tensor], batch_size: int, vocab_size: int, *,
                num_beams: int = 3, max_length: int = 16, min_length: int = 0, decoder_start_token_id: int,
                eos_token_id: Optional[int], pad_token_id: Optional[int] = None,
                forced_eos_token_id: Optional[int] = None, early_stopping=True, length_penalty: float = 1.0,
                device="cuda", select: Optional[Callable] = None) -> torch.Tensor:
    """Beam search as ``transformers`` 5.5 runs it for an encoder-decoder model whose prompt is the single
    ``decoder_start_token_id``.  ``step_logits(prefixes, parents)``: prefixes (batch*beams, cur_len) int64, parents
    (batch*beams,) int64 = for every row the row of the PREVIOUS call it continues (None on the first call; what a
    key/value cache is re-ordered by) -> (batch*beams, vocab) next-token logits.  Returns (batch, <= max_length) token ids, the best finished hypothesis per batch item,
    filled with ``pad_token_id`` (or eos) after its end.

    Per step: log-softmax, MinLength (eos = -inf while cur_len < min_length) and ForcedEOS (at cur_len ==
    max_length - 1 only ``forced_eos_token_id`` survives) processors, add the running beam scores, keep the top
    2 * beams continuations over all beams, retire those that ended (eos or max_length) into the finished set
    (score / generated_length ** length_penalty, best ``num_beams`` kept), continue with the best ``num_beams``
    unfinished ones; stop when every batch item has ``num_beams`` finished hypotheses (early_stopping=True), when
    no running beam can beat the worst finished one, or when nothing can continue.

    ``select(logits, k, banned)`` -> (log-probabilities (rows, k) descending, token ids (rows, k)): optional fused
    per-row log-softmax + top-k (``topk_logprobs``).  The best ``2 * beams`` continuations of a batch item are always
    among the best ``2 * beams`` of each of its beams, so the V-wide log-softmax / add / top-k of the plain path
    shrink to a top-k over ``beams * 2 * beams`` numbers."""
    if num_beams < 2:
        raise ValueError("beam_search needs num_beams >= 2 (the library decodes greedily through another routine)")
    dev = torch.device(device)
    B, nb, V = batch_size, num_beams, vocab_size
    keep = 2 * nb                                                 # (number of eos tokens + 1) * beams, at least 2x
    fill = pad_token_id if pad_token_id else (eos_token_id if eos_token_id is not None else -1)
    prompt = 1
    cur = prompt
    running = torch.full((B, nb, max_length), fill, dtype=torch.int64, device=dev)
    running[:, :, 0] = decoder_start_token_id
    finished = running.clone()
    run_score = torch.zeros(B, nb, device=dev)
    run_score[:, 1:] = NEG                                        # all beams start identical: only the first counts
    fin_score = torch.full((B, nb), NEG, device=dev)
    is_fin = torch.zeros(B, nb, dtype=torch.bool, device=dev)
    can_improve = torch.ones(B, 1, dtype=torch.bool, device=dev)
    fin_len = torch.zeros(B, nb, dtype=torch.int64, device=dev)   # generated length of the stored hypotheses
    top_mask = torch.zeros(keep, dtype=torch.bool, device=dev)
    top_mask[:nb] = True
    strict = early_stopping is True
    parents = None
    row0 = torch.arange(B, device=dev).view(B, 1) * nb

    while True:
        logits = step_logits(running[:, :, :cur].reshape(B * nb, cur), parents).to(torch.float32)
        ban_eos = eos_token_id is not None and cur < min_length
        force = forced_eos_token_id is not None and cur == max_length - 1
        if select is None:
            logp = torch.log_softmax(logits, dim=-1)
            if ban_eos:
                logp[:, eos_token_id] = float("-inf")
            if force:
                forced = torch.full_like(logp, float("-inf"))
                forced[:, forced_eos_token_id] = 0
                logp = forced
            acc = (logp.view(B, nb, V) + run_score[:, :, None]).view(B, nb * V)
            top_val, top_idx = torch.topk(acc, k=keep)
            src_beam, tokens = top_idx // V, top_idx % V
        else:
            kk = min(keep, V)
            if force:                                             # one continuation per beam, log-probability 0
                row_val = torch.full((B * nb, kk), float("-inf"), device=dev)
                row_val[:, 0] = 0
                row_idx = torch.full((B * nb, kk), forced_eos_token_id, dtype=torch.int64, device=dev)
            else:
                row_val, row_idx = select(logits, kk, eos_token_id if ban_eos else -1)
            acc = (row_val.view(B, nb, kk) + run_score[:, :, None]).view(B, nb * kk)
            top_val, pos = torch.topk(acc, k=keep)
            src_beam, tokens = pos // kk, torch.gather(row_idx.view(B, nb * kk), 1, pos)
        cand = _take(running, src_beam)
        cand[:, :, cur] = tokens
        # which candidates just ended: eos as last token, or the length limit
        ended = torch.full((B, keep), cur + 1 >= max_length, dtype=torch.bool, device=dev)
        if eos_token_id is not None:
            ended = ended | (cand[:, :, cur] == eos_token_id)
        # the best `nb` that have NOT ended keep running
        open_val = top_val + ended.to(torch.float32) * NEG
        nxt = torch.topk(open_val, k=nb)[1]
        running, run_score = _take(cand, nxt), _take(open_val, nxt)
        parents = (_take(src_beam, nxt) + row0).reshape(B * nb)
        # retire the ended ones among the top `nb` candidates into the finished set
        just = ended & top_mask[None, :]
        score = top_val / float((cur + 1 - prompt) ** length_penalty)
        score = score + (is_fin.all(dim=-1, keepdim=True) & strict).to(torch.float32) * NEG
        score = score + (~can_improve).to(torch.float32) * NEG
        score = score + (~just) * NEG
        m_seq = torch.cat((finished, cand), dim=1)
        m_score = torch.cat((fin_score, score), dim=1)
        m_fin = torch.cat((is_fin, just), dim=1)
        m_len = torch.cat((fin_len, torch.full((B, keep), cur + 1 - prompt, dtype=torch.int64, device=dev)), dim=1)
        best = torch.topk(m_score, k=nb)[1]
        finished, fin_score, is_fin, fin_len = _take(m_seq, best), _take(m_score, best), _take(m_fin, best), \
            _take(m_len, best)
        cur += 1
        # can any running beam still beat the worst finished hypothesis of its batch item?
        hyp_len = (max_length - prompt) if (early_stopping == "never" and length_penalty > 0.0) else (cur - prompt)
        best_running = run_score[:, :1] / float(hyp_len ** length_penalty)
        worst_fin = torch.where(is_fin, fin_score.min(dim=1, keepdim=True)[0], torch.full_like(fin_score, NEG))
        can_improve = can_improve & (best_running > worst_fin).any(dim=-1, keepdim=True)
        go_on = can_improve.any() & ~(is_fin.all() & strict) & ~ended.all()
        if not bool(go_on):                                       # the one host sync per step (as in the library loop)
            break

    out = finished[:, 0, :]
    out_len = prompt + int(fin_len[:, 0].max())
    return out[:, :out_len]


def topk_logprobs(logits: torch.Tensor, k: int, banned: int = -1):
    """Fused per-row log-softmax + top-k (``eegx_logsoftmax_topk_f32``): one read of the fp32 logits."""
    from . import _lib
    if not (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1):
        raise _lib.EegxError("topk_logprobs needs a CUDA float32 (rows, V) tensor with unit inner stride")
    rows, V = logits.shape
    val = torch.empty(rows, k, dtype=torch.float32, device=logits.device)
    idx = torch.empty(rows, k, dtype=torch.int64, device=logits.device)
    _lib.check(_lib.lib().eegx_logsoftmax_topk_f32(_lib.ptr(logits), logits.stride(0), rows, V, int(k), int(banned),
                                                   _lib.ptr(val), _lib.ptr(idx), _lib.stream_ptr()),
               "eegx_logsoftmax_topk_f32")
    return val, idx


# ---------------------------------------------------------------------------------------------- model side
@torch.no_grad()
def decoder_step_logits(bart_decoder, mem: torch.Tensor, prefixes: torch.Tensor, beams: int) -> torch.Tensor:
    """Next-token logits (fp32, (rows, V)) for ``prefixes`` (rows = batch*beams, cur_len) against the EEG memory
    ``mem`` ((batch*n_mem, d) bf16, one block per batch item): the decoder of ``BARTDecoder._forward_fused`` in
    eval mode, then the LM head + final_logits_bias on the last position only."""
    from . import nn_ops, ops
    bart = bart_decoder.bart
    rows, L = prefixes.shape
    d = bart.config.d_model
    n_mem = mem.shape[0] * beams // rows
    mem_rows = mem.view(rows // beams, 1, n_mem, d).expand(-1, beams, -1, -1).reshape(rows * n_mem, d)
    h = bart_decoder._decoder_hidden(mem_rows, prefixes, training=False)          # (rows * L, d) bf16
    last = h.view(rows, L, d)[:, -1].contiguous()
    w16 = nn_ops._w_linear(bart.lm_head.weight)                                   # (V padded to 8, d) bf16
    V = bart.lm_head.weight.shape[0]
    bias = torch.zeros(w16.shape[0], dtype=torch.float32, device=last.device)
    bias[:V] = bart.final_logits_bias.reshape(-1).float()
    return ops.gemm(last, w16, bias, out_dtype=torch.float32)[:, :V]


class CachedDecoder:
    """``step_logits`` callable for ``beam_search``: one new token per row and step, self-attention over a
    key/value cache, cross-attention over EEG-memory keys/values projected once (BartDecoderLayer,
    modeling_bart.py, eval mode; same kernels as ``BARTDecoder._forward_fused``)."""

    def __init__(self, bart_decoder, mem: torch.Tensor, batch: int, beams: int):
        from . import nn_ops
        self.owner = bart_decoder
        bart = bart_decoder.bart
        self.dec = bart.model.decoder
        self.d = bart.config.d_model
        self.rows = batch * beams
        self.n_mem = mem.shape[0] // batch
        self.cross = []
        for layer in self.dec.layers:
            a = layer.encoder_attn
            kv = nn_ops.linear_cat(mem, a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias)   # (B*n_mem, 2d)
            kv = kv.view(batch, 1, self.n_mem, 2 * self.d).expand(-1, beams, -1, -1)
            self.cross.append(kv.reshape(self.rows * self.n_mem, 2 * self.d).contiguous())
        self.cache = [None] * len(self.dec.layers)
        self.w16 = nn_ops._w_linear(bart.lm_head.weight)
        self.V = bart.lm_head.weight.shape[0]
        self.bias = torch.zeros(self.w16.shape[0], dtype=torch.float32, device=mem.device)
        self.bias[:self.V] = bart.final_logits_bias.reshape(-1).float()

    @torch.no_grad()
    def __call__(self, prefixes: torch.Tensor, parents: Optional[torch.Tensor]) -> torch.Tensor:
        return self.step(prefixes[:, -1], prefixes.shape[1] - 1, parents)

    @torch.no_grad()
    def step(self, tokens: torch.Tensor, t: int, parents: Optional[torch.Tensor]) -> torch.Tensor:
        """tokens: (rows,) the token at position t of every row."""
        from . import fused, nn_ops, ops
        dec, d, rows = self.dec, self.d, self.rows
        emb = dec.embed_tokens(tokens)                                      # includes embed_scale
        pos = dec.embed_positions.weight[dec.embed_positions.offset + t]
        ln = dec.layernorm_embedding
        h = fused.layer_norm((emb + pos).to(torch.bfloat16), ln.weight, ln.bias, ln.eps, p=0.0, training=False)
        for i, layer in enumerate(dec.layers):
            a = layer.self_attn
            H = a.num_heads
            qkv = nn_ops.linear_cat(h, a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.k_proj.bias,
                                    a.v_proj.weight, a.v_proj.bias)         # (rows, 3d)
            kv_new = qkv[:, d:].reshape(rows, 1, 2 * d)
            if self.cache[i] is None:
                self.cache[i] = kv_new.contiguous()
            else:
                self.cache[i] = torch.cat((self.cache[i].index_select(0, parents), kv_new), dim=1)
            o = fused.attn_cross(qkv[:, :d], self.cache[i].view(rows * (t + 1), 2 * d), rows, 1, t + 1, H,
                                 training=False)
            o = nn_ops.linear(o, a.out_proj.weight, a.out_proj.bias)
            ln = layer.self_attn_layer_norm
            h = fused.layer_norm(fused.add_dropout(h, o, training=False), ln.weight, ln.bias, ln.eps)
            a = layer.encoder_attn
            q = nn_ops.linear(h, a.q_proj.weight, a.q_proj.bias)
            o = fused.attn_cross(q, self.cross[i], rows, 1, self.n_mem, H, training=False)
            o = nn_ops.linear(o, a.out_proj.weight, a.out_proj.bias)
            ln = layer.encoder_attn_layer_norm
            h = fused.layer_norm(fused.add_dropout(h, o, training=False), ln.weight, ln.bias, ln.eps)
            f = nn_ops.linear(h, layer.fc1.weight, layer.fc1.bias)
            f = fused.gelu_dropout(f, p=0.0, training=False)
            f = nn_ops.linear(f, layer.fc2.weight, layer.fc2.bias)
            ln = layer.final_layer_norm
            h = fused.layer_norm(fused.add_dropout(h, f, training=False), ln.weight, ln.bias, ln.eps)
        return ops.gemm(h, self.w16, self.bias, out_dtype=torch.float32)[:, :self.V]


class GraphedDecoder:
    """``CachedDecoder`` with every step captured in its own CUDA graph (one per position: the cache grows, so
    shapes differ between steps but are fixed per step).  A decode step is ~130 small launches; issued from Python
    they take longer on the host than on the GPU.  Graph 0 also holds the bf16 weight packs and the memory
    key/value projections, so every ``bind`` + replay of step 0 re-derives them from the live weights."""

    def __init__(self, bart_decoder, batch: int, beams: int, max_length: int, device):
        from . import nn_ops
        cfg = bart_decoder.bart.config
        self.rows, self.steps = batch * beams, max_length - 1
        self.mem = torch.zeros(batch * cfg.encoder_layers, cfg.d_model, dtype=torch.bfloat16, device=device)
        self.tok = [torch.zeros(self.rows, dtype=torch.int64, device=device) for _ in range(self.steps)]
        self.par = [torch.zeros(self.rows, dtype=torch.int64, device=device) for _ in range(self.steps)]
        self.signature = self.weights_signature(bart_decoder)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):                               # eager dry run: lazy initialisations happen here
            core = CachedDecoder(bart_decoder, self.mem, batch, beams)
            for t in range(min(2, self.steps)):
                core.step(self.tok[t], t, self.par[t] if t else None)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        nn_ops.clear_pack_cache()                                   # the packs must be (re)built INSIDE graph 0
        self.graphs, self.out, pool, core = [], [], None, None
        for t in range(self.steps):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                if t == 0:
                    core = CachedDecoder(bart_decoder, self.mem, batch, beams)
                self.out.append(core.step(self.tok[t], t, self.par[t] if t else None))
            pool = g.pool()
            self.graphs.append(g)
        nn_ops.clear_pack_cache()                                   # eager code builds its own packs again

    @staticmethod
    def weights_signature(bart_decoder):
        return tuple(p.data_ptr() for p in bart_decoder.bart.model.decoder.parameters()) + \
            (bart_decoder.bart.lm_head.weight.data_ptr(), bart_decoder.bart.final_logits_bias.data_ptr())

    def bind(self, mem: torch.Tensor) -> "GraphedDecoder":
        self.mem.copy_(mem)
        return self

    def __call__(self, prefixes: torch.Tensor, parents: Optional[torch.Tensor]) -> torch.Tensor:
        t = prefixes.shape[1] - 1
        self.tok[t].copy_(prefixes[:, -1])
        if parents is not None:
            self.par[t].copy_(parents)
        self.graphs[t].replay()
        return self.out[t]


SUPPORTED = {"max_length", "min_length", "num_beams", "early_stopping", "decoder_start_token_id", "length_penalty"}

# generation_config entries that switch on a logits processor / stopping criterion / decoding mode this path does
# not implement, with the value that means "off" (transformers GenerationConfig defaults)
_INACTIVE = {
    "do_sample": (False, None), "num_beam_groups": (1, None), "diversity_penalty": (0.0, None), "penalty_alpha": (None,),
    "repetition_penalty": (1.0, None), "encoder_repetition_penalty": (1.0, None),
    "no_repeat_ngram_size": (0, None), "encoder_no_repeat_ngram_size": (0, None),
    "bad_words_ids": (None,), "force_words_ids": (None,), "constraints": (None,), "sequence_bias": (None,),
    "forced_bos_token_id": (None,), "suppress_tokens": (None,), "begin_suppress_tokens": (None,),
    "forced_decoder_ids": (None,), "renormalize_logits": (False, None), "exponential_decay_length_penalty": (None,),
    "remove_invalid_values": (False, None), "guidance_scale": (None, 1.0), "num_return_sequences": (1, None),
    "max_new_tokens": (None,), "min_new_tokens": (None,), "stop_strings": (None,), "max_time": (None,),
    "dola_layers": (None,), "watermarking_config": (None,), "token_healing": (False, None),
    "output_scores": (False, None), "output_logits": (False, None, ), "return_dict_in_generate": (False, None),
}


def unsupported_generation_options(generation_config) -> list:
    """Names of the options in a model's (checkpoint-inherited) ``generation_config`` that are active but not
    implemented by ``beam_search`` -- ``transformers.generate`` applies them silently (e.g. ``no_repeat_ngram_size``
    of a fine-tuned checkpoint), so the native path must not be taken when this list is non-empty."""
    active = []
    for name, off in _INACTIVE.items():
        val = getattr(generation_config, name, None)
        if isinstance(val, (list, tuple, dict)) and len(val) == 0:
            continue
        if not any(val is o or (o is not None and not isinstance(val, (list, tuple, dict)) and val == o) for o in off):
            active.append(name)
    return active


@torch.no_grad()
def generate(bart_decoder, eeg_feat: torch.Tensor, **gen) -> torch.Tensor:
    """``BARTDecoder.generate_from_eeg`` on our kernels; ``gen`` holds the (already defaulted) generation options."""
    from .layers import run_sequential
    bart = bart_decoder.bart
    cfg, gc = bart.config, bart.generation_config
    B = eeg_feat.shape[0]
    nb = int(gen.get("num_beams", 1))
    proj = run_sequential(bart_decoder.eeg_to_bart, eeg_feat.to(torch.bfloat16))                 # (B, d)
    n_mem = cfg.encoder_layers
    mem = proj.unsqueeze(1).expand(B, n_mem, cfg.d_model).reshape(B * n_mem, cfg.d_model).contiguous()
    eos = gc.eos_token_id if gc.eos_token_id is not None else cfg.eos_token_id
    if isinstance(eos, (list, tuple)):
        if len(eos) != 1:
            raise ValueError("several eos tokens are not supported on this path")
        eos = eos[0]
    max_length = int(gen.get("max_length", 32))
    mode = getattr(bart_decoder, "generate_mode", "graph")
    if mode == "graph":
        cache = bart_decoder.__dict__.setdefault("_gen_graphs", {})
        key = (B, nb, max_length, eeg_feat.device.index)
        graphed = cache.get(key)
        if graphed is None or graphed.signature != GraphedDecoder.weights_signature(bart_decoder):
            cache.clear()                                          # one configuration at a time (each holds a cache pool)
            graphed = cache[key] = GraphedDecoder(bart_decoder, B, nb, max_length, eeg_feat.device)
        step = graphed.bind(mem)
    elif mode == "cache":
        step = CachedDecoder(bart_decoder, mem, B, nb)
    else:                                                          # "reencode"
        step = lambda p, parents: decoder_step_logits(bart_decoder, mem, p, nb)      # noqa: E731
    return beam_search(step, B, cfg.vocab_size,
                       num_beams=nb, max_length=max_length,
                       min_length=int(gen.get("min_length", gc.min_length or 0)),
                       decoder_start_token_id=int(gen.get("decoder_start_token_id", cfg.decoder_start_token_id)),
                       eos_token_id=eos, pad_token_id=gc.pad_token_id if gc.pad_token_id is not None else cfg.pad_token_id,
                       forced_eos_token_id=gc.forced_eos_token_id, early_stopping=gen.get("early_stopping", False),
                       length_penalty=float(gen.get("length_penalty", gc.length_penalty if gc.length_penalty is not None else 1.0)),
                       device=eeg_feat.device,
                       select=topk_logprobs if getattr(bart_decoder, "fused_select", True) else None)

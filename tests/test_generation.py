"""Beam-search generation.

CPU:  generation.beam_search driven by the stock transformers model's own logits must return exactly what
      transformers' generate() returns (token ids: bit-exact) -- this pins the bookkeeping restatement against the
      installed library, which is what the reference calls (bart_decoder.py:59-79).
GPU:  the model side on our kernels: next-token logits against the stock fp32 forward (stated bf16 bound), and the
      end-to-end generate_from_eeg against transformers.generate on a model with a peaked output distribution.
"""
import pytest
import torch

from imagined_speech_translation_b200 import generation


def _tiny_bart(seed, vocab=60, eos=5):
    from transformers import BartConfig, BartForConditionalGeneration
    torch.manual_seed(seed)
    cfg = BartConfig(vocab_size=vocab, d_model=32, encoder_layers=1, decoder_layers=2, encoder_attention_heads=2,
                     decoder_attention_heads=2, encoder_ffn_dim=64, decoder_ffn_dim=64, max_position_embeddings=64,
                     pad_token_id=0, bos_token_id=1, eos_token_id=eos, decoder_start_token_id=1,
                     forced_eos_token_id=2, init_std=0.5)
    return BartForConditionalGeneration(cfg).eval()


def _torch_select(logits, k, banned):
    """Reference for the fused per-row log-softmax + top-k."""
    logp = torch.log_softmax(logits.float(), dim=-1)
    if banned >= 0:
        logp[:, banned] = float("-inf")
    return torch.topk(logp, k=k, dim=-1)


def _library_and_ours(model, enc, select=None, **gen):
    from transformers.modeling_outputs import BaseModelOutput
    B, n_mem, _ = enc.shape
    mask = torch.ones(B, n_mem)
    cfg = model.config
    with torch.no_grad():
        want = model.generate(encoder_outputs=BaseModelOutput(last_hidden_state=enc.clone()), attention_mask=mask,
                              decoder_start_token_id=cfg.decoder_start_token_id, use_cache=False, **gen)
    nb = gen["num_beams"]

    def step(prefixes, parents):
        e = enc.repeat_interleave(nb, dim=0)
        out = model(encoder_outputs=BaseModelOutput(last_hidden_state=e), attention_mask=mask.repeat_interleave(nb, 0),
                    decoder_input_ids=prefixes, use_cache=False, return_dict=True)
        return out.logits[:, -1, :]

    got = generation.beam_search(step, B, cfg.vocab_size, num_beams=nb, max_length=gen["max_length"],
                                 min_length=gen.get("min_length", 0),
                                 decoder_start_token_id=cfg.decoder_start_token_id, eos_token_id=cfg.eos_token_id,
                                 pad_token_id=cfg.pad_token_id,
                                 forced_eos_token_id=model.generation_config.forced_eos_token_id,
                                 early_stopping=gen.get("early_stopping", False),
                                 length_penalty=gen.get("length_penalty", 1.0), device="cpu", select=select)
    return want, got


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("gen", [
    dict(num_beams=3, max_length=16, min_length=4, early_stopping=True),          # the reference's eval config
    dict(num_beams=3, max_length=32, early_stopping=True),                         # generate_from_eeg defaults
    dict(num_beams=4, max_length=12, min_length=2, early_stopping=False, length_penalty=2.0),
    dict(num_beams=2, max_length=9, early_stopping="never", length_penalty=0.5),
    dict(num_beams=5, max_length=10, min_length=3, early_stopping=True),
])
def test_beam_search_equals_transformers_generate(seed, gen):
    model = _tiny_bart(seed)
    torch.manual_seed(100 + seed)
    enc = torch.randn(5, 6, 32)
    want, got = _library_and_ours(model, enc, **gen)
    assert want.shape == got.shape, (want.shape, got.shape)
    assert torch.equal(want, got)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_beam_search_with_per_row_selection_equals_transformers_generate(seed):
    """The per-row top-2k shortcut (what the fused kernel feeds) selects exactly the library's continuations."""
    model = _tiny_bart(seed)
    torch.manual_seed(200 + seed)
    enc = torch.randn(6, 6, 32)
    for gen in (dict(num_beams=3, max_length=16, min_length=4, early_stopping=True),
                dict(num_beams=4, max_length=12, min_length=2, early_stopping=False, length_penalty=2.0)):
        want, got = _library_and_ours(model, enc, select=_torch_select, **gen)
        assert want.shape == got.shape and torch.equal(want, got)


def test_beam_search_early_finish_and_fill():
    """A model that emits eos immediately after min_length: every hypothesis ends early, the output is cropped to
    the longest one and the library's fill rule (pad_token_id, or eos when pad is 0 / None) is followed."""
    V, eos = 11, 7

    def step(prefixes, parents):
        logits = torch.zeros(prefixes.shape[0], V)
        logits[:, 3] = 2.0
        if prefixes.shape[1] >= 3:
            logits[:, eos] = 9.0
        return logits

    out = generation.beam_search(step, 2, V, num_beams=3, max_length=16, min_length=3, decoder_start_token_id=1,
                                 eos_token_id=eos, pad_token_id=0, early_stopping=True, device="cpu")
    assert out.tolist() == [[1, 3, 3, eos]] * 2


# ------------------------------------------------------------------------------------------------- GPU
def _decoder(seed=0, peaked=True):
    from imagined_speech_translation_b200.model import BARTDecoder
    torch.manual_seed(seed)
    dec = BARTDecoder(hidden_dim=768).cuda().eval()
    if peaked:
        # random-init logits are nearly flat (std 0.02 embeddings): give the output distribution structure so that
        # beam decisions are not decided by the last bit -- larger tied embedding / LM head and a spread-out bias
        with torch.no_grad():
            dec.bart.model.shared.weight.mul_(6.0)
            dec.bart.final_logits_bias.copy_(torch.randn_like(dec.bart.final_logits_bias) * 1.5)
    return dec


@pytest.mark.gpu
def test_step_logits_match_stock_forward():
    from transformers.modeling_outputs import BaseModelOutput
    dec = _decoder()
    torch.manual_seed(1)
    B, nb = 4, 3
    feat = torch.randn(B, 768, device="cuda")
    from imagined_speech_translation_b200.layers import run_sequential
    proj = run_sequential(dec.eeg_to_bart, feat.to(torch.bfloat16))
    mem = proj.unsqueeze(1).expand(B, 6, 768).reshape(B * 6, 768).contiguous()
    for L in (1, 2, 7, 16):
        prefixes = torch.randint(1, 51271, (B * nb, L), device="cuda")
        prefixes[:, 0] = 101
        got = generation.decoder_step_logits(dec, mem, prefixes, nb)
        enc = proj.float().unsqueeze(1).expand(B, 6, 768).repeat_interleave(nb, dim=0)
        with torch.no_grad():
            want = dec.bart(encoder_outputs=BaseModelOutput(last_hidden_state=enc), decoder_input_ids=prefixes,
                            attention_mask=torch.ones(B * nb, 6, device="cuda"), use_cache=False).logits[:, -1].float()
        assert got.shape == want.shape == (B * nb, 51271) and got.dtype == torch.float32
        rel = float((got - want).abs().max() / want.abs().max())
        assert rel <= 5e-2, (L, rel)                       # bf16 activations against the fp32 library forward
        lp_g, lp_w = torch.log_softmax(got, -1), torch.log_softmax(want, -1)
        top = lp_w.topk(5, dim=-1).indices
        assert float((lp_g.gather(1, top) - lp_w.gather(1, top)).abs().max()) <= 0.15, L


@pytest.mark.gpu
@pytest.mark.parametrize("rows,V,k,banned", [(7, 51271, 6, 102), (3, 51271, 6, -1), (5, 1000, 16, 0), (4, 9, 8, 3),
                                             (2, 300, 1, -1)])
def test_fused_logsoftmax_topk_matches_torch(rows, V, k, banned):
    torch.manual_seed(rows + V)
    ld = (V + 7) // 8 * 8
    buf = torch.randn(rows, ld, device="cuda") * 3
    logits = buf[:, :V]
    logits[0, : min(V, 5)] = 2.5                                   # exact ties: lower index first
    val, idx = generation.topk_logprobs(logits, k, banned)
    want_v, want_i = _torch_select(logits.clone(), k, banned)
    assert float((val - want_v).abs().max()) <= 2e-5
    ties = want_v[:, 1:] == want_v[:, :-1]
    keep = torch.ones_like(want_i, dtype=torch.bool)
    keep[:, 1:] &= ~ties
    keep[:, :-1] &= ~ties
    assert torch.equal(idx[keep], want_i[keep])
    if banned >= 0:
        assert not (idx == banned).any()
    # a row start that is not 16-byte aligned takes the one-element-per-thread path: same answer
    shifted = torch.empty(rows, ld + 8, device="cuda")[:, 1:1 + V]
    shifted.copy_(logits)
    val2, idx2 = generation.topk_logprobs(shifted, k, banned)
    assert torch.equal(idx2, idx) and float((val2 - val).abs().max()) <= 2e-6
    same = val[0, 1:] == val[0, :-1]                               # among equal values: ascending token ids
    assert bool((idx[0, 1:][same] > idx[0, :-1][same]).all())


@pytest.mark.gpu
def test_cached_decoder_matches_reencoding_under_beam_reordering():
    """Key/value-cache path against re-encoding the whole prefix, with random parent permutations between steps
    (the cache must follow the beams).  Same kernels, same bf16 roundings: logits agree to 1e-3 relative."""
    dec = _decoder()
    torch.manual_seed(3)
    B, nb = 5, 3
    rows = B * nb
    from imagined_speech_translation_b200.layers import run_sequential
    proj = run_sequential(dec.eeg_to_bart, torch.randn(B, 768, device="cuda").to(torch.bfloat16))
    mem = proj.unsqueeze(1).expand(B, 6, 768).reshape(B * 6, 768).contiguous()
    cached = generation.CachedDecoder(dec, mem, B, nb)
    prefixes = torch.full((rows, 1), 101, device="cuda")
    parents = None
    for t in range(8):
        got = cached(prefixes, parents)
        want = generation.decoder_step_logits(dec, mem, prefixes, nb)
        rel = float((got - want).abs().max() / want.abs().max())
        assert rel <= 1e-3, (t, rel)
        # next step: every row continues a random row of ITS batch item
        parents = (torch.randint(0, nb, (B, nb), device="cuda") + torch.arange(B, device="cuda")[:, None] * nb).reshape(-1)
        prefixes = torch.cat((prefixes[parents], torch.randint(1, 51271, (rows, 1), device="cuda")), dim=1)


@pytest.mark.gpu
def test_generate_from_eeg_matches_transformers_generate():
    """End to end against the library path on the same weights.  Token ids are discrete decisions on bf16 logits, so
    the bar is stated as agreement rates: >= 90 % of sequences identical, >= 97 % of tokens."""
    dec = _decoder()
    torch.manual_seed(2)
    feat = torch.randn(48, 768, device="cuda")
    gen = dict(max_length=16, min_length=4, num_beams=3, early_stopping=True)
    got = dec.generate_from_eeg(feat, **gen)                                   # CUDA-graphed steps (default)
    assert torch.equal(dec.generate_from_eeg(feat, **gen), got)                # replaying the graphs: same result
    for mode in ("cache", "reencode"):
        dec.generate_mode = mode
        other = dec.generate_from_eeg(feat, **gen)
        n_ = min(other.shape[1], got.shape[1])
        assert (other[:, :n_] == got[:, :n_]).float().mean().item() >= 0.97, mode
    dec.generate_mode = "graph"
    with torch.no_grad():                                                      # weights change in place: graph 0 re-packs them
        dec.bart.final_logits_bias.add_(torch.randn_like(dec.bart.final_logits_bias))
    got = dec.generate_from_eeg(feat, **gen)
    dec.native_generate = False
    dec.autocast_dtype = None
    want = dec.generate_from_eeg(feat, **gen)
    assert got.shape[0] == want.shape[0] == 48 and got.dtype == torch.int64
    n = min(got.shape[1], want.shape[1])
    same_tok = (got[:, :n] == want[:, :n]).float().mean().item()
    same_seq = sum(torch.equal(a[:n], b[:n]) for a, b in zip(got, want)) / 48
    assert got.shape[1] == want.shape[1]
    assert (got[:, 0] == 101).all() and same_seq >= 0.90 and same_tok >= 0.97, (same_seq, same_tok)


@pytest.mark.gpu
def test_generate_unsupported_options_use_the_library_path():
    dec = _decoder(peaked=False)
    feat = torch.randn(2, 768, device="cuda")
    out = dec.generate_from_eeg(feat, max_length=8, num_beams=2, no_repeat_ngram_size=2)     # not in generation.SUPPORTED
    assert out.shape[0] == 2 and out.shape[1] <= 8


def test_checkpoint_generation_config_options_are_detected():
    """Options a checkpoint's generation_config switches on (transformers.generate applies them silently) must be
    reported, so generate_from_eeg leaves the native beam search for the library path."""
    from transformers import GenerationConfig
    assert generation.unsupported_generation_options(GenerationConfig()) == []
    assert generation.unsupported_generation_options(_tiny_bart(0).generation_config) == []
    for name, val in (("no_repeat_ngram_size", 3), ("repetition_penalty", 1.2), ("bad_words_ids", [[7]]),
                      ("num_beam_groups", 2), ("do_sample", True), ("suppress_tokens", [4]), ("num_return_sequences", 2)):
        gc = GenerationConfig()
        setattr(gc, name, val)
        assert generation.unsupported_generation_options(gc) == [name]


@pytest.mark.gpu
def test_generate_honours_the_checkpoints_generation_config():
    """no_repeat_ngram_size set on bart.generation_config (not passed as a kwarg): the result must be the
    library's, which never repeats a bigram -- the native path would."""
    from transformers.modeling_outputs import BaseModelOutput
    dec = _decoder(peaked=True)
    feat = torch.randn(3, 768, device="cuda")
    dec.bart.generation_config.no_repeat_ngram_size = 2
    try:
        out = dec.generate_from_eeg(feat, max_length=12, min_length=10, num_beams=2)
        enc, mask = dec.create_encoder_sequence(feat)
        want = dec.bart.generate(encoder_outputs=BaseModelOutput(last_hidden_state=enc.contiguous()), attention_mask=mask,
                                 max_length=12, min_length=10, num_beams=2, early_stopping=True,
                                 decoder_start_token_id=dec.bart.config.decoder_start_token_id)
    finally:
        dec.bart.generation_config.no_repeat_ngram_size = 0
    assert torch.equal(out.cpu(), want.cpu())
    for row in out.tolist():
        grams = list(zip(row[1:], row[2:]))
        body = [g for g in grams if g[0] != 0 and g[1] != 0]
        assert len(body) == len(set(body))


def test_load_bart_raises_without_a_checkpoint(monkeypatch):
    """No silent random-weight fallback: a missing checkpoint is an error unless random init is requested."""
    from imagined_speech_translation_b200.model import _load_bart
    monkeypatch.setenv("EEGX_BART_RANDOM_INIT", "0")
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    with pytest.raises(Exception):
        _load_bart("fnlp/definitely-not-cached-bart")
    assert _load_bart("random").config.vocab_size == 51271

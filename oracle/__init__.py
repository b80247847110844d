"""CPU oracle for the EEG preprocessing + encoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product path (``imagined_speech_translation_b200``)
never imports this package and has no CPU fallback.

Parity status (see DESIGN.md "Oracle"):

* reference-actual normalisation (``process_raw_eeg`` / ``normalize_regions`` /
  ``zscore_time``): PINNED against the reference's own ``EEGDataset`` imported
  from ``/root/reference`` -> ``tests/golden/normalize_ref.npz``
  (generator: ``tests/golden/make_golden.py``).
* DSP stages (window -> FIR -> STFT -> log-power -> z-score): the reference has
  no such code, so for these stages PARITY IS UNPINNED
  by the reference; the oracle restates the written spec of SURVEY.md
  section 8(c) in fp64 and is itself pinned against ``scipy.signal.firwin`` +
  ``torch.nn.functional.conv1d`` + ``torch.stft`` (the calls the spec names)
  -> ``tests/golden/dsp_spec.npz``.
* encoder modules: pinned by golden tensors produced by the reference modules
  imported from ``/root/reference`` -> ``tests/golden/encoder_*.pt``.
"""

"""CPU oracle for the waveform -> speaker-embedding path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (torch fp32 / numpy) restatement of the
reference's algorithm for the hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker (or as the timed CPU baseline) -- never on the product path.  The product package
(``voiceprintrecognition-pytorch_b200``) never imports ``oracle`` and fails loudly when its
CUDA library is missing.

Pinning: the reference ships no tests / golden vectors, so the oracle is
pinned against outputs of the reference itself, generated in the authoring container by
``tests/golden/make_golden.py`` (imports /root/reference) and committed under ``tests/golden``.
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures on every run, and
``tests/test_oracle_vs_reference.py`` re-checks it live against /root/reference when present.

Reference citations are relative to /root/reference (``kaldi.py`` = torchaudio/compliance/kaldi.py,
``functional.py`` = torchaudio/functional/functional.py of torchaudio 2.11.0).
"""

"""CPU oracle for the MASR inference hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement of the reference's algorithm for the path named by BASELINE.json's
``north_star`` (fbank -> Conformer-family encoder -> CTC softmax -> greedy / prefix beam),
each function citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker / CPU baseline.  ``masr_b200`` (the product) never
imports it and fails loudly when its CUDA library is missing.

Pinning (see DESIGN.md "Oracle"):
* fbank, Conformer forward / forward_chunk, CTC softmax, greedy decode: PINNED — checked against
  the unmodified reference run in the build container through ``oracle/ref_shims.py`` (golden
  vectors under ``tests/golden`` made by ``tests/golden/make_golden.py``; the reference itself has
  no tests or golden vectors).
* CTC prefix beam search: PARITY UNPINNED — the reference delegates to the un-vendored, absent
  ``paddlespeech_ctcdecoders`` C++ library (masr/decoders/swig_wrapper.py:1); the restatement
  follows the algorithm's public definition and is self-checked only.
"""

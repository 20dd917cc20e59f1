"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the SIDEKIT speaker-verification inference hot path
(audio -> log-Mel/MFCC -> HalfResNet34/TDNN -> pooling -> embedding -> cosine /
PLDA / two-covariance / as-norm scoring).  Nothing under ``sidekit_b200/`` may
import this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or the timed CPU baseline -- never as the product path.

Parity status: the reference ships no tests or golden vectors,
so the restatement is pinned against the reference ITSELF, imported in the
build container by ``oracle/ref_import.py`` (``tests/test_oracle_vs_reference.py``
runs whenever /root/reference is present) and against fixtures generated from
it and committed under ``tests/golden/`` by ``oracle/make_golden.py``; the
scoring functions are additionally pinned by the seed-based known-answer
vectors of SURVEY.md Appendix B.

Modules
  extract_ref.py  torch-CPU fp32/fp64 restatement of Xtractor.forward
  scoring_ref.py  numpy restatement of iv_scoring / score_normalization
  ref_import.py   import of the real reference with stub modules + patches P1/P2
  make_golden.py  regenerates tests/golden/*.npz from the real reference
"""

"""oracle/ -- TEST INFRASTRUCTURE ONLY (not product code).

CPU restatement of the reference's FP8 fake-quantization hot path
(Qualcomm-AI-research/FP8-quantization @ ede19a9).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here, and there only as the
checker or as the timed CPU baseline -- never as part of the shipped path.
The product package ``fp8_quantization_b200`` must never import ``oracle``.

Parity pin: ``oracle.fp8_oracle`` is checked bit-for-bit against the *real*
reference (imported from /root/reference in the build container through
``oracle.reference_loader``) by ``tests/golden/make_golden.py``; the outputs of
that run are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` wherever the suite runs.
"""

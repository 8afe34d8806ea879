"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the hot path used as the *checker* for the CUDA product path.  Nothing in
``supersdr_b200/`` may import, call, link or execute anything in this package; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

Parity status:

* Tier P (``tier_p.py``): restates arithmetic that EXISTS in the reference
  (``utils_supersdr.py:333-348, 780-813, 879-898, 1044-1076, 1106-1148``).  PINNED: checked
  against the unmodified reference functions imported headless (``ref_import.py``) and against the
  committed fixtures ``tests/golden/tier_p_*.npz`` generated from them (``make_golden.py``).
* Tier U (``tier_u.py``, ``c/ssdr_oracle.c``): FFT / log-magnitude / demodulation.  This arithmetic
  is NOT in the reference (it runs on the remote KiwiSDR server, which is not vendored and has no
  pinned version).  PARITY UNPINNED: the spec is builder-defined; only the
  wire conventions the reference pins are anchored (byte = dBm + 255, mode pass-bands, AGC
  parameter names, 12 kHz / 512-sample frames, big-endian int16 I/Q).
"""

"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's Y-Net forecasting hot path
(vita-epfl/motion-style-transfer).  It exists so that the CUDA product path in
``motion_style_transfer_b200`` can be checked against something that follows the
reference line by line.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package never imports ``oracle`` and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * rasterisation, sampling, k-means, soft-argmax, CWS, network forward, ADE/FDE:
    pinned against the reference itself run in the build container
    (``oracle/gen_golden.py`` imports /root/reference unmodified and writes
    ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them).
  * ``loralib==0.1.1`` (un-vendored dependency of the reference,
    requirements.txt:11): restated in ``oracle/loralib_restatement.py`` from the
    published 0.1.1 algorithm -- PARITY UNPINNED for that module (no reference
    test or golden vector exists for it; only ``train.py:46-59 --init_check``
    pins "B = 0 is an exact no-op", which the tests cover).
"""

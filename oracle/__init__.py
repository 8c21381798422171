"""CPU oracle for the OCBO MTS hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy, the Gaussian-process arithmetic that the
reference (fusion-ml/OCBO) reaches through its un-vendored dependency
``dragonfly-opt==0.1.4`` (reference ``requirements.txt:3``), behind the call
sites in ``src/gp/gp_interface.py:67-138``.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path and Dragonfly itself is absent from this
image, so the arithmetic below is pinned only by
  * the one in-tree formula ``DragonflyGP.get_pt_relations``
    (``src/gp/gp_interface.py:112-119``),
  * closed-form n in {0, 1, 2} posteriors and scipy ``cho_factor/cho_solve``
    cross-checks (``tests/test_oracle.py``),
  * the reference's own, unmodified strategy classes run on top of this
    restatement (``tests/golden/make_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline
legs may import anything from here.  The product package ``ocbo_b200`` never
does; it fails loudly when its CUDA library is missing.
"""

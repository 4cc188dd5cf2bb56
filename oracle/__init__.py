"""CPU oracle for the SCOUTER forward hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker (or as the timed CPU baseline),
never as the thing shipped.  The product path (``scouter_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Parity status: **parity unpinned by the reference's own tests** -- the reference
(wbw520/scouter) ships no tests, golden vectors or fixtures.
The oracle is instead pinned against the *reference itself*, imported and run on
the CPU of the build container by ``oracle/make_golden.py`` (which commits the
resulting vectors under ``tests/golden/``), and re-checked by
``tests/test_oracle_golden.py`` on every run.
"""

"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference GP prior term.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker or as the
timed CPU baseline -- never as the thing shipped.  The product package
``gppvae_b200`` does not import this package and fails loudly when its CUDA
library is missing.
"""

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the WITW retrieval hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / CPU baseline.
"""

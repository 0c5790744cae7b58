"""CPU oracle for the Rydberg-CZ hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` may be imported by the product package
(``noisyquantumsimulator_b200``).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs use it, and only as
the checker or as the timed CPU baseline.
"""

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TurboAE hot path.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as a fallback for the CUDA path.
"""

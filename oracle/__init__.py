"""TEST INFRASTRUCTURE ONLY - CPU oracle for the DiffusionHandles activation-lifting / 3D-warp hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline - never as a fallback of the CUDA path.
"""

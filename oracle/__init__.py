"""CPU oracle for the RMA-Net chamfer / se(3)-warp hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it, and there only as the checker (or the timed CPU baseline), never as the thing shipped.
"""

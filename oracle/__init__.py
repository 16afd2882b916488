"""CPU oracle for the CARLE environment step (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and there only as the checker / the timed CPU baseline.
The product path (``carle_b200``) never imports this package.
"""

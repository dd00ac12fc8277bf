"""TEST INFRASTRUCTURE ONLY - CPU restatement of the FullSubNet enhancement hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker (or as the CPU arm being timed).
The product path (``fullsubnet_b200``) never imports this package and fails
loudly when its CUDA library is missing.
"""

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the WildlifeMapper tile-detection hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / CPU baseline -- never as the
thing measured or shipped.  The product path (``wildlifemapper_b200``) never imports
this package and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against the *reference itself*, imported from ``/root/reference`` in
the build container by ``tests/golden/make_golden.py``; the outputs are committed as
small fixtures under ``tests/golden/`` and re-checked on every CPU test run.
"""

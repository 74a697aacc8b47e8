"""ORACLE — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``pi-gan-thz_b200/``) never does.  Each function cites the
reference file:line it follows, and the package is pinned against outputs of the reference itself
(``tests/golden/*.npz``, produced by ``tools/make_golden.py`` importing ``/root/reference``).
"""

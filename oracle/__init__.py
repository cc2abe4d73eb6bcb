"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package
``bluesky-simulator>=1.0.7`` (reference ``pyproject.toml:10``), which is neither vendored under
/root/reference nor installable in this image (no network), and the reference ships no tests,
golden vectors or fixtures.  Everything tagged [UPSTREAM-RECALL] below is a restatement of the
*published* TUDelft-CNS-ATM/bluesky algorithm from memory, anchored on the reference's own call
sites (cited per function).  The in-tree parts (env reset / action / obs / reward logic,
``common/functions.py``) are restated from the reference files directly and cited file:line.

What pins exist (tests/test_oracle_*.py): closed-form geo/aero cases, the ISA ``vcas2tas`` table,
``creconfs`` -> ``detect`` round trips, and the distribution-level episode-length pins derived from
the reference's shipped training logs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``bluesky_gym_sasha_b200``) never does.
"""

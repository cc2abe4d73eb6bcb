"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package
``bluesky-simulator>=1.0.7`` (reference ``pyproject.toml:10``), which is neither vendored under
/root/reference nor installable in this image (no network), and the reference ships no tests,
golden vectors or fixtures.  Everything tagged [UPSTREAM-RECALL] below is a restatement of the
*published* TUDelft-CNS-ATM/bluesky algorithm from memory, anchored on the reference's own call
sites (cited per function).  The in-tree parts (env reset / action / obs / reward logic,
``common/functions.py``) are restated from the reference files directly and cited file:line.

PARTIAL PIN (tests/golden/, tests/test_golden.py): the in-tree half of the path IS pinned -- the
reference's own env files and ``common/functions.py`` are executed unmodified in the build container
(``oracle/bs_shim.py`` stands in for the absent ``bluesky`` package, backed by oracle/traffic.py) and
their reset / step outputs are committed as golden vectors; oracle/envs.py reproduces them to 1e-9 and
the CUDA path to the stated float32 tolerances.  The simulator core under ``bs.*`` stays unpinned.

Other pins (tests/test_oracle.py): closed-form geo/aero cases, the ISA ``vcas2tas`` table,
``creconfs`` -> ``detect`` round trips, and the distribution-level episode-length pins derived from
the reference's shipped training logs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``bluesky_gym_sasha_b200``) never does.
"""

"""TEST INFRASTRUCTURE -- the CPU oracle package.  NOT part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this.  ``shippingenv_b200`` never does (tests/test_no_oracle_in_product.py
enforces it).

``oracle.lib`` wraps ``oracle/shipenv_oracle.c`` (built by ``oracle/Makefile`` into
``oracle/_build/liboracle.so``); ``oracle.philox_shim`` / ``oracle.ref_harness`` drive the
real reference where a checkout exists.  Parity of the C oracle is pinned by
``tests/test_oracle_golden.py`` against fixtures the unmodified reference produced.
"""

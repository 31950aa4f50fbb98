"""oracle/ -- TEST INFRASTRUCTURE, not product code.

A CPU (torch fp32 / numpy) restatement of the reference's WACNN (`-m cnn`)
forward pass, pinned against golden vectors produced by the *unmodified*
reference (see tests/golden/make_golden.py, oracle/ref_shim.py).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import this package, and only as the
checker / CPU baseline.  The product (`resdsic_b200/`) never imports it and
fails loudly when its CUDA library is missing.
"""

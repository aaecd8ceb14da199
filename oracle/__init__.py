"""TEST INFRASTRUCTURE.  CPU restatements of the reference's hot-path algorithms (the oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import anything from this package, and only as the checker / the timed CPU baseline -- never on
the product path.  effocr_b200/ must not import it (tests/test_abi.py::test_product_never_imports_oracle enforces).
"""

"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the TAM-TR detection-head hot path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import anything from this package, and only as the
checker or the timed CPU baseline -- never as part of the product path (`tamtr_b200/` never imports it).
"""

"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's geometry hot path (honglianghe/CDNet).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this package, and only as the checker / the reported CPU baseline -- never as a product path.
`cdnet_b200` must never import it.
"""

"""oracle/ -- TEST INFRASTRUCTURE.  CPU restatements of the reference algorithms for the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker / reported baseline.  The product (cvpr2023-unidistill_b200/)
never imports it.
"""

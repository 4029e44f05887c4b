"""TEST INFRASTRUCTURE ONLY: CPU oracle for GPclust's collapsed-VB hot path.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
only.  The product package `gpclust_b200` never imports it and has no CPU fallback.
"""

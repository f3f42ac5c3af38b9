"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see the header of asc_oracle.c).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
from .capi import Oracle, build_oracle  # noqa: F401

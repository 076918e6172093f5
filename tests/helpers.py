"""Test helpers (thin re-export of the oracle bridge)."""
from oracle.bridge import oracle_dynamics, rel_err  # noqa: F401

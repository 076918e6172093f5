"""Test helpers (thin re-export of the oracle bridge)."""
from oracle.bridge import entry_err, oracle_dynamics, rel_err  # noqa: F401

"""CPU oracle (test infrastructure only; see stv_oracle.py)."""

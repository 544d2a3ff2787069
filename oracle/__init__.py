"""CPU oracle (test infrastructure only -- see oracle/sa_oracle.c header)."""

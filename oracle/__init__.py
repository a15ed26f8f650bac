"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py, oracle/dbspm_oracle.c)."""

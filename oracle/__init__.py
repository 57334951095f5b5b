"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/gw_oracle.c header)."""

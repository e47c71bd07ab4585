"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/emrifd_oracle.c header)."""

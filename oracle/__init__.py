"""CPU oracle -- test infrastructure only (see oracle/vdl_oracle.c header). Parity unpinned."""

"""CPU oracle: test infrastructure only (see oracle/bn254.py, oracle/halo2_cpu.c headers)."""

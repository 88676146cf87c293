"""Test infrastructure: CPU oracle for the hot path (see palantir_oracle.py)."""

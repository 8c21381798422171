"""Measurement and profiling helpers used by bench.py and from the command line."""

"""Parity and host-logic tests (pytest -m gpu / -m "not gpu")."""

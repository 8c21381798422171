"""Continuous-strategy registry in the reference's format
(src/cstrats/__init__.py:12-15)."""
from .profile_cts import prof_strats

cstrats = list(prof_strats)

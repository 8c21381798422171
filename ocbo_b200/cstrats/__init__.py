"""Continuous-strategy registry in the reference's format
(src/cstrats/__init__.py:12-15)."""
from argparse import Namespace

from .baselines import AgnEI, AgnTS, RandOpt
from .postmax_cts import REVI
from .profile_cts import prof_strats

cstrats = list(prof_strats) + [Namespace(impl=c, name=c.get_strat_name())
                               for c in (RandOpt, AgnEI, AgnTS, REVI)]

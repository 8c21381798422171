"""Strategy registry in the reference's format (src/strategies/__init__.py:13-27):
``Namespace(impl=cls, name=str)`` entries matched case-insensitively by the
drivers (src/ocbo.py:176-177).  The MTS family is the accelerated hot path
(SURVEY 8a); MEI / joint-MEI are the first "next" row (8f, f2) and share its kernels."""
from argparse import Namespace

from .corr_opt import REVI
from .baselines import AgnEI, AgnThompson, JointAgnEI, JointAgnThompson, JointRandom, RandomOpt
from .joint_mei import JointMEI
from .joint_mts import JointMTS
from .mei import MEI
from .mts import MTS

strategies = [
    Namespace(impl=MTS, name=MTS.get_opt_method_name()),
    Namespace(impl=JointMTS, name=JointMTS.get_opt_method_name()),
    Namespace(impl=MEI, name=MEI.get_opt_method_name()),
    Namespace(impl=JointMEI, name=JointMEI.get_opt_method_name()),
    # baselines of the option files' method lists (not on the hot path, same device entry points)
    Namespace(impl=AgnEI, name=AgnEI.get_opt_method_name()),
    Namespace(impl=AgnThompson, name=AgnThompson.get_opt_method_name()),
    Namespace(impl=JointAgnThompson, name=JointAgnThompson.get_opt_method_name()),
    Namespace(impl=JointAgnEI, name=JointAgnEI.get_opt_method_name()),
    Namespace(impl=RandomOpt, name=RandomOpt.get_opt_method_name()),
    Namespace(impl=JointRandom, name=JointRandom.get_opt_method_name()),
    # REVI: knowledge gradient over the joint GP
    Namespace(impl=REVI, name=REVI.get_opt_method_name()),
]

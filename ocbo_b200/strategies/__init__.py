"""Strategy registry in the reference's format (src/strategies/__init__.py:13-27):
``Namespace(impl=cls, name=str)`` entries matched case-insensitively by the
drivers (src/ocbo.py:176-177).  Only the MTS family is on the accelerated hot
path."""
from argparse import Namespace

from .joint_mts import JointMTS
from .mts import MTS

strategies = [
    Namespace(impl=MTS, name=MTS.get_opt_method_name()),
    Namespace(impl=JointMTS, name=JointMTS.get_opt_method_name()),
]

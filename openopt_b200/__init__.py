"""openopt_b200 -- B200-native backend for OpenOpt's differential-evolution path.

    from openopt_b200 import GLP, objectives
    p = GLP(objectives.Rastrigin(), x0, lb=lb, ub=ub)
    r = p.solve('de', population=10000)

See DESIGN.md.  Only the `p.solve('de')` path of OpenOpt 0.43 is provided.
"""
__version__ = '0.1'

from . import objectives  # noqa: F401
from .host import GLP, NLP, OpenOptException, oosolver  # noqa: F401

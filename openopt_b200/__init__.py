"""openopt_b200 -- B200-native backend for OpenOpt's differential-evolution path.

    from openopt_b200 import GLP, objectives
    p = GLP(objectives.Rastrigin(), x0, lb=lb, ub=ub)
    r = p.solve('de', population=10000)

See DESIGN.md.  Only the `p.solve('de')` path of OpenOpt 0.43 is provided.
"""
__version__ = '0.1'
OPENOPT_API_VERSION = '0.43'      # the OpenOpt release whose solve('de') path this mirrors (shown in the log banner)

from . import objectives  # noqa: F401
from .host import GLP, NLP, OpenOptException, oosolver  # noqa: F401

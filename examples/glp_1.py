"""OpenOpt's examples/glp_1.py on the CUDA `de` backend.

The reference example (openopt/examples/glp_1.py:4-10) is

    f = lambda x: (x[0]-1.5)**2 + sin(0.8 * x[1] ** 2 + 15)**4 + cos(0.8 * x[2] ** 2 + 15)**4 + (x[3]-7.5)**4
    p = GLP(f, lb = -ones(4),  ub = ones(4),  maxIter = 1e3,  maxFunEvals = 1e5,  maxTime = 3,  maxCPUTime = 3)
    r = p.solve('de', plot=1)

Here the objective is the same formula written as an expression (openopt_b200.expr): the host evaluates it with
numpy exactly as the lambda does, the GPU runs the generated device twin.  Everything else is the reference's call
(plotting is not part of this backend).  `python examples/glp_1.py` needs a CUDA device.
"""
import os
import sys

from numpy import cos, ones, sin

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openopt_b200 import GLP, expr as E, objectives  # noqa: E402

# the reference's objective, verbatim (used by the oracle and the parity tests)
f_ref = lambda x: (x[0]-1.5)**2 + sin(0.8 * x[1] ** 2 + 15)**4 + cos(0.8 * x[2] ** 2 + 15)**4 + (x[3]-7.5)**4

# the same formula as a device objective: one term per column, summed over the row
x = E.X
f = E.ExprObjective(E.by_column([(x - 1.5) ** 2, E.sin(0.8 * x ** 2 + 15) ** 4, E.cos(0.8 * x ** 2 + 15) ** 4, (x - 7.5) ** 4]),
                    name='glp1')
assert f.key() == objectives.get('glp1').key()      # the same objective the test suite pins against the reference


def problem(**kw):
    args = dict(lb=-ones(4), ub=ones(4), maxIter=1e3, maxFunEvals=1e5, maxTime=3, maxCPUTime=3)
    args.update(kw)
    return GLP(f, **args)


if __name__ == '__main__':
    p = problem()
    r = p.solve('de')
    x_opt, f_opt = r.xf, r.ff
    print('x_opt =', x_opt, ' f_opt =', f_opt)

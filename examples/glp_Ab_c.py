"""OpenOpt's examples/glp_Ab_c.py on the CUDA `de` backend: the glp_1 objective with linear inequality constraints
A x <= b and two nonlinear constraints c(x) <= 0 (openopt/examples/glp_Ab_c.py:6-33).

    c = lambda x: (x[0] ** 2 + x[2] ** 2 - 0.15,  1.5 * x[0] ** 2 + x[1] ** 2 - 1.5)

is a diagonal quadratic form, i.e. the registered device constraint DiagQuadratic([[1,0,1,0],[1.5,1,0,0]], [0.15,1.5]).
The reference passes `mutationRate = 0.15` to solve(), which `de` does not know (it is reported and ignored there,
kernel/runProbSolver.py:99-104); it is left out here.
"""
import os
import sys

from numpy import array, ones

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glp_1 import f, f_ref  # noqa: E402,F401
from openopt_b200 import GLP  # noqa: E402
from openopt_b200.constraints import DiagQuadratic  # noqa: E402

lb, ub = -ones(4), ones(4)
A = array([[1, 0, 0, 1], [0, 1, 0, 1]])        # the reference writes mat('1 0 0 1; 0 1 0 1') (numpy.mat is gone in numpy 2)
b = [0.15, 1.5]
c_ref = lambda x: (x[0] ** 2 + x[2] ** 2 - 0.15,  1.5 * x[0] ** 2 + x[1] ** 2 - 1.5)
c = DiagQuadratic([[1, 0, 1, 0], [1.5, 1, 0, 0]], [0.15, 1.5])


def problem(**kw):
    args = dict(lb=lb, ub=ub, A=A, b=b, c=c, maxIter=250, maxFunEvals=1e5, maxTime=30, maxCPUTime=30)
    args.update(kw)
    return GLP(f, **args)


if __name__ == '__main__':
    p = problem()
    r = p.solve('de')
    x_opt, f_opt = r.xf, r.ff
    print('x_opt =', x_opt, ' f_opt =', f_opt, ' feasible:', r.isFeasible)

"""Registered device-side objectives.

The reference calls an arbitrary user callable ``f`` one row at a time
(kernel/nonLinFuncs.py:190-200).  The CUDA backend needs an objective it can evaluate on the
device, so ``f`` must be one of the objects below: each is an ordinary host callable (numpy, one
row or a 2-D array of rows -- this numpy form is the parity definition; the
driver still calls it on the host for F(x0), the GLP maxNonSuccess test and the final ff,
kernel/runProbSolver.py:259,326, kernel/GLP.py:24-28) and carries the id / per-column data of the
matching ``Obj<>`` specialisation in csrc/de_device.cuh.
"""
import math

import numpy as np

from . import _lib

TWO_PI = 2 * math.pi


class DeviceObjective:
    """Base class: a host callable with a device twin."""
    device_id = None
    name = None

    def aux(self, D):
        """Per-column data the device kernel needs ([D] float64) or None."""
        return None

    def __call__(self, x):
        raise NotImplementedError

    def __repr__(self):
        return '<device objective %s>' % self.name


class Sphere(DeviceObjective):
    device_id, name = _lib.OBJ_SPHERE, 'sphere'

    def __call__(self, x):
        x = np.asarray(x)
        return (x * x).sum(-1)


class ShiftedSphere(DeviceObjective):
    """((x - shift)**2).sum()  -- examples/glp_3.py:6 uses shift = arange(N)."""
    device_id, name = _lib.OBJ_SHIFTED_SPHERE, 'shifted_sphere'

    def __init__(self, shift):
        self.shift = np.asarray(shift, dtype=np.float64)

    def aux(self, D):
        return np.broadcast_to(self.shift, (D,)).copy()

    def __call__(self, x):
        return ((np.asarray(x) - self.shift) ** 2).sum(-1)


class Rosenbrock(DeviceObjective):
    device_id, name = _lib.OBJ_ROSENBROCK, 'rosenbrock'

    def __call__(self, x):
        x = np.asarray(x)
        return (100.0 * (x[..., 1:] - x[..., :-1] ** 2) ** 2 + (1.0 - x[..., :-1]) ** 2).sum(-1)


class Rastrigin(DeviceObjective):
    """openopt/tests/rastrigin.py:4."""
    device_id, name = _lib.OBJ_RASTRIGIN, 'rastrigin'

    def __call__(self, x):
        x = np.asarray(x)
        return (x * x - 10 * np.cos(TWO_PI * x) + 10).sum(-1)


class Ackley(DeviceObjective):
    device_id, name = _lib.OBJ_ACKLEY, 'ackley'

    def __call__(self, x):
        x = np.asarray(x)
        D = x.shape[-1]
        return (-20.0 * np.exp(-0.2 * np.sqrt((x * x).sum(-1) / D))
                - np.exp(np.cos(TWO_PI * x).sum(-1) / D) + 20.0 + math.e)


class Griewank(DeviceObjective):
    device_id, name = _lib.OBJ_GRIEWANK, 'griewank'

    def aux(self, D):
        return np.sqrt(np.arange(1, D + 1))

    def __call__(self, x):
        x = np.asarray(x)
        D = x.shape[-1]
        return 1.0 + (x * x).sum(-1) / 4000.0 - np.cos(x / np.sqrt(np.arange(1, D + 1))).prod(-1)


BY_NAME = {'sphere': Sphere, 'rosenbrock': Rosenbrock, 'rastrigin': Rastrigin, 'ackley': Ackley,
           'griewank': Griewank}


def _glp1():
    """examples/glp_1.py:4 of the reference as an expression objective (openopt_b200/expr.py): one term per column."""
    from . import expr as E
    x = E.X
    o = E.ExprObjective(E.by_column([(x - 1.5) ** 2, E.sin(0.8 * x ** 2 + 15) ** 4, E.cos(0.8 * x ** 2 + 15) ** 4,
                                     (x - 7.5) ** 4]), name='glp1')
    return o


BY_NAME['glp1'] = _glp1


def get(name):
    return BY_NAME[name]()

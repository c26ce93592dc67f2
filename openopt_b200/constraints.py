"""Registered device-side nonlinear constraints.

The reference evaluates the user's ``c`` (c(x) <= 0) and ``h`` (h(x) = 0) one row at a time through
kernel/nonLinFuncs.py:190-200, and de's _eval_pop folds them into one value per row
(de_oo.py:314-316; kernel/Point.py:128-167).  Like objectives, a constraint function that is to run
on the GPU must be one of the objects below: an ordinary host callable (the numpy form is the parity
definition and is what the host driver evaluates at single points) that also carries the data of its
device twin (``de_constr_kernel`` in csrc/de_kernels.cuh).  Linear constraints (A, b, Aeq, beq) need
no registration.
"""
import numpy as np


class DeviceConstraint:
    """Base class: a host callable with a device twin."""
    kind = None


class DiagQuadratic(DeviceConstraint):
    """m constraints  g_k(x) = sum_j Q[k, j] * x[j]**2 - r[k]   (use as c: g_k(x) <= 0, or h: g_k(x) = 0).

    examples/glp_Ab_c.py's  c = (x0**2 + x2**2 - 0.15, 1.5*x0**2 + x1**2 - 1.5)  is
    DiagQuadratic([[1, 0, 1, 0], [1.5, 1, 0, 0]], [0.15, 1.5])."""
    kind = 'diag_quadratic'

    def __init__(self, Q, r):
        self.Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
        self.r = np.asarray(r, dtype=np.float64).ravel()
        assert self.Q.shape[0] == self.r.size

    @property
    def m(self):
        return self.r.size

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        return (self.Q * (x * x)[..., None, :]).sum(-1) - self.r

    def __repr__(self):
        return '<device constraint diag_quadratic, %d rows>' % self.m

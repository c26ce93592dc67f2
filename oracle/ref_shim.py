"""Loader for the UNMODIFIED reference (OpenOpt 0.43 at /root/reference) -- test infrastructure only.

The reference is pure Python + numpy but predates numpy 2 / Python 3.12
(``numpy.asfarray``/``asscalar`` and ``time.clock`` are gone), so it is loaded behind a
three-item runtime shim; nothing under /root/reference is edited or copied.

Only ``tests/golden/make_golden.py`` (run in the build container, where /root/reference is
mounted) uses this file.  /root/reference does not exist on the GPU box: nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` imports this module.
"""
import sys
import time

import numpy

REFERENCE_ROOT = '/root/reference'


def load_reference(root=REFERENCE_ROOT):
    """Return (openopt module, de_oo module) of the unmodified reference."""
    if not hasattr(numpy, 'asfarray'):
        numpy.asfarray = lambda a, dtype=numpy.float64: numpy.asarray(a, dtype=dtype)
        numpy.__all__.append('asfarray')       # the reference uses `from numpy import *`
    if not hasattr(numpy, 'asscalar'):
        numpy.asscalar = lambda a: numpy.asarray(a).item()
        numpy.__all__.append('asscalar')
    if not hasattr(time, 'clock'):
        time.clock = time.process_time
    sys.dont_write_bytecode = True             # /root/reference is read-only
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import openopt
        from openopt.solvers.UkrOpt import de_oo
    return openopt, de_oo

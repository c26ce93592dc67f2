"""Loader for the UNMODIFIED reference `galileo` solver (openopt/solvers/Standalone/galileo_oo.py) -- test
infrastructure only, used by tests/golden/make_golden_galileo.py and tests/test_galileo_vs_reference.py in the build
container (where /root/reference is mounted).

galileo_oo.py is Python-2 code.  Three Python-2 behaviours are restored at run time, on the loaded module's classes
(nothing under /root/reference is edited or copied):
  1. ordering through `__cmp__` (:188-191): Python 3's `list.sort()` only looks for `__lt__`;
     `Chromosome.__lt__ = lambda a, b: a.__cmp__(b) < 0` is what Python 2's rich-comparison fallback evaluated.
  2. simple slices through `__getslice__` (:169-176): Python 3 passes a slice object to `__getitem__` (:158-164).
  3. a seed: `Population.__init__` (:223) builds an unseeded `Random()`; the module-level name `Random` is replaced
     by a factory returning `random.Random(seed)` so that runs can be recorded and replayed.
"""
import random
import warnings

from ref_shim import load_reference


def load_galileo(seed, root=None):
    """Return (openopt module, galileo_oo module) with the generator seeded by `seed`.  root: where the unmodified
    reference lies (default /root/reference; bench.py's cpu_baseline leg passes baseline/_ref)."""
    oo, _ = load_reference(root) if root else load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from openopt.solvers.Standalone import galileo_oo as g
    C = g.Chromosome
    if not getattr(C, '_py2_semantics_restored', False):
        item = C.__getitem__

        def getitem(self, key):
            if isinstance(key, slice):
                i, j, _ = key.indices(len(self.genes))
                return self.__getslice__(i, j)
            return item(self, key)

        C.__getitem__ = getitem
        C.__lt__ = lambda a, b: a.__cmp__(b) < 0
        C._py2_semantics_restored = True
    g.Random = lambda: random.Random(seed)
    return oo, g

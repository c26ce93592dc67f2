"""Open objective registry: objectives written as expressions, compiled into device code at run time.

The reference accepts any Python callable ``f`` and calls it one row at a time (kernel/nonLinFuncs.py:190-200;
e.g. examples/glp_1.py:4).  On the GPU an objective is device code -- a specialisation of the kernels' ``Obj<>``
template (csrc/de_device.cuh) -- so an arbitrary callable cannot run there.  This module closes the gap for every
objective of the form

    f(x) = finalize( reduce_0(t_0(x_j, j)), reduce_1(t_1(x_j, j)), ..., D )

i.e. per-gene terms (each may depend on the gene, its column index and one per-column parameter array), reduced over
the row by sums and/or products, and a closing scalar expression -- which covers the separable test functions and
per-column formulas such as examples/glp_1.py's.  The expression is written once with ordinary Python operators:

    from openopt_b200 import expr as E
    rastrigin = E.ExprObjective(E.X * E.X - 10 * E.cos(2 * math.pi * E.X) + 10)
    glp1 = E.ExprObjective(E.by_column([(E.X - 1.5) ** 2, E.sin(0.8 * E.X ** 2 + 15) ** 4,
                                        E.cos(0.8 * E.X ** 2 + 15) ** 4, (E.X - 7.5) ** 4]))

and yields BOTH forms of a registered device objective (objectives.DeviceObjective):
  * the host callable -- evaluated with numpy, operation for operation as written (numpy's ``.sum()`` pairwise order,
    ``**2`` as a square, ...): this is the parity definition the oracle uses, and what the driver calls at single points;
  * the device twin -- generated CUDA (one IEEE operation per node, never contracted; the kernels reproduce numpy's
    pairwise summation order), compiled with nvcc against csrc/trial_obj.cu into its own shared object (cached under
    openopt_b200/_objcache/ by content hash) and handed to liboode with oode_register_objective.
Elementary functions (sin, cos, exp, log, pow, ...) are CUDA's on the device and numpy's on the host: they agree to
1-2 ulp, inside the 1e-12 relative parity bar; everything built from + - * / sqrt is bit-exact.
"""
import hashlib
import math
import os
import subprocess

import numpy as np

from . import _lib
from .objectives import DeviceObjective

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.path.join(HERE, '_objcache')
CUSTOM_ID = 6          # Obj<6> inside the separately compiled translation unit (de_launch.h: OODE_OBJ_CUSTOM)


# --------------------------------------------------------------------------------------------
# expression tree
# --------------------------------------------------------------------------------------------
def _wrap(v):
    if isinstance(v, Expr):
        return v
    if isinstance(v, (int, float, np.integer, np.floating)):
        return Const(float(v))
    raise TypeError('cannot use %r in an objective expression' % (v,))


class Expr:
    def __add__(self, o): return Bin('+', self, _wrap(o))
    def __radd__(self, o): return Bin('+', _wrap(o), self)
    def __sub__(self, o): return Bin('-', self, _wrap(o))
    def __rsub__(self, o): return Bin('-', _wrap(o), self)
    def __mul__(self, o): return Bin('*', self, _wrap(o))
    def __rmul__(self, o): return Bin('*', _wrap(o), self)
    def __truediv__(self, o): return Bin('/', self, _wrap(o))
    def __rtruediv__(self, o): return Bin('/', _wrap(o), self)
    def __neg__(self): return Un('neg', self)
    def __pos__(self): return self
    def __pow__(self, o): return Pow(self, _wrap(o))
    def __rpow__(self, o): return Pow(_wrap(o), self)
    def __abs__(self): return Un('abs', self)
    def __lt__(self, o): return Cmp('<', self, _wrap(o))
    def __le__(self, o): return Cmp('<=', self, _wrap(o))
    def __gt__(self, o): return Cmp('>', self, _wrap(o))
    def __ge__(self, o): return Cmp('>=', self, _wrap(o))


class Const(Expr):
    def __init__(self, v): self.v = float(v)
    def np(self, env): return self.v
    def cu(self):
        if math.isnan(self.v):
            return '__longlong_as_double(0x7FF8000000000000LL)'
        if math.isinf(self.v):
            return ('-' if self.v < 0 else '') + '__longlong_as_double(0x7FF0000000000000LL)'
        return '(%s)' % float(self.v).hex()          # hexadecimal literal: exactly this double


class Var(Expr):
    CU = {'x': 'x', 'j': '((double)col)', 'a': 'aux', 'D': '((double)D)'}

    def __init__(self, name): self.name = name
    def np(self, env): return env[self.name]
    def cu(self): return self.CU[self.name] if self.name in self.CU else 's[%d]' % int(self.name[1:])


class Bin(Expr):
    CU = {'+': 'dadd', '-': 'dsub', '*': 'dmul', '/': 'ddiv'}

    def __init__(self, op, a, b): self.op, self.a, self.b = op, a, b

    def np(self, env):
        a, b = self.a.np(env), self.b.np(env)
        return a + b if self.op == '+' else a - b if self.op == '-' else a * b if self.op == '*' else a / b

    def cu(self): return '%s(%s, %s)' % (self.CU[self.op], self.a.cu(), self.b.cu())


class Un(Expr):
    NP = {'neg': np.negative, 'abs': np.abs, 'sin': np.sin, 'cos': np.cos, 'tan': np.tan, 'exp': np.exp, 'log': np.log,
          'sqrt': np.sqrt, 'tanh': np.tanh, 'floor': np.floor, 'ceil': np.ceil}
    CU = {'neg': '(-(%s))', 'abs': 'fabs(%s)', 'sin': 'sin(%s)', 'cos': 'cos(%s)', 'tan': 'tan(%s)', 'exp': 'exp(%s)',
          'log': 'log(%s)', 'sqrt': '__dsqrt_rn(%s)', 'tanh': 'tanh(%s)', 'floor': 'floor(%s)', 'ceil': 'ceil(%s)'}

    def __init__(self, fn, a): self.fn, self.a = fn, a
    def np(self, env): return self.NP[self.fn](self.a.np(env))
    def cu(self): return self.CU[self.fn] % self.a.cu()


class Pow(Expr):
    """a ** b with numpy's fast paths for constant exponents (2: square, 0.5: sqrt, 1, 0, -1), otherwise pow()."""

    def __init__(self, a, b): self.a, self.b = a, b

    def np(self, env):
        return self.a.np(env) ** self.b.np(env)

    def cu(self):
        a = self.a.cu()
        if isinstance(self.b, Const):
            e = self.b.v
            if e == 2.0:
                return '_sq(%s)' % a
            if e == 0.5:
                return '__dsqrt_rn(%s)' % a
            if e == 1.0:
                return a
            if e == -1.0:
                return 'ddiv(1.0, %s)' % a
        return 'pow(%s, %s)' % (a, self.b.cu())


class Cmp:
    def __init__(self, op, a, b): self.op, self.a, self.b = op, a, b

    def np(self, env):
        a, b = self.a.np(env), self.b.np(env)
        return {'<': np.less, '<=': np.less_equal, '>': np.greater, '>=': np.greater_equal}[self.op](a, b)

    def cu(self): return '(%s %s %s)' % (self.a.cu(), self.op, self.b.cu())


class Where(Expr):
    def __init__(self, c, a, b): self.c, self.a, self.b = c, _wrap(a), _wrap(b)
    def np(self, env): return np.where(self.c.np(env), self.a.np(env), self.b.np(env))
    def cu(self): return '(%s ? %s : %s)' % (self.c.cu(), self.a.cu(), self.b.cu())


class ByColumn(Expr):
    """A different term expression per column: a list (one entry per column) or {column: expr} with a default."""

    def __init__(self, cols, default=None):
        self.cols = dict(enumerate(cols)) if isinstance(cols, (list, tuple)) else dict(cols)
        self.cols = {int(k): _wrap(v) for k, v in self.cols.items()}
        self.default = None if default is None else _wrap(default)
        self.exact = isinstance(cols, (list, tuple))        # exactly len(cols) columns

    def np(self, env):
        x = env['x']
        out = np.empty_like(np.asarray(x, dtype=np.float64))
        D = out.shape[-1]
        if self.default is None and any(k not in self.cols for k in range(D)):
            raise ValueError('by_column: no expression for some of the %d columns' % D)
        for k in range(D):
            e = self.cols.get(k, self.default)
            sub = dict(env, x=x[..., k], j=float(k))
            if 'a' in env and env['a'] is not None:
                sub['a'] = env['a'][..., k]
            out[..., k] = e.np(sub)
        return out

    def cu(self):
        raise TypeError('by_column can only be a whole term, not part of a larger expression')

    def cu_statement(self, target):
        lines = ['switch (col) {']
        for k in sorted(self.cols):
            lines.append('    case %d: %s = %s; break;' % (k, target, self.cols[k].cu()))
        lines.append('    default: %s = %s; break;' % (target, self.default.cu() if self.default is not None else '0.0'))
        lines.append('}')
        return '\n        '.join(lines)


X, J, A, DIM = Var('x'), Var('j'), Var('a'), Var('D')       # the gene, its column index, the per-column parameter, D


def S(k):
    """The k-th reduced slot inside `finalize` (sum slots first, then product slots)."""
    return Var('S%d' % k)


def _fn(name):
    return lambda a: Un(name, _wrap(a))


sin, cos, tan, exp, log, sqrt, tanh, floor, ceil = (_fn(n) for n in ('sin', 'cos', 'tan', 'exp', 'log', 'sqrt', 'tanh', 'floor', 'ceil'))


def where(cond, a, b):
    return Where(cond, a, b)


def by_column(cols, default=None):
    return ByColumn(cols, default)


# --------------------------------------------------------------------------------------------
# the objective
# --------------------------------------------------------------------------------------------
class ExprObjective(DeviceObjective):
    """f(x) = finalize(S0, S1, ..., D) with S_k = sum (or product) over the row of term_k(x_j, j, a_j).

    terms     one expression, or a list of expressions / ('sum' | 'prod', expression) pairs (sums first)
    finalize  expression over S(0), S(1), ... and DIM; default S(0)
    aux       per-column parameter: an array of D values or a callable D -> array (the `A` symbol)
    """

    def __init__(self, terms, finalize=None, aux=None, name=None):
        if isinstance(terms, Expr):
            terms = [terms]
        norm = [(t if isinstance(t, tuple) else ('sum', t)) for t in terms]
        self.sums = [_wrap(e) for k, e in norm if k == 'sum']
        self.prods = [_wrap(e) for k, e in norm if k == 'prod']
        if len(self.sums) + len(self.prods) != len(norm) or not norm:
            raise ValueError("terms: expressions or ('sum' | 'prod', expression) pairs")
        if not self.sums:
            raise ValueError('at least one sum slot is required (the kernels order sum slots first)')
        if len(norm) > 2:
            raise ValueError('at most two reduction slots per row are supported by the kernels')
        self.finalize = _wrap(finalize) if finalize is not None else S(0)
        self._aux = aux
        self.name = name or 'expr_' + self.key()[:10]
        self._device_id = None

    # ---- host form (the parity definition) ----
    def aux(self, D):
        if self._aux is None:
            return None
        a = self._aux(D) if callable(self._aux) else self._aux
        return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (D,)))

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        D = x.shape[-1]
        env = {'x': x, 'j': np.arange(D, dtype=np.float64), 'a': self.aux(D), 'D': float(D)}
        red = {}
        with np.errstate(all='ignore'):
            for k, e in enumerate(self.sums):
                t = np.broadcast_to(e.np(env), x.shape)
                red['S%d' % k] = np.ascontiguousarray(t).sum(-1)
            for k, e in enumerate(self.prods):
                t = np.broadcast_to(e.np(env), x.shape)
                red['S%d' % (len(self.sums) + k)] = np.ascontiguousarray(t).prod(-1)
            return self.finalize.np(dict(red, D=float(D)))

    # ---- device form ----
    def header(self):
        body = []
        for k, e in enumerate(self.sums + self.prods):
            if isinstance(e, ByColumn):
                body.append(e.cu_statement('t[%d]' % k))
            else:
                body.append('t[%d] = %s;' % (k, e.cu()))
        return ('// generated by openopt_b200/expr.py -- do not edit\n'
                '__device__ __forceinline__ double _sq(double v) { return dmul(v, v); }\n'
                'template <> struct Obj<%d> {\n'
                '    static constexpr int NS = %d, NPR = %d, HALO = 0;\n'
                '    static constexpr bool AUX = %s;\n'
                '    template <bool CK = true> __device__ static __forceinline__ void term(double x, double, double aux, int col, double *t) {\n'
                '        (void)x; (void)aux; (void)col;\n'
                '        %s\n'
                '    }\n'
                '    __device__ static __forceinline__ double finalize(const double *s, int D) { (void)D; return %s; }\n'
                '};\n') % (CUSTOM_ID, len(self.sums), len(self.prods), 'true' if self._aux is not None else 'false',
                           '\n        '.join(body), self.finalize.cu())

    def key(self):
        return hashlib.sha256(self.header().encode()).hexdigest()

    def build(self, verbose=False):
        """Compile the device twin (cached by content hash of the generated code, the kernel sources and the flags)."""
        from . import build as b
        h = hashlib.sha256()
        h.update(self.header().encode())
        for f in sorted(os.listdir(b.CSRC)):
            if f.endswith(('.cu', '.cuh', '.h')):
                h.update(open(os.path.join(b.CSRC, f), 'rb').read())
        h.update(' '.join(b.NVCC_FLAGS).encode())
        tag = h.hexdigest()[:20]
        os.makedirs(CACHE, exist_ok=True)
        lib = os.path.join(CACHE, 'obj_%s.so' % tag)
        if not os.path.exists(lib):
            hdr = os.path.join(CACHE, 'obj_%s.cuh' % tag)
            with open(hdr, 'w') as fh:
                fh.write(self.header())
            cmd = [b._nvcc()] + b.NVCC_FLAGS + ['-shared', '-DOODE_OBJ=%d' % CUSTOM_ID, '-DOODE_CUSTOM_HEADER="%s"' % hdr,
                                                os.path.join(b.CSRC, 'trial_obj.cu'), '-o', lib + '.tmp']
            if verbose:
                print(' '.join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise _lib.LibraryError('nvcc failed for objective %s:\n%s' % (self.name, (r.stdout + r.stderr)[-6000:]))
            os.replace(lib + '.tmp', lib)
        return lib

    @property
    def device_id(self):
        if self._device_id is None:
            self._device_id = _lib.register_objective(self.build())
        return self._device_id

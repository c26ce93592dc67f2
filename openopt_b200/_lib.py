"""ctypes binding of liboode.so (include/oode.h) -- the only way host Python reaches the kernels.

There is no CPU fallback: if the library is missing or no CUDA device is present the calls fail
loudly (OpenOptException from the plugin, RuntimeError here).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('OODE_LIB_PATH') or os.path.join(HERE, 'liboode.so')     # override: A/B timing of two builds

OODE_ABI_VERSION = 3
OBJ_SPHERE, OBJ_ROSENBROCK, OBJ_RASTRIGIN, OBJ_ACKLEY, OBJ_GRIEWANK, OBJ_SHIFTED_SPHERE = range(6)
RNG_PHILOX, RNG_REPLAY, RNG_CPYTHON = range(3)
SHARD_NONE, SHARD_ROWS, SHARD_COLS = range(3)
XCHG_NONE, XCHG_NCCL, XCHG_PEER = range(3)
XCHG_NAMES = {XCHG_NONE: 'none', XCHG_NCCL: 'nccl all-gather', XCHG_PEER: 'peer stores over NVLink'}
RNG_BY_NAME = {'philox': RNG_PHILOX, 'replay': RNG_REPLAY, 'cpython': RNG_CPYTHON}
BASE_BY_NAME = {'random': 0, 'best': 1}
DIRECTION_BY_NAME = {'random': 0, 'best': 1}
FACTOR_BY_NAME = {'random': 0, 'constant': 1}
MAX_HNDVI = 512

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('dim', C.c_int32), ('population', C.c_int64),
                ('lb', _dp), ('ub', _dp), ('x0', _dp),
                ('diff_factor', C.c_double), ('cross_rate', C.c_double), ('seed', C.c_uint64),
                ('objective', C.c_int32), ('invert', C.c_int32), ('obj_aux', _dp),
                ('rng', C.c_int32), ('device', C.c_int32), ('world_size', C.c_int32), ('rank', C.c_int32),
                ('shard', C.c_int32), ('max_batch', C.c_int32), ('nccl_id', C.c_void_p),
                ('base_strategy', C.c_int32), ('direction_strategy', C.c_int32), ('factor_strategy', C.c_int32),
                ('hndvi', C.c_int32)]


class GenRecord(C.Structure):
    _fields_ = [('trial_best_f', C.c_double), ('best_f', C.c_double), ('trial_best_idx', C.c_int64),
                ('n_accepted', C.c_int64), ('n_repaired', C.c_int64), ('best_changed', C.c_int32),
                ('generation', C.c_int32), ('trial_best_c', C.c_double), ('trial_best_feasible', C.c_int32),
                ('reserved', C.c_int32)]


class Constraints(C.Structure):
    _fields_ = [('n_lin_ineq', C.c_int32), ('n_lin_eq', C.c_int32), ('n_quad_ineq', C.c_int32), ('n_quad_eq', C.c_int32),
                ('A', _dp), ('b', _dp), ('Aeq', _dp), ('beq', _dp), ('Qc', _dp), ('rc', _dp), ('Qh', _dp), ('rh', _dp),
                ('contol', C.c_double)]


class GaConfig(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('dim', C.c_int32), ('population', C.c_int64), ('lb', _dp), ('ub', _dp),
                ('crossover_rate', C.c_double), ('mutation_rate', C.c_double), ('seed', C.c_uint64),
                ('objective', C.c_int32), ('invert', C.c_int32), ('obj_aux', _dp), ('rng', C.c_int32),
                ('device', C.c_int32), ('max_batch', C.c_int32), ('reserved', C.c_int32)]


class GaRecord(C.Structure):
    _fields_ = [('best_f', C.c_double), ('sum_fitness', C.c_double), ('best_idx', C.c_int64), ('n_appended', C.c_int64),
                ('n_aliased', C.c_int64), ('n_duplicates', C.c_int64), ('generation', C.c_int32), ('reserved', C.c_int32)]


class GaTiming(C.Structure):
    _fields_ = [('total_ms', C.c_double), ('phase_ms', C.c_double * 8), ('generations', C.c_int64),
                ('profiled_generations', C.c_int64), ('gpu_launches', C.c_int64)]


class ReplayDraws(C.Structure):
    _fields_ = [('ib', _i64p), ('r1', _i64p), ('r2', _i64p), ('Uf', _dp), ('Uc', _dp)]


class Counters(C.Structure):
    _fields_ = [('kernel_ms', C.c_double), ('generations', C.c_int64), ('trial_evals', C.c_int64),
                ('gpu_launches', C.c_int64), ('bytes_algorithmic', C.c_int64), ('h2d_bytes', C.c_int64),
                ('d2h_bytes', C.c_int64), ('trial_kernel_ms', C.c_double), ('select_kernel_ms', C.c_double),
                ('profiled_generations', C.c_int64), ('exchange_wait_ms', C.c_double)]


EXPORTS = ['oode_abi_version', 'oode_device_count', 'oode_last_error', 'oode_nccl_unique_id', 'oode_create',
           'oode_destroy', 'oode_init', 'oode_step', 'oode_get_best_x', 'oode_get_population',
           'oode_set_population', 'oode_get_accept_mask', 'oode_get_trial_vals', 'oode_eval_points',
           'oode_get_counters', 'oode_cpython_stream', 'oode_set_profiling', 'oode_shard_layout', 'oode_exchange_mode',
           'oode_release_cached_memory', 'oode_set_constraints', 'oode_get_constr_vals', 'oode_trial_kernel_name',
           'oode_state_size', 'oode_save_state', 'oode_load_state', 'oode_crossover_threshold', 'oode_get_rows',
           'oode_get_donor_indices', 'oode_register_objective']
# include/ooga.h: the galileo genetic algorithm behind the same boundary
EXPORTS_GA = ['ooga_create', 'ooga_destroy', 'ooga_init', 'ooga_step', 'ooga_get_population', 'ooga_get_selection',
              'ooga_gpu_launches', 'ooga_set_profiling', 'ooga_get_timing', 'ooga_scan_values', 'ooga_set_population',
              'ooga_state_size', 'ooga_save_state', 'ooga_load_state']
GA_PHASES = ['evaluate', 'children', 'aliases', 'hash', 'objective', 'duplicates', 'sort', 'truncate']

_lib = None


class LibraryError(RuntimeError):
    pass


def load():
    """Load liboode.so; raises LibraryError if it has not been built (python -m openopt_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError('liboode.so is not built (%s missing): run `python -m openopt_b200.build`; '
                           'there is no CPU fallback for the de solver' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.oode_last_error.restype = C.c_char_p
    lib.oode_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.oode_destroy.argtypes = [C.c_void_p]
    lib.oode_destroy.restype = None
    lib.oode_init.argtypes = [C.c_void_p, _dp, C.POINTER(GenRecord)]
    lib.oode_step.argtypes = [C.c_void_p, C.c_int32, C.POINTER(ReplayDraws), C.POINTER(GenRecord)]
    lib.oode_get_best_x.argtypes = [C.c_void_p, C.c_int64, _dp]
    lib.oode_get_population.argtypes = [C.c_void_p, _dp, _dp]
    lib.oode_set_population.argtypes = [C.c_void_p, _dp, _dp]
    lib.oode_get_accept_mask.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
    lib.oode_get_trial_vals.argtypes = [C.c_void_p, _dp]
    lib.oode_eval_points.argtypes = [C.c_void_p, _dp, C.c_int64, _dp]
    if hasattr(lib, 'oode_get_rows'):          # absent from older A/B builds loaded through OODE_LIB_PATH
        lib.oode_get_rows.argtypes = [C.c_void_p, C.c_int32, _i64p, C.c_int64, _dp]
        lib.oode_get_donor_indices.argtypes = [C.c_void_p, _i64p, _i64p, _i64p]
    lib.oode_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.oode_exchange_mode.argtypes = [C.c_void_p]
    lib.oode_crossover_threshold.argtypes = [C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    lib.oode_state_size.argtypes = [C.c_void_p]
    lib.oode_state_size.restype = C.c_int64
    lib.oode_save_state.argtypes = [C.c_void_p, C.c_void_p]
    lib.oode_load_state.argtypes = [C.c_void_p, C.c_void_p]
    lib.oode_trial_kernel_name.argtypes = [C.c_void_p]
    lib.oode_trial_kernel_name.restype = C.c_char_p
    lib.oode_set_constraints.argtypes = [C.c_void_p, C.POINTER(Constraints)]
    lib.oode_get_constr_vals.argtypes = [C.c_void_p, _dp]
    lib.oode_release_cached_memory.restype = None
    lib.oode_nccl_unique_id.argtypes = [C.c_void_p]
    if hasattr(lib, 'oode_register_objective'):
        lib.oode_register_objective.argtypes = [C.c_char_p]
    lib.oode_set_profiling.argtypes = [C.c_void_p, C.c_int32]
    lib.oode_shard_layout.argtypes = [C.c_int32] * 4 + [C.POINTER(C.c_int32)] * 4
    lib.oode_cpython_stream.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, _i64p, C.c_int64, _dp]
    if hasattr(lib, 'ooga_create'):            # absent from older A/B builds loaded through OODE_LIB_PATH
        lib.ooga_create.argtypes = [C.POINTER(GaConfig), C.POINTER(C.c_void_p)]
        lib.ooga_destroy.argtypes = [C.c_void_p]
        lib.ooga_destroy.restype = None
        lib.ooga_init.argtypes = [C.c_void_p]
        lib.ooga_step.argtypes = [C.c_void_p, C.c_int32, C.POINTER(GaRecord), _dp]
        lib.ooga_get_population.argtypes = [C.c_void_p, _dp, _dp]
        lib.ooga_get_selection.argtypes = [C.c_void_p, _i64p, _i64p]
        lib.ooga_gpu_launches.argtypes = [C.c_void_p]
        lib.ooga_gpu_launches.restype = C.c_int64
        lib.ooga_scan_values.argtypes = [C.c_void_p, _dp, _dp, _dp, _i64p]
        if hasattr(lib, 'ooga_set_population'):
            lib.ooga_set_population.argtypes = [C.c_void_p, _dp, _dp]
            lib.ooga_state_size.argtypes = [C.c_void_p]
            lib.ooga_state_size.restype = C.c_int64
            lib.ooga_save_state.argtypes = [C.c_void_p, C.c_void_p]
            lib.ooga_load_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        lib.ooga_set_profiling.argtypes = [C.c_void_p, C.c_int32]
        lib.ooga_get_timing.argtypes = [C.c_void_p, C.POINTER(GaTiming)]
    if lib.oode_abi_version() != OODE_ABI_VERSION:
        raise LibraryError('liboode.so ABI version mismatch; rebuild it')
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


def last_error():
    return load().oode_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise LibraryError('%s failed (%d): %s' % (what, rc, last_error()))


def release_cached_memory():
    """Return the population blocks liboode keeps between handles to the CUDA driver."""
    load().oode_release_cached_memory()


def device_count():
    return load().oode_device_count()


def nccl_unique_id():
    """128-byte NCCL id (rank 0 creates it and ships it to the other ranks)."""
    buf = C.create_string_buffer(128)
    _check(load().oode_nccl_unique_id(buf), 'oode_nccl_unique_id')
    return buf.raw


def register_objective(library_path):
    """Load a separately compiled objective (openopt_b200/expr.py builds them) and return its objective id."""
    rc = load().oode_register_objective(os.fsencode(library_path))
    if rc < 0:
        raise LibraryError('oode_register_objective failed (%d): %s' % (rc, last_error()))
    return rc


def shard_layout(dim, objective, world_size, rank):
    """(col0, ncols, leaf0, nleaves) of a rank under column sharding."""
    out = [C.c_int32() for _ in range(4)]
    _check(load().oode_shard_layout(dim, objective, world_size, rank, *[C.byref(o) for o in out]), 'oode_shard_layout')
    return tuple(o.value for o in out)


def crossover_threshold(cross_rate):
    """(threshold, keep_all): the Philox-mode crossover test on the raw 32-bit word (oode_crossover_threshold)."""
    t, a = C.c_uint32(), C.c_int32()
    _check(load().oode_crossover_threshold(float(cross_rate), C.byref(t), C.byref(a)), 'oode_crossover_threshold')
    return t.value, bool(a.value)


def cpython_stream(seed, below, n_ints, n_doubles):
    ints = np.empty(max(n_ints, 1), dtype=np.int64)
    dbl = np.empty(max(n_doubles, 1), dtype=np.float64)
    _check(load().oode_cpython_stream(seed, below, n_ints, ints.ctypes.data_as(_i64p), n_doubles, _d(dbl)),
           'oode_cpython_stream')
    return ints[:n_ints], dbl[:n_doubles]


class DeviceDE:
    """One device-resident DE population (an oode_handle).  Mirrors the state de.__solver__ keeps in
    local variables (pop, vals, best, Best; de_oo.py:147-155,178-286)."""

    def __init__(self, objective, lb, ub, x0=None, population=None, diff_factor=0.8, cross_rate=0.5,
                 seed=150880, obj_aux=None, rng='philox', invert=False, device=0, max_batch=64,
                 world_size=1, rank=0, shard=SHARD_NONE, nccl_id=None, base='random', direction='random',
                 factor='random', hndvi=1):
        self.lib = load()
        self.lb = np.ascontiguousarray(lb, dtype=np.float64).ravel()
        self.ub = np.ascontiguousarray(ub, dtype=np.float64).ravel()
        self.D = self.lb.size
        self.NP = int(10 * self.D if population is None else population)
        self._x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).ravel()
        self._aux = None if obj_aux is None else np.ascontiguousarray(obj_aux, dtype=np.float64).ravel()
        self._nccl = nccl_id
        cfg = Config()
        cfg.abi_version = OODE_ABI_VERSION
        cfg.dim, cfg.population = self.D, self.NP
        cfg.lb, cfg.ub = _d(self.lb), _d(self.ub)
        cfg.x0 = _d(self._x0) if self._x0 is not None else None
        cfg.diff_factor, cfg.cross_rate, cfg.seed = float(diff_factor), float(cross_rate), int(seed)
        cfg.objective, cfg.invert = int(objective), int(bool(invert))
        cfg.obj_aux = _d(self._aux) if self._aux is not None else None
        cfg.rng = RNG_BY_NAME[rng] if isinstance(rng, str) else int(rng)
        cfg.device, cfg.world_size, cfg.rank, cfg.shard = int(device), int(world_size), int(rank), int(shard)
        cfg.max_batch = int(max_batch)
        cfg.base_strategy, cfg.direction_strategy = BASE_BY_NAME[base], DIRECTION_BY_NAME[direction]
        cfg.factor_strategy, cfg.hndvi = FACTOR_BY_NAME[factor], int(hndvi)
        self.strategy = (base, direction, factor, int(hndvi))
        self._nccl_buf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        cfg.nccl_id = C.cast(self._nccl_buf, C.c_void_p) if nccl_id is not None else None
        self.max_batch = int(max_batch)
        self.rng = cfg.rng
        self.h = C.c_void_p()
        _check(self.lib.oode_create(C.byref(cfg), C.byref(self.h)), 'oode_create')
        self.generation = -1
        self.constrained = False

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            self.lib.oode_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_constraints(self, A=None, b=None, Aeq=None, beq=None, c=None, h=None, contol=1e-6):
        """General constraints (call before init).  A/b, Aeq/beq: arrays; c, h: DiagQuadratic objects."""
        cs = Constraints()
        keep = []

        def mat(M, rows_of=None):
            if M is None:
                return None, 0
            M = np.ascontiguousarray(np.atleast_2d(np.asarray(M, dtype=np.float64)))
            if M.size == 0:
                return None, 0
            assert M.shape[1] == self.D
            keep.append(M)
            return _d(M), M.shape[0]

        def vec(v, n):
            if not n:
                return None
            v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
            assert v.size == n
            keep.append(v)
            return _d(v)

        cs.A, cs.n_lin_ineq = mat(A)
        cs.b = vec(b, cs.n_lin_ineq)
        cs.Aeq, cs.n_lin_eq = mat(Aeq)
        cs.beq = vec(beq, cs.n_lin_eq)
        cs.Qc, cs.n_quad_ineq = mat(None if c is None else c.Q)
        cs.rc = vec(None if c is None else c.r, cs.n_quad_ineq)
        cs.Qh, cs.n_quad_eq = mat(None if h is None else h.Q)
        cs.rh = vec(None if h is None else h.r, cs.n_quad_eq)
        cs.contol = float(contol)
        _check(self.lib.oode_set_constraints(self.h, C.byref(cs)), 'oode_set_constraints')
        self.constrained = True

    def constr_vals(self):
        v = np.empty(self.NP, dtype=np.float64)
        _check(self.lib.oode_get_constr_vals(self.h, _d(v)), 'oode_get_constr_vals')
        return v

    def init(self, u0=None):
        rec = GenRecord()
        up = None
        if u0 is not None:
            u0 = np.ascontiguousarray(u0, dtype=np.float64)
            assert u0.size == self.NP * self.D
            up = _d(u0)
        _check(self.lib.oode_init(self.h, up, C.byref(rec)), 'oode_init')
        self.generation = 0
        return rec

    def step(self, n_gens=1, draws=None):
        """draws (replay mode): one entry per generation -- a tuple (ib, r1, r2, Uf, Uc) or an object with
        those attributes (plus `r`, the 2h indices of the directed search); items a strategy does not
        draw are None."""
        recs = (GenRecord * n_gens)()
        dp = None
        keep = []
        if draws is not None:
            assert len(draws) == n_gens
            base, direction, factor, h = self.strategy
            arr = (ReplayDraws * n_gens)()

            def ints(a, n):
                if a is None:
                    return None
                a = np.ascontiguousarray(a, dtype=np.int64).ravel()
                assert a.size == n
                keep.append(a)
                return a.ctypes.data_as(_i64p)

            def dbls(a):
                if a is None:
                    return None
                a = np.ascontiguousarray(a, dtype=np.float64)
                assert a.size == self.NP * self.D
                keep.append(a)
                return _d(a)

            for k, d in enumerate(draws):
                if isinstance(d, (tuple, list)):
                    ib, r1, r2, Uf, Uc = d
                    r = None
                else:
                    ib, r1, r2, Uf, Uc, r = d.ib, d.r1, d.r2, d.Uf, d.Uc, getattr(d, 'r', None)
                arr[k].ib = ints(ib, self.NP) if base == 'random' else None
                if direction == 'random':
                    arr[k].r1, arr[k].r2 = ints(r1, h * self.NP), ints(r2, h * self.NP)
                else:
                    arr[k].r1, arr[k].r2 = ints(r, 2 * h), None
                arr[k].Uf = dbls(Uf) if factor == 'random' else None
                arr[k].Uc = dbls(Uc)
            dp = arr
        _check(self.lib.oode_step(self.h, n_gens, dp, recs), 'oode_step')
        self.generation += n_gens
        return list(recs)

    def best_x(self, generation=None):
        x = np.empty(self.D, dtype=np.float64)
        g = self.generation if generation is None else generation
        _check(self.lib.oode_get_best_x(self.h, g, _d(x)), 'oode_get_best_x')
        return x

    def population(self):
        pop = np.empty((self.NP, self.D), dtype=np.float64)
        vals = np.empty(self.NP, dtype=np.float64)
        _check(self.lib.oode_get_population(self.h, _d(pop), _d(vals)), 'oode_get_population')
        return pop, vals

    def population_vals(self):
        """f of the current population only (no row data crosses the bus)."""
        vals = np.empty(self.NP, dtype=np.float64)
        _check(self.lib.oode_get_population(self.h, None, _d(vals)), 'oode_get_population')
        return vals

    def set_population(self, pop, vals):
        pop = np.ascontiguousarray(pop, dtype=np.float64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert pop.shape == (self.NP, self.D) and vals.shape == (self.NP,)
        _check(self.lib.oode_set_population(self.h, _d(pop), _d(vals)), 'oode_set_population')

    def save_state(self):
        """Checkpoint: everything the next generation depends on, as one uint8 array."""
        n = self.lib.oode_state_size(self.h)
        if n < 0:
            raise LibraryError('oode_state_size failed: %s' % last_error())
        buf = np.empty(n, dtype=np.uint8)
        _check(self.lib.oode_save_state(self.h, buf.ctypes.data_as(C.c_void_p)), 'oode_save_state')
        return buf

    def load_state(self, buf):
        """Resume from save_state() of a handle with the same configuration (call init() first)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        assert buf.size == self.lib.oode_state_size(self.h)
        _check(self.lib.oode_load_state(self.h, buf.ctypes.data_as(C.c_void_p)), 'oode_load_state')
        hdr = np.frombuffer(buf.tobytes()[:24], dtype=np.int64)
        self.generation = int(hdr[2])

    def accept_mask(self):
        m = np.empty(self.NP, dtype=np.uint8)
        _check(self.lib.oode_get_accept_mask(self.h, m.ctypes.data_as(C.POINTER(C.c_uint8))), 'oode_get_accept_mask')
        return m

    def trial_vals(self):
        v = np.empty(self.NP, dtype=np.float64)
        _check(self.lib.oode_get_trial_vals(self.h, _d(v)), 'oode_get_trial_vals')
        return v

    ROWS_CURRENT, ROWS_SPARE, ROWS_TRIAL, ROWS_PARENT = range(4)

    def rows(self, idx, which=0):
        """[len(idx)][D] copy of the rows `idx` (global row numbers): their current version, spare slot, the trial the
        last generation built for them or the parent it read (oode_get_rows) -- cheap at any population size."""
        idx = np.ascontiguousarray(idx, dtype=np.int64).ravel()
        out = np.zeros((idx.size, self.D), dtype=np.float64)
        _check(self.lib.oode_get_rows(self.h, int(which), idx.ctypes.data_as(_i64p), idx.size, _d(out)), 'oode_get_rows')
        return out

    def donor_indices(self):
        """(ib, r1, r2), each [NP] int64: the donor rows the NEXT generation will use."""
        a = [np.empty(self.NP, dtype=np.int64) for _ in range(3)]
        _check(self.lib.oode_get_donor_indices(self.h, *[x.ctypes.data_as(_i64p) for x in a]), 'oode_get_donor_indices')
        return tuple(a)

    def eval_points(self, x):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        assert x.shape[1] == self.D
        f = np.empty(x.shape[0], dtype=np.float64)
        _check(self.lib.oode_eval_points(self.h, _d(x), x.shape[0], _d(f)), 'oode_eval_points')
        return f

    def exchange_mode(self):
        """XCHG_* of this handle: how a sharded population moves its leaf sums between the GPUs."""
        return self.lib.oode_exchange_mode(self.h)

    def trial_kernel_name(self):
        return self.lib.oode_trial_kernel_name(self.h).decode()

    def set_profiling(self, on=True):
        _check(self.lib.oode_set_profiling(self.h, int(bool(on))), 'oode_set_profiling')

    def counters(self):
        c = Counters()
        _check(self.lib.oode_get_counters(self.h, C.byref(c)), 'oode_get_counters')
        return {k: getattr(c, k) for k, _ in Counters._fields_}


class DeviceGalileo:
    """One device-resident GA population (an ooga_handle, include/ooga.h).  Mirrors the `Population` object
    galileo.__solver__ builds (openopt/solvers/Standalone/galileo_oo.py:35-70)."""

    def __init__(self, objective, lb, ub, population=15, crossover_rate=1.0, mutation_rate=0.05, seed=150880,
                 obj_aux=None, rng='philox', invert=False, device=0, max_batch=64):
        self.lib = load()
        self.lb = np.ascontiguousarray(lb, dtype=np.float64).ravel()
        self.ub = np.ascontiguousarray(ub, dtype=np.float64).ravel()
        self.D, self.N = self.lb.size, int(population)
        self.P = (self.N + 1) // 2
        self._aux = None if obj_aux is None else np.ascontiguousarray(obj_aux, dtype=np.float64).ravel()
        cfg = GaConfig()
        cfg.abi_version = OODE_ABI_VERSION
        cfg.dim, cfg.population = self.D, self.N
        cfg.lb, cfg.ub = _d(self.lb), _d(self.ub)
        cfg.crossover_rate, cfg.mutation_rate, cfg.seed = float(crossover_rate), float(mutation_rate), int(seed)
        cfg.objective, cfg.invert = int(objective), int(bool(invert))
        cfg.obj_aux = _d(self._aux) if self._aux is not None else None
        cfg.rng = RNG_BY_NAME[rng] if isinstance(rng, str) else int(rng)
        cfg.device, cfg.max_batch = int(device), int(max_batch)
        self.max_batch = int(max_batch)
        self.h = C.c_void_p()
        _check(self.lib.ooga_create(C.byref(cfg), C.byref(self.h)), 'ooga_create')
        self.generation = -1

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            self.lib.ooga_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def init(self):
        _check(self.lib.ooga_init(self.h), 'ooga_init')
        self.generation = 0

    def step(self, n_gens=1):
        """-> (records, best_x[n_gens][D]): what __solver__ hands to p.iterfcn per iteration (galileo_oo.py:86-89)."""
        recs = (GaRecord * n_gens)()
        xs = np.empty((n_gens, self.D), dtype=np.float64)
        _check(self.lib.ooga_step(self.h, n_gens, recs, _d(xs)), 'ooga_step')
        self.generation += n_gens
        return list(recs), xs

    def population(self):
        """(genes [N][D], cached fitness [N]) of currentGeneration in list order."""
        genes = np.empty((self.N, self.D), dtype=np.float64)
        fit = np.empty(self.N, dtype=np.float64)
        _check(self.lib.ooga_get_population(self.h, _d(genes), _d(fit)), 'ooga_get_population')
        return genes, fit

    def set_population(self, genes, fitness=None):
        """Replace currentGeneration (list order as given); fitness=None: evaluated on the device."""
        genes = np.ascontiguousarray(genes, dtype=np.float64)
        assert genes.shape == (self.N, self.D)
        fp = None
        if fitness is not None:
            fitness = np.ascontiguousarray(fitness, dtype=np.float64)
            assert fitness.shape == (self.N,)
            fp = _d(fitness)
        _check(self.lib.ooga_set_population(self.h, _d(genes), fp), 'ooga_set_population')

    def save_state(self):
        """Checkpoint: everything the next iteration depends on, as one uint8 array."""
        n = self.lib.ooga_state_size(self.h)
        if n < 0:
            raise LibraryError('ooga_state_size failed: %s' % last_error())
        buf = np.empty(n, dtype=np.uint8)
        _check(self.lib.ooga_save_state(self.h, buf.ctypes.data_as(C.c_void_p)), 'ooga_save_state')
        return buf

    def load_state(self, buf):
        """Resume from save_state() of a handle with the same configuration (call init() first)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        _check(self.lib.ooga_load_state(self.h, buf.ctypes.data_as(C.c_void_p), buf.size), 'ooga_load_state')
        self.generation = int(np.frombuffer(buf.tobytes()[16:24], dtype=np.int64)[0])

    def selection(self):
        """(picks [2P], cut [P]) of the last iteration (verification tap)."""
        picks = np.empty(2 * self.P, dtype=np.int64)
        cut = np.empty(self.P, dtype=np.int64)
        _check(self.lib.ooga_get_selection(self.h, picks.ctypes.data_as(_i64p), cut.ctypes.data_as(_i64p)), 'ooga_get_selection')
        return picks, cut

    def gpu_launches(self):
        return int(self.lib.ooga_gpu_launches(self.h))

    def scan_values(self, values):
        """(running sums [N], total, index of the first largest value) of `values` [N] as the evaluate step computes them."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert values.shape == (self.N,)
        out = np.empty(self.N, dtype=np.float64)
        tot, bi = C.c_double(), C.c_int64()
        _check(self.lib.ooga_scan_values(self.h, _d(values), _d(out), C.byref(tot), C.byref(bi)), 'ooga_scan_values')
        return out, tot.value, bi.value

    def set_profiling(self, on=True):
        _check(self.lib.ooga_set_profiling(self.h, int(bool(on))), 'ooga_set_profiling')

    def timing(self):
        """Device time so far: total_ms over `generations`, and per phase (GA_PHASES) over `profiled_generations`."""
        t = GaTiming()
        _check(self.lib.ooga_get_timing(self.h, C.byref(t)), 'ooga_get_timing')
        return dict(total_ms=t.total_ms, generations=t.generations, profiled_generations=t.profiled_generations,
                    gpu_launches=t.gpu_launches, phase_ms=dict(zip(GA_PHASES, list(t.phase_ms))))

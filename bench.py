#!/usr/bin/env python
"""Benchmark of the DE generation loop (BASELINE.json metric: DE trial evals/s; % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

A *step* is one generation: one pass of the hot path (mutation, crossover, bound repair, objective,
selection, argmin) over the whole synthetic population.  Default workload = BASELINE configs[2]
(the configuration the north-star target is quoted on): Ackley 1000-D, population 1M, which fits
one GPU (2 x 8 GB slot arrays), so it is also the N=1 workload.

One JSON line is printed by rank 0:
  value     whole-job trial evaluations / s, population resident in HBM, timed on the device
            (CUDA events on the library's stream around exactly K generations; max over ranks)
  e2e       the same metric through the public API `GLP(...).solve('de')` with host buffers
            (wall clock around the call: allocation, H2D of lb/ub/x0, K generations, D2H of the
            per-generation records and Best.x)
  roofline  de_trial_kernel: algorithmic bytes (40*D+16 per trial) / its average
            launch duration (CUDA events inside the library), against the measured HBM peak
  e2e_cold  the same call as the first solve of a process (liboode's block cache emptied first)
  roofline  the trial kernel: algorithmic bytes (40*D+16 per trial) / its average launch duration
            (CUDA events inside the library, recorded INSIDE the timed region when a generation is long enough),
            against the measured HBM peak; `traffic` = ncu DRAM bytes per launch, reported only while
            profiles/traffic.json is stamped with the hash of the current kernel sources
  cpu_baseline  oracle/de_oracle.c (the C port of the reference's algorithm) on every host core, and
            `reference_python`: the UNMODIFIED reference (baseline/_ref) timed on one core of this box
  parity_vs_1gpu, per_rank, wait_ms  (N > 1) a reduced-population twin run on every rank alone and sharded over all
            ranks must be identical; each rank's trial kernel / selection pass / wait for the slowest rank

`--impl reference` times the C port alone, all host threads (set by the bench itself: torchrun exports
OMP_NUM_THREADS=1), and carries the unmodified reference's own speed beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (objective, D, NP, lb, ub)  -- SURVEY.md section 8d
    'c2': ('rastrigin', 100, 10000, -5.12, 5.12),
    'c3': ('ackley', 1000, 1000000, -32.768, 32.768),
    'c4': ('griewank', 200, 262144, -600.0, None),
    'c5_1m': ('rastrigin', 100, 1 << 20, -5.12, 5.12),
    'c5_16m': ('rastrigin', 100, 1 << 24, -5.12, 5.12),
}
SEED, F, CR = 150880, 0.8, 0.5
# the sibling population solver: galileo on the C5 row shape.  goal='max' makes the fitness
# (-p.f) positive, so the roulette spreads over the whole population (on a minimised positive objective the reference's
# wheel only ever lands on the first or the last chromosome).
GA_WORKLOADS = {
    # name: (objective, D, N, lb, ub, goal)
    'ga_c5_64k': ('rastrigin', 100, 1 << 16, -5.12, 5.12, 'max'),
    'ga_c5_1m': ('rastrigin', 100, 1 << 20, -5.12, 5.12, 'max'),
}


def workload(name):
    obj, D, NP, lo, hi = WORKLOADS[name]
    lb = lo * np.ones(D)
    ub = np.linspace(100.0, 900.0, D) if hi is None else hi * np.ones(D)
    x0 = lb + 0.25 * (ub - lb)
    return obj, D, NP, lb, ub, x0


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons while the timed region runs: NVML (nvidia_ml_py) every 5 ms when it is
    available, else `nvidia-smi` every 200 ms."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag, self.index, self.source = False, index, None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[index]) if visible and all(v.strip().isdigit() for v in visible.split(',')) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.nv = pynvml
            self.source = 'nvml'
        except Exception:
            self.nv = None
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (('hw_slowdown', nv.nvmlClocksThrottleReasonHwSlowdown),
                          ('hw_thermal_slowdown', nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ('sw_thermal_slowdown', nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ('sw_power_cap', nv.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
        v = [x.strip() for x in out.strip().split(',')]
        if len(v) >= 6:
            if v[0].replace('.', '').isdigit():
                self.sm.append(float(v[0]))
            if v[1].replace('.', '').isdigit():
                self.mx.append(float(v[1]))
            for name, x in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), v[2:6]):
                if x.lower().startswith('active'):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self.nv is not None:          # NVML misbehaves: fall back to nvidia-smi
                    self.nv, self.source = None, 'nvidia-smi'
            time.sleep(0.005 if self.nv is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': self.source}


# the translation unit of the trial kernel (csrc/trial_obj.cu and every header it pulls in): what its DRAM traffic depends on
TRIAL_KERNEL_SOURCES = ('trial_obj.cu', 'de_launch.h', 'de_kernels.cuh', 'de_device.cuh', 'de_trial_tma.cuh', 'de_trial_rows.cuh')


def csrc_sha256():
    """Hash of the CODE of the trial kernel's translation unit (comments and blank lines stripped):
    profiles/traffic.json is stamped with it when an ncu capture is recorded (scripts/update_traffic.py), so a
    DRAM-traffic figure taken from an older kernel is never reported, while editing a comment -- or host code and other
    kernels (csrc/oode.cu, the galileo sources) -- does not invalidate a capture."""
    import hashlib
    import re
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'openopt_b200', 'csrc')
    for f in sorted(TRIAL_KERNEL_SOURCES):
        src = open(os.path.join(d, f), encoding='utf-8').read()
        src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)                 # block comments
        src = re.sub(r'//[^\n]*', '', src)                              # line comments (no string literal here holds //)
        code = '\n'.join(ln.strip() for ln in src.splitlines() if ln.strip())
        h.update(f.encode())
        h.update(code.encode())
    return h.hexdigest()[:16]


def traffic_for(workload_name, kernel_name):
    """(dram bytes per launch, provenance) of the workload's trial kernel from the last `ncu --set full` capture,
    or (None, why) when there is none for the CURRENT kernel sources."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            e = json.load(f).get(workload_name)
    except Exception:
        e = None
    if not e:
        return None, 'no ncu capture recorded for this workload'
    if e.get('csrc_sha256') != csrc_sha256():
        return None, 'stale: recorded for kernel sources %s (%s), current %s' % (e.get('csrc_sha256'), e.get('git_sha'), csrc_sha256())
    return e.get('dram_bytes_per_launch'), {k: e.get(k) for k in ('report', 'kernel', 'git_sha', 'csrc_sha256', 'captured')}


def cpu_port_run(obj, D, lb, ub, x0, sample_rows, gens):
    """The C port (oracle/de_oracle.c) on `sample_rows` rows of the same workload, all host threads (set here:
    torchrun exports OMP_NUM_THREADS=1 to its ranks, which must not shrink the baseline)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import c_oracle
    c_oracle.set_threads()
    c = c_oracle.COracleDE(obj, lb, ub, x0=x0, population=sample_rows, F=F, Cr=CR, seed=SEED, stream='philox')
    t0 = time.perf_counter()
    c.step(gens)
    dt = time.perf_counter() - t0
    threads = c.threads
    c.close()
    return sample_rows * gens / dt, dt, threads


REF_PY_SIZES = {   # workload -> (NP, generations) the pure-Python reference finishes in a few seconds
    'c2': (10000, 3), 'c3': (1000, 2), 'c4': (2000, 3), 'c5_1m': (10000, 3), 'c5_16m': (10000, 3)}


def reference_python(name):
    """The UNMODIFIED reference (baseline/_ref, OpenOpt 0.43) timed on one core of this box through its own
    `GLP(...).solve('de')`, at a population it can finish (evals/s is ~independent of NP at fixed D, so larger
    populations are extrapolations), plus the line with its per-element Python RNG loop patched out."""
    obj, D, NP, lo, hi = WORKLOADS[name]
    rows, gens = REF_PY_SIZES[name]
    out = {}
    for key, flag in (('stock', []), ('numpy_rng', ['--numpy-rng'])):
        cmd = [sys.executable, os.path.join(ROOT, 'baseline', 'time_reference.py'), obj, str(D), str(rows), str(gens),
               repr(lo), 'lin' if hi is None else repr(hi)] + flag
        try:
            env = {k: v for k, v in os.environ.items() if k not in ('OMP_NUM_THREADS',)}
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
            out[key] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:          # noqa: BLE001 -- a missing baseline/_ref must not cost the bench line
            out[key] = {'unavailable': '%s: %s' % (type(e).__name__, e)}
    out['note'] = ('unmodified OpenOpt 0.43 de (baseline/_ref), 1 core (single-threaded by construction), NP=%d of %d rows, '
                   '%d generations; numpy_rng = Rand/Randint monkeypatched to numpy RandomState' % (rows, NP, gens))
    return out


def run_reference(args):
    obj, D, NP, lb, ub, x0 = workload(args.workload)
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rows = min(NP, max(1024, int(2.0e7 // D)))           # ~2e7 gene evaluations per step
    cpu_port_run(obj, D, lb, ub, x0, rows, max(1, args.warmup))      # warm-up
    evals_s, dt, threads = cpu_port_run(obj, D, lb, ub, x0, rows, args.steps)
    sample = '%d of %d rows, %d generations per run' % (rows, NP, args.steps)
    line = {
        'impl': 'reference', 'metric': 'de_trial_evals_per_sec', 'value': evals_s, 'unit': 'evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': describe(args.workload), 'sample': sample},
        'cpu_baseline': {'value': evals_s, 'unit': 'evals/s', 'cores': threads, 'nproc': os.cpu_count(), 'kind': 'port',
                         'sample': sample, 'reference_python': reference_python(args.workload)},
        'e2e': {'value': evals_s, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'value = C port of the reference algorithm (oracle/de_oracle.c, OpenMP over rows, every host core): '
                'the strongest CPU form of the path.  cpu_baseline.reference_python = the unmodified pure-Python '
                'reference itself on this box (1 core), ~2 orders of magnitude slower than the port',
    }
    return line


def describe(name):
    obj, D, NP, lo, hi = WORKLOADS[name]
    return "GLP 'de' %s %d-D, population %d, F=%.1f Cr=%.1f, Philox draws" % (obj, D, NP, F, CR)


def parity_vs_one_gpu(dobj, D, lb, ub, x0, local_rank, extra, rows=65536, gens=5):
    """N > 1: every rank runs a reduced-population twin of the workload on its own GPU alone and the same problem
    sharded over all ranks; records (trial-best index and f, Best f, accepted / repaired counts), Best.x and the
    post-selection vals must be IDENTICAL (draws are keyed by global row / column, the leaf sums are folded in the
    same tree order on every rank).  The result is the AND over all ranks."""
    import hashlib
    import torch
    import torch.distributed as dist
    from openopt_b200 import _lib

    def rec(rs):
        return [(r.generation, r.trial_best_idx, r.trial_best_f, r.best_f, r.n_accepted, r.n_repaired, r.best_changed) for r in rs]

    kw = dict(x0=x0, population=rows, diff_factor=F, cross_rate=CR, seed=SEED, obj_aux=dobj.aux(D), rng='philox',
              device=local_rank, max_batch=gens)
    with _lib.DeviceDE(dobj.device_id, lb, ub, **kw) as one:
        r1 = [one.init()] + one.step(gens)
        bx1, v1 = one.best_x(), one.population_vals()
    with _lib.DeviceDE(dobj.device_id, lb, ub, **kw, **extra) as sh:
        r2 = [sh.init()] + sh.step(gens)
        bx2, v2 = sh.best_x(), sh.population_vals()
    mine = [rec(r1) == rec(r2), bool(np.array_equal(bx1, bx2)), bool(np.array_equal(v1, v2))]
    t = torch.tensor([int(x) for x in mine], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = [bool(x) for x in t.tolist()]
    return {'records': ok[0], 'best_x': ok[1], 'vals': ok[2], 'vals_sha': hashlib.sha256(v2.tobytes()).hexdigest()[:16],
            'vals_sha_1gpu': hashlib.sha256(v1.tobytes()).hexdigest()[:16], 'population': rows, 'generations': gens,
            'ranks_checked': dist.get_world_size()}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from openopt_b200 import GLP, _lib, objectives

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    obj, D, NP, lb, ub, x0 = workload(args.workload)
    dobj = objectives.get(obj)
    K, W = args.steps, args.warmup
    peak, peak_src = measured_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    extra = {}
    shard_kind = None
    if world > 1:
        # the population is column-sharded on the pairwise-sum tree: every GPU
        # holds all rows of its column block; per generation only the leaf sums cross the GPUs
        from openopt_b200 import distributed as ddist
        try:                                  # rows too narrow to give every GPU a leaf (C5: 100 columns): row shards
            _lib.shard_layout(D, dobj.device_id, world, rank)
            shard_kind = _lib.SHARD_COLS
        except _lib.LibraryError:
            shard_kind = _lib.SHARD_ROWS
        extra = dict(world_size=world, rank=rank, shard=shard_kind, nccl_id=ddist.shared_nccl_id())
    parity = None
    if world > 1:
        parity = parity_vs_one_gpu(dobj, D, lb, ub, x0, local_rank, extra)
    dev = _lib.DeviceDE(dobj.device_id, lb, ub, x0=x0, population=NP, diff_factor=F, cross_rate=CR, seed=SEED,
                        obj_aux=dobj.aux(D), rng='philox', device=local_rank, max_batch=max(K, W, 1), **extra)
    xchg = _lib.XCHG_NAMES[dev.exchange_mode()]
    kernel_name = dev.trial_kernel_name()
    dev.init()
    if W:
        dev.step(W)
    # ---- timed region: exactly K generations, device time from CUDA events on the launch stream ----
    # Per-kernel CUDA events (trial / selection) are recorded INSIDE the timed region when a generation is long enough
    # for three event records to be noise (< 0.1 %); for small populations (generations of tens of microseconds) they
    # come from a second pass over the same K generations so that they do not perturb `value`.
    inline_events = NP * D / max(world, 1) >= 2.0e7
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if inline_events:
        dev.set_profiling(True)
    c0 = dev.counters()
    t0 = time.perf_counter()
    recs = dev.step(K)
    barrier()
    wall = time.perf_counter() - t0
    c1 = dev.counters()
    ms = c1['kernel_ms'] - c0['kernel_ms']
    launches = c1['gpu_launches'] - c0['gpu_launches']
    if inline_events:
        dev.set_profiling(False)
        p0, p1, nprof = c0, c1, K
    else:
        # ---- roofline pass: same K generations with per-kernel events ----
        dev.set_profiling(True)
        p0 = dev.counters()
        dev.step(K)
        p1 = dev.counters()
        dev.set_profiling(False)
        nprof = 2 * K
    clocks = sampler.summary()
    trial_ms = (p1['trial_kernel_ms'] - p0['trial_kernel_ms']) / K
    select_ms = (p1['select_kernel_ms'] - p0['select_kernel_ms']) / K
    # time this rank's selection pass spent waiting for the other ranks' leaf sums (inter-rank skew), per generation
    wait_ms = (p1['exchange_wait_ms'] - c0['exchange_wait_ms']) / nprof if world > 1 else 0.0
    acc = float(np.mean([r.n_accepted for r in recs])) / NP
    dev.close()
    del dev

    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = NP * K / (ms * 1e-3)
    if world > 1:
        t = torch.tensor([trial_ms, select_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        trial_ms, select_ms = float(t[0].item()), float(t[1].item())
        mine = torch.tensor([(p1['trial_kernel_ms'] - p0['trial_kernel_ms']) / K, (p1['select_kernel_ms'] - p0['select_kernel_ms']) / K,
                             wait_ms], dtype=torch.float64, device='cuda')
        allr = [torch.zeros(3, dtype=torch.float64, device='cuda') for _ in range(world)]
        dist.all_gather(allr, mine)
        # per rank: its own trial kernel, its selection pass (which starts by waiting for the other ranks' leaf sums) and
        # that wait -- the inter-rank skew as numbers
        per_rank = [{'trial_ms': float(a[0]), 'select_ms': float(a[1]), 'wait_ms': float(a[2])} for a in allr]
        wait_ms = [r['wait_ms'] for r in per_rank]
    # per launch and per GPU: this rank's share of the columns of all NP rows
    bytes_per_launch = NP * (40 * D + 16) / world
    achieved = bytes_per_launch / (trial_ms * 1e-3) / 1e9

    # ---- e2e: the call a user makes, host buffers in, result out ----
    def solve_e2e():
        barrier()
        t0 = time.perf_counter()
        p = GLP(dobj, lb=lb, ub=ub, x0=x0, maxIter=K + 1, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9, iprint=-1)
        r = p.solve('de', population=NP, gensPerCall=max(1, min(K, 64)))
        wall_s = time.perf_counter() - t0
        sc = p.solver.counters
        if world > 1:
            t = torch.tensor([wall_s], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall_s = float(t.item())
        return {'value': sc['trial_evals'] / wall_s, 'unit': 'evals/s',
                'h2d_bytes_per_step': sc['h2d_bytes'] / max(1, sc['generations']),
                'd2h_bytes_per_step': sc['d2h_bytes'] / max(1, sc['generations']),
                'generations': sc['generations'], 'wall_s': wall_s, 'ff': float(r.ff), 'istop': int(r.istop)}

    # cold: the first solve of a process -- liboode's block cache is emptied first, so the call pays the
    # cudaMalloc of both slot arrays; warm (`e2e`): the same call again, blocks reused from the cache
    _lib.release_cached_memory()
    e2e_cold = solve_e2e()
    e2e_cold['note'] = 'block cache emptied first (oode_release_cached_memory): includes cudaMalloc of the slot arrays'
    e2e = solve_e2e()
    e2e['note'] = 'device blocks reused from the process-wide cache (second and later solves of a process)'

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    traffic, traffic_src = traffic_for(args.workload, kernel_name)
    rows = min(NP, max(1024, int(2.0e7 // D)))
    cpu = None
    if world == 1:
        cpu_port_run(obj, D, lb, ub, x0, rows, 2)                                       # warm-up (threads, caches)
        cpu_gens = 24                                                                   # ~1-2 s wall = 15-30 core-seconds
        cpu_evals_s, cpu_dt, threads = cpu_port_run(obj, D, lb, ub, x0, rows, cpu_gens)
        cpu = {'value': cpu_evals_s, 'unit': 'evals/s', 'cores': threads, 'nproc': os.cpu_count(), 'kind': 'port',
               'reference_python': reference_python(args.workload),
               'sample': '%d of %d rows, %d generations, %.2f s wall (oracle/de_oracle.c, OpenMP)'
                         % (rows, NP, cpu_gens, cpu_dt)}
    line = {
        'metric': 'de_trial_evals_per_sec', 'value': value, 'unit': 'evals/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': describe(args.workload), 'objective': obj, 'D': D, 'population': NP,
                   'rng': 'philox4x32-10', 'l2': 'inputs larger than L2 (2 x %.1f GB slot arrays)' % (NP * D * 8 / 1e9)
                   if NP * D * 8 > 2.5e8 else 'population fits L2; reported as is', 'generations_per_s': 1e3 * K / ms,
                   'acceptance_rate': acc,
                   'parallelism': (('column shards x%d on the pairwise-sum tree, leaf sums exchanged by %s' % (world, xchg))
                                   if shard_kind == _lib.SHARD_COLS else
                                   ('row shards x%d, accepted rows pushed into every replica over NVLink' % world))
                   if world > 1 else 'single GPU',
                   'trial_kernel': kernel_name},
        'roofline': {'bound': 'hbm', 'kernel': kernel_name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'peak_source': peak_src, 'traffic': traffic if world == 1 else None,
                     'traffic_source': traffic_src if world == 1 else None,
                     'bytes_per_launch': bytes_per_launch, 'kernel_ms': trial_ms,
                     'kernel_events': 'inside the timed region' if inline_events else 'second pass over the same generations', 'select_kernel_ms': select_ms,
                     'frac_whole_generation': bytes_per_launch / (ms / K * 1e-3) / 1e9 / peak},
        'cpu_baseline': cpu,
        'parity_vs_1gpu': parity, 'wait_ms': wait_ms if world > 1 else None, 'per_rank': per_rank if world > 1 else None,
        'e2e': e2e, 'e2e_cold': e2e_cold, 'gpu_launches': int(launches), 'clocks': clocks, 'wall_s_timed_region': wall,
    }
    return line


GA_REF_POPULATION = 512       # what the unmodified galileo finishes in about a second per run (its duplicate scan is O(N^2 * D))


def ga_describe(name):
    obj, D, N, lo, hi, goal = GA_WORKLOADS[name]
    return ('galileo (genetic algorithm) %s %d-D, population %d, goal=%s, crossoverRate 1.0, mutationRate 0.05' % (obj, D, N, goal))


def ga_cpu_port(obj, lb, ub, goal, n_cpu=32768, it_cpu=3):
    """(entries/s, sample description) of the numpy oracle's vectorised step (oracle/galileo_oracle.py: the CPU port of
    the path; one core, as numpy's element-wise loops are single-threaded) on a bounded population of the same rows."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import de_oracle as orc
    import galileo_oracle as gorc
    o = gorc.OracleGalileo(orc.OBJECTIVES[obj], lb, ub, population=n_cpu, seed=SEED, stream='philox', invert=(goal == 'max'))
    o.step_fast()                                # warm-up
    c0 = time.perf_counter()
    for _ in range(it_cpu):
        o.step_fast()
    cdt = time.perf_counter() - c0
    return 2 * ((n_cpu + 1) // 2) * it_cpu / cdt, ('oracle/galileo_oracle.py step_fast (numpy), population %d, %d iterations, %.2f s'
                                                    % (n_cpu, it_cpu, cdt))


def run_reference_galileo(args):
    """--impl reference --workload ga_*: the UNMODIFIED reference solver (baseline/_ref behind oracle/galileo_ref_shim.py)
    on one core of this box -- it is single-threaded by construction -- on a bounded population of the same row shape."""
    if int(os.environ.get('RANK', '0')) != 0:
        return None
    obj, D, N, lo, hi, goal = GA_WORKLOADS[args.workload]

    def once(iters):
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'baseline', 'time_reference_galileo.py'), obj, str(D),
                              str(GA_REF_POPULATION), str(iters + 1), str(lo), str(hi), goal], capture_output=True, text=True, timeout=900)
        if out.returncode != 0:
            raise RuntimeError(out.stderr[-500:])
        return json.loads(out.stdout.strip().splitlines()[-1])

    if args.warmup:
        once(1)
    r = once(args.steps)
    if 'unavailable' in r:
        return {'impl': 'reference', 'unavailable': r['unavailable']}
    sample = 'population %d of %d, %d iterations, %.2f s wall (unmodified galileo_oo.py, Python-2 behaviours restored at run time)' % (
        GA_REF_POPULATION, N, r['iterations'], r['wall_s'])
    v = r['entries_per_s']
    return {'impl': 'reference', 'metric': 'ga_entries_per_sec', 'value': v, 'unit': 'entries/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * r['wall_s'] / max(1, r['iterations']),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': ga_describe(args.workload), 'sample': sample},
            'cpu_baseline': {'value': v, 'unit': 'entries/s', 'cores': 1, 'nproc': os.cpu_count(), 'kind': 'reference', 'sample': sample},
            'e2e': {'value': v, 'unit': 'entries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}


def run_galileo(args):
    """--workload ga_*: one step = one iteration of galileo's loop (select, crossover, mutate, replace) over the whole
    population.  metric: entries per second (2*ceil(N/2) entries are built, evaluated, checked for duplicates and
    ranked per iteration).  roofline: against the bytes the generation must move in this design's data layout with
    everything fused -- read the two parents of a pair and write its two entries (16*D per entry) and move the N
    survivors into list order (16*D per survivor) -- so `frac` also shows what the separate hash / objective passes cost;
    `phases_ms` is the per-phase device time from a second, profiled pass."""
    import torch
    from openopt_b200 import GLP, _lib, objectives
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        if int(os.environ.get('RANK', '0')) == 0:
            return {'metric': 'ga_entries_per_sec', 'unavailable': 'the galileo solver is single-GPU (replicas only)'}
        return None
    obj, D, N, lo, hi, goal = GA_WORKLOADS[args.workload]
    lb, ub = lo * np.ones(D), hi * np.ones(D)
    dobj = objectives.get(obj)
    K, W = args.steps, args.warmup
    peak, peak_src = measured_peak()
    torch.cuda.set_device(0)
    entries = 2 * ((N + 1) // 2)
    dev = _lib.DeviceGalileo(dobj.device_id, lb, ub, population=N, seed=SEED, obj_aux=dobj.aux(D), rng='philox',
                             invert=(goal == 'max'), max_batch=max(K, W, 1))
    dev.init()
    if W:
        dev.step(W)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    t0 = dev.timing()
    w0 = time.perf_counter()
    recs, _ = dev.step(K)
    wall = time.perf_counter() - w0
    t1 = dev.timing()
    clocks = sampler.summary()
    ms = t1['total_ms'] - t0['total_ms']
    launches = t1['gpu_launches'] - t0['gpu_launches']
    dev.set_profiling(True)                      # second pass: the same number of iterations, phase by phase
    dev.step(K)
    t2 = dev.timing()
    dev.set_profiling(False)
    phases = {k: v / K for k, v in t2['phase_ms'].items()}
    appended = float(np.mean([r.n_appended for r in recs]))
    dev.close()

    def solve_e2e():
        torch.cuda.synchronize()
        w = time.perf_counter()
        p = GLP(dobj, lb=lb, ub=ub, maxIter=K, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9, iprint=-1, goal=goal)
        r = p.solve('galileo', population=N, seed=SEED, gensPerCall=max(1, min(K, 64)))
        dt = time.perf_counter() - w
        done = len(r.iterValues.f) - 1
        return {'value': entries * done / dt, 'unit': 'entries/s', 'h2d_bytes_per_step': 16.0 * D / max(done, 1),
                'd2h_bytes_per_step': 56 + 8 * D, 'iterations': done, 'wall_s': dt, 'ff': float(r.ff), 'istop': int(r.istop)}

    _lib.release_cached_memory()
    e2e_cold = solve_e2e()
    e2e = solve_e2e()
    # CPU: the numpy oracle (a port) on a bounded population, and the unmodified reference on one core
    port_v, port_sample = ga_cpu_port(obj, lb, ub, goal)
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'baseline', 'time_reference_galileo.py'), obj, str(D), '512', '4',
                              str(lo), str(hi), goal], capture_output=True, text=True, timeout=300)
        ref_py = json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else {'unavailable': out.stderr[-300:]}
    except Exception as e:          # noqa: BLE001
        ref_py = {'unavailable': repr(e)}
    bytes_per_step = entries * 16.0 * D + N * 16.0 * D
    achieved = bytes_per_step / (ms / K * 1e-3) / 1e9
    return {
        'metric': 'ga_entries_per_sec', 'value': entries * K / (ms * 1e-3), 'unit': 'entries/s', 'n_gpus': 1, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': ga_describe(args.workload), 'objective': obj, 'D': D, 'population': N, 'rng': 'philox4x32-10',
                   'l2': 'inputs larger than L2' if N * D * 8 > 2.5e8 else 'population fits L2; reported as is',
                   'appended_per_iteration': appended, 'parallelism': 'single GPU'},
        'roofline': {'bound': 'hbm', 'kernel': 'whole generation (%d launches)' % (launches // max(K, 1)), 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src, 'traffic': None,
                     'bytes_per_launch': bytes_per_step, 'kernel_ms': ms / K, 'phases_ms': phases},
        'cpu_baseline': {'value': port_v, 'unit': 'entries/s', 'cores': 1, 'nproc': os.cpu_count(), 'kind': 'port',
                         'sample': port_sample, 'reference_python': ref_py},
        'e2e': e2e, 'e2e_cold': e2e_cold, 'gpu_launches': int(launches), 'clocks': clocks, 'wall_s_timed_region': wall,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS) + sorted(GA_WORKLOADS))
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints while the run is in progress
    # (NCCL's version banner, warnings) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.workload in GA_WORKLOADS:
            line = run_galileo(args) if args.impl == 'ours' else run_reference_galileo(args)
        else:
            line = run_reference(args) if args.impl == 'reference' else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()

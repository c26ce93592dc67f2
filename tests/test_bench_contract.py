"""bench.py's reference arm prints exactly one JSON line with the contract's keys (CPU; the GPU arm is run
by the driver on a B200)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--workload', 'c2'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['metric'] == 'de_trial_evals_per_sec' and d['unit'] == 'evals/s'
    assert d['dtype'] == 'f64' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['value'] > 0 and 'workload' in d['config']


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                          '--warmup', '1', '--workload', 'c2'], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_galileo_reference_arm_times_the_unmodified_solver():
    """--workload ga_* --impl reference: one contract line, kind "reference" (baseline/_ref, built by __graft_entry__.build())."""
    if not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'openopt')):
        import pytest
        pytest.skip('baseline/_ref is not installed')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '0',
                          '--workload', 'ga_c5_64k'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'ga_entries_per_sec' and d['unit'] == 'entries/s' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] == 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'entries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}

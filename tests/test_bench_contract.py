"""bench.py contract pieces that need no GPU: the reference arm (CPU oracle port timed on the host cores) prints ONE JSON
line with the keys the driver reads, on the same metric / unit / config as the GPU arm."""
import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--ref-n', '300', '--ref-full', '0'], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('logML+gradient evaluations/s at n=20000')
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 0 and d['value'] > 0
    assert d['config']['n'] == 20000 and 'workload' in d['config']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and 'sample' in d['cpu_baseline']
    assert d['e2e'] == dict(value=d['value'], unit='evals/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    # the figure is an extrapolation of a bounded sample and says so in machine-readable form: every phase with its own
    # exponent, the measured sample time kept next to the extrapolated step time
    assert d['extrapolated'] is True and d['n_measured'] == 300 and d['ms_per_step_measured_sample'] < d['ms_per_step']
    sc = d['scale']
    assert sc['exponents_nominal'] == dict(gram=2, chol=3, solve=2, inverse=3, dgram=2)
    for p, e in sc['exponents_nominal'].items():
        want = sc['phases_seconds_measured'][p] * (20000 / 300) ** e
        assert abs(sc['phases_seconds_extrapolated'][p] - want) <= 1e-9 * want
    assert abs(sum(sc['phases_seconds_used'].values()) * 1e3 - d['ms_per_step']) <= 1e-6 * d['ms_per_step']
    assert d['cpu_baseline']['cores'] == (os.cpu_count() or 1) or d['cpu_baseline']['cores'] >= 1


def test_reference_arm_uses_all_cores_under_torchrun_env():
    """ torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every host core """
    env = dict(os.environ, OMP_NUM_THREADS='1', RANK='0', WORLD_SIZE='2', LOCAL_RANK='0')
    res = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                          '--warmup', '0', '--ref-n', '300', '--ref-full', '0'], capture_output=True, text=True,
                         timeout=300, cwd=str(ROOT), env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([l for l in res.stdout.splitlines() if l.startswith('{')][0])
    assert d['cpu_baseline']['cores'] == (os.cpu_count() or 1)


def test_reference_arm_other_ranks_exit_quietly():
    """ under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output """
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    res = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                          '--warmup', '0', '--ref-n', '300', '--ref-full', '0'], capture_output=True, text=True, timeout=120, cwd=str(ROOT),
                         env=env)
    assert res.returncode == 0 and not [l for l in res.stdout.splitlines() if l.startswith('{')]

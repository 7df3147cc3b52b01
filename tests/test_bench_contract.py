"""bench.py's reference arm (the reference algorithm on the host cores: oracle port of model/unet.py + train.py:383-402) runs
without a GPU and prints the JSON line the driver parses: same metric / unit / config keys as the GPU arm."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '4'],
                         capture_output=True, text=True, env=env, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    import bench
    lines = [l for l in _run().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC and d['unit'] == 'STC/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1
    assert d['config']['workload'] == bench.WORKLOAD['net4'] and d['config']['batch_per_gpu'] == 4
    assert d['config'] == bench.config_of('net4', 4, 1, 8)          # key for key what the GPU arm prints at N = 1 (bench.run_ours)
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert d['e2e'] == {'value': d['value'], 'unit': 'STC/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == ''

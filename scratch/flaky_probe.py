"""Repeat the fixture parity test many times in one process to surface run-to-run variation."""
import os, sys, traceback
sys.path.insert(0, '.')
import torch
from tests import test_unet_gpu as T
from tests._util import CONFIGS
gd = os.path.join('tests', 'golden')
fails = 0
reps = int(os.environ.get('REPS', '12'))
for rep in range(reps):
    for name in sorted(CONFIGS):
        for path in ('simt', 'tc'):
            try:
                T.test_train_forward_backward_matches_reference_fixture(name, path, gd)
            except AssertionError as e:
                fails += 1
                print('FAIL rep %d %s %s: %s' % (rep, name, path, str(e)[:600].replace('\n', ' | ')), flush=True)
            except Exception as e:  # noqa: BLE001
                fails += 1
                print('EXC rep %d %s %s: %r' % (rep, name, path, e), flush=True)
print('done: %d failures in %d x %d x 2 runs' % (fails, reps, len(CONFIGS)))

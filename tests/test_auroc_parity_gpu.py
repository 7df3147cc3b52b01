"""AUROC parity (north star: test.py AUROC within +-0.2 % of the reference): the whole train.py -> test.py flow of vec_vad_b200.pipeline
run twice on the same on-disk dataset, same seeds, same batches -- once on the CUDA engine (fp32 SIMT, tf32 and fp16-operand tiles) and
once with the pinned CPU oracle (oracle/unet_oracle.py: the reference's arithmetic) standing in for the network behind the same
pipeline code -- must end at the same frame-level AUROC within 0.002, and the per-frame anomaly scores must rank the frames alike.

No real UCSDped2 frames exist on the box (no network): the dataset is tests/_synthetic_dataset.py's UCSDped2-shaped stand-in, whose
anomalous frames are well separated, so this checks that the pipeline + engine reproduce the reference's decision, not that 97 % is
reached on the real benchmark.
"""
import os

import numpy as np
import pytest
import torch

from oracle import unet_oracle as orc
from tests import _synthetic_dataset as syn
from vec_vad_b200 import pipeline as pl

pytestmark = pytest.mark.gpu


class OracleNet(torch.nn.Module):
    """The CPU oracle behind the surface vec_vad_b200.pipeline drives (CompletionNet's): tensors arrive on the GPU, are computed on
    the CPU in the reference's fp32 arithmetic, results go back."""

    def __init__(self, kind, **kw):
        super().__init__()
        kw = {k: v for k, v in kw.items() if k not in ('patch_size', 'use_tensor_cores')}
        self.m = orc.CompletionNetOracle(kind, **kw)
        self._opt = None

    def to(self, *a, **k):
        return self

    def init_adam(self, **kw):
        self._opt = orc.make_adam(self.m)

    def train_step(self, x, x_of, lambda_raw=1.0, lambda_of=1.0, reduce_grads=None, **kw):
        self.m.train()
        lr_, lo_ = orc.train_step(self.m, self._opt, x.cpu(), x_of.cpu(), lambda_raw, lambda_of)
        return torch.tensor([lr_, lo_], device=x.device)

    @torch.no_grad()
    def score(self, x, x_of):
        self.m.eval()
        r, o = orc.score_cubes(self.m, x.cpu(), x_of.cpu())
        return r.to(x.device), (None if o is None else o.to(x.device))

    def state_dict(self, *a, **k):
        return self.m.state_dict()

    def load_state_dict(self, sd, *a, **k):
        return self.m.load_state_dict(sd)

    def share_workspace(self, pool):
        return self


def _run(root, tag, build):
    """train + test in-process under fixed seeds -> (AUROC, per-frame scores)"""
    import shutil
    for d in ('data', 'results'):
        shutil.rmtree(os.path.join(root, d), ignore_errors=True)
    torch.manual_seed(1234)
    np.random.seed(1234)
    orig = pl.build_network
    pl.build_network = build(orig)
    try:
        pl.train('config.cfg')
        torch.manual_seed(99)
        auc = pl.test('config.cfg', 'results')
    finally:
        pl.build_network = orig
    res = np.load(os.path.join(root, 'results', 'UCSDped2', 'raw2flow_obj_det_with_motion_SelfComplete_frame_results.npz'))
    masks = [torch.load(os.path.join(root, 'results', 'UCSDped2', 'score_mask', str(i)), weights_only=False) for i in range(10)]
    return float(auc), float(res['roc_auc']), np.array([float(np.max(m)) for m in masks])


def test_pipeline_auroc_matches_oracle_pipeline(tmp_path, monkeypatch):
    root = syn.make(str(tmp_path / 'ws'), cfg_overrides={'context_of_num': 0, 'epochs': 4, 'batch_size': 32})
    monkeypatch.chdir(root)

    def oracle_build(orig):
        def build(cfg, **kw):
            st = torch.get_rng_state()
            e = orig(cfg, **kw)                          # only to read the constructor arguments the pipeline resolved
            torch.set_rng_state(st)                      # ... the oracle then draws the SAME initial weights (seeded-init parity)
            kind, ctor = e.kind, dict(e._ctor)
            return OracleNet(kind, **ctor)
        return build

    def engine_build(prec):
        def wrap(orig):
            return lambda cfg, **kw: orig(cfg, **dict(kw, use_tensor_cores=prec))
        return wrap
    auc_o, auc_o2, frames_o = _run(root, 'oracle', oracle_build)
    assert auc_o == auc_o2
    rows = {}
    for prec in (0, 1, 2):
        auc, _, frames = _run(root, 'prec%d' % prec, engine_build(prec))
        rows[prec] = (auc, float(np.corrcoef(frames, frames_o)[0, 1]))
        assert abs(auc - auc_o) <= 0.002, (prec, auc, auc_o, rows)
        # the per-frame anomaly scores (max of the score mask) follow the oracle pipeline's: same ranking of normal vs anomalous frames
        assert rows[prec][1] > 0.98, rows
    assert auc_o > 0.9
    try:                                                     # evidence for profiles/: AUROC and score correlation per operand type
        import json
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
        os.makedirs(out, exist_ok=True)
        json.dump({'oracle_pipeline_auroc': auc_o, 'engine': {('fp32', 'tf32', 'f16')[k]: {'auroc': v[0], 'frame_score_corr': v[1]} for k, v in rows.items()}},
                  open(os.path.join(out, 'auroc_parity.json'), 'w'), indent=1)
    except OSError:
        pass

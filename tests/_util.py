"""Shared helpers for the parity tests (digest format written by tests/golden/make_golden.py)."""
import numpy as np
import torch

N_SAMP = 8

CONFIGS = {
    'net4_flow_b2': ('net4', dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None,
                                  useFlow=True, padding=False)),
    'net4_noflow_b4': ('net4', dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None,
                                    useFlow=False, padding=False)),
    'net4_pad_b2': ('net4', dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None,
                                 useFlow=True, padding=True)),
    'full_b2': ('full', dict(features_root=32, tot_raw_num=5, tot_of_num=5, border_mode='predict', rawRange=None,
                             useFlow=True, padding=False)),
    'net1raw1of_b2': ('1raw1of', dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None,
                                      useFlow=True, padding=False)),
}


def digest(named):
    stats, samp, names = [], [], []
    for k, t in named:
        t = t.detach().to('cpu', torch.float64).reshape(-1)
        n = t.numel()
        idx = (torch.arange(N_SAMP, dtype=torch.int64) * 2654435761 + 12345) % n
        stats.append([t.sum().item(), t.abs().sum().item(), t.pow(2).sum().sqrt().item()])
        samp.append(t[idx].numpy())
        names.append(k)
    return np.array(names), np.array(stats, dtype=np.float64), np.array(samp, dtype=np.float64)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))

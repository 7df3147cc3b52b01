"""Generate golden fixtures from the REAL reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports /root/reference/model/unet.py, vad_datasets.py and utils.py unmodified (with the
``np.int = int`` shim numpy>=1.24 needs) and dumps small .npz fixtures next
to this script.  The fixtures travel to the GPU box; /root/reference does not.
Each UNet fixture holds the synthetic cubes, the seeded-init checksum, the forward 4-tuple,
the two losses, a per-parameter gradient digest, the post-Adam parameter digest after one
and two steps and the BatchNorm running statistics digest, plus eval-mode per-cube scores.
"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
np.int = int  # noqa  (reference uses the removed alias: vad_datasets.py:74, utils.py:22)
sys.path.insert(0, '/root/reference')
sys.path.insert(1, REPO)

from model.unet import SelfCompleteNet4, SelfCompleteNetFull, SelfCompleteNet1raw1of  # noqa: E402  (reference)
import vad_datasets as ref_ds  # noqa: E402  (reference)
import utils as ref_utils  # noqa: E402  (reference)
from oracle.unet_oracle import synthetic_cubes  # noqa: E402  (only the seeded input generator)

N_SAMP = 8


def digest(named):
    """[n,3] (sum, abs-sum, l2) in float64 + [n,N_SAMP] sampled elements at fixed indices."""
    stats, samp, names = [], [], []
    for k, t in named:
        t = t.detach().to(torch.float64).reshape(-1)
        n = t.numel()
        idx = (torch.arange(N_SAMP, dtype=torch.int64) * 2654435761 + 12345) % n
        stats.append([t.sum().item(), t.abs().sum().item(), t.pow(2).sum().sqrt().item()])
        samp.append(t[idx].numpy())
        names.append(k)
    return np.array(names), np.array(stats, dtype=np.float64), np.array(samp, dtype=np.float64)


def unet_fixture(name, cls, kw, batch, t_of, seed_w=0, seed_x=1234, lam=(1.0, 1.0)):
    torch.manual_seed(seed_w)
    torch.set_num_threads(1)   # deterministic reduction order
    model = cls(**kw)
    raw_u8, flow = synthetic_cubes(batch, t_of=t_of, seed=seed_x)
    ds = ref_ds.cube_to_train_dataset(raw_u8, target=flow)
    items = [ds[i] for i in range(batch)]
    x = torch.stack([it[0] for it in items]).float()
    x_of = torch.stack([it[1] for it in items]).float()
    out = {'raw_u8': raw_u8, 'flow': flow, 'x': x.numpy(), 'x_of': x_of.numpy(),
           'lambda': np.array(lam), 'seed_w': seed_w}
    keys = list(model.state_dict().keys())
    out['state_keys'] = np.array(keys)
    n0, s0, e0 = digest(model.state_dict().items())
    out['init_stats'], out['init_samp'] = s0, e0
    use_flow = kw.get('useFlow', True)
    opt = torch.optim.Adam(model.parameters(), eps=1e-7, weight_decay=0.0)
    mse = torch.nn.MSELoss()
    model.train()
    for step in (1, 2):
        of_o, raw_o, of_t, raw_t = model(x, x_of)
        loss_raw = mse(raw_t.detach(), raw_o)
        if use_flow:
            loss_of = mse(of_t.detach(), of_o)
            loss = lam[0] * loss_raw + lam[1] * loss_of
        else:
            loss_of, loss = torch.zeros(()), loss_raw
        opt.zero_grad()
        loss.backward()
        if step == 1:
            out['raw_out'] = raw_o.detach().numpy()
            out['raw_tgt'] = raw_t.detach().numpy()
            if use_flow:
                out['of_out'] = of_o.detach().numpy()
                out['of_tgt'] = of_t.detach().numpy()
            pn, gs, ge = digest([(k, p.grad) for k, p in model.named_parameters()])
            out['param_names'], out['grad_stats'], out['grad_samp'] = pn, gs, ge
            # a few complete gradient tensors (small ones) for element-wise checks
            for k, p in model.named_parameters():
                if k in ('inc0.conv.conv.0.weight', 'inc.conv.conv.0.weight', 'outc0.conv.weight', 'outc.conv.weight',
                         'up03.up.weight', 'up3.up.weight', 'inc0.conv.conv.1.weight', 'inc0.conv.conv.4.bias',
                         'outc_of.conv.weight', 'outc_of0.conv.weight', 'up01.conv.conv.1.bias'):
                    out['grad::' + k] = p.grad.detach().numpy()
        out['loss_raw_%d' % step] = np.float64(loss_raw.item())
        out['loss_of_%d' % step] = np.float64(loss_of.item())
        opt.step()
        _, ps, pe = digest(model.state_dict().items())
        out['state_stats_%d' % step], out['state_samp_%d' % step] = ps, pe
    model.eval()
    with torch.no_grad():
        of_o, raw_o, of_t, raw_t = model(x, x_of)
        out['score_raw'] = ((raw_t - raw_o) ** 2).sum(dim=(1, 2, 3)).numpy()
        if use_flow:
            out['score_of'] = ((of_t - of_o) ** 2).sum(dim=(1, 2, 3)).numpy()
        out['eval_raw_out'] = raw_o.numpy()
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', 'loss', out['loss_raw_1'], out['loss_of_1'])


def dataset_fixture():
    raw_u8, flow = synthetic_cubes(3, t_of=5, seed=77)
    ds = ref_ds.cube_to_train_dataset(raw_u8, target=flow)
    a, b, c = ds[1]
    raw4 = raw_u8[:, 0]            # 4-D input path (vad_datasets.py:132-135)
    ds4 = ref_ds.cube_to_train_dataset(raw4, target=flow[:, 0])
    a4, b4, c4 = ds4[2]
    # integer/index paths
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, size=(5, 3, 60, 90)).astype(np.uint8)
    boxes = np.array([[3.2, 4.7, 40.1, 33.3], [10.0, 12.0, 42.0, 44.0], [0.0, 0.0, 89.5, 59.2], [50.5, 20.5, 70.49, 41.51]],
                     dtype=np.float32)
    fg4 = ref_ds.get_foreground(img, boxes, 32)
    fg3 = ref_ds.get_foreground(img[0], boxes, 32)
    imgf = rng.randn(5, 2, 60, 90).astype(np.float32)
    fgf = ref_ds.get_foreground(imgf, boxes, 32)
    blk = {}
    bb = rng.uniform(0, 1, size=(64, 4))
    for mode in (1, 5, 9):
        res = []
        for r in bb:
            x0, x1 = sorted([r[0] * 360, r[2] * 360])
            y0, y1 = sorted([r[1] * 240, r[3] * 240])
            got = sorted(ref_utils.calc_block_idx(x0, x1, y0, y1, 240 / 3, 360 / 4, mode=mode))
            flat = -np.ones(18, dtype=np.int64)
            flat[:2 * len(got)] = np.array(got).reshape(-1)
            res.append(flat)
        blk['block_idx_mode%d' % mode] = np.array(res)
    path = os.path.join(HERE, 'dataset.npz')
    np.savez_compressed(path, raw_u8=raw_u8, flow=flow, item1_in=a.numpy(), item1_tgt=b.numpy(), item1_copy=c.numpy(),
                        item4_in=a4.numpy(), item4_tgt=b4.numpy(), img=img, imgf=imgf, boxes=boxes, fg4=fg4, fg3=fg3, fgf=fgf,
                        block_boxes=bb, **blk)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def index_fixture():
    """context_range (vad_datasets.py:277-354) of the reference on synthetic video-length lists; -1 rows = it raised."""
    class Obj:
        pass
    out = {}
    layouts = {'a': [10, 7, 12], 'b': [5, 5], 'c': [3, 20], 'd': [9]}
    for lname, lens in layouts.items():
        fvi = []
        for v, n in enumerate(lens):
            fvi += [v + 1] * n
        for mode in ('elastic', 'predict', 'hard'):
            for c in (1, 2, 4):
                o = Obj()
                o.border_mode, o.context_frame_num, o.tot_frame_num, o.frame_video_idx = mode, c, len(fvi), fvi
                width = c + 1 if mode == 'predict' else 2 * c + 1
                res = -np.ones((len(fvi), width), dtype=np.int64)
                for i in range(len(fvi)):
                    try:
                        import io, contextlib
                        with contextlib.redirect_stdout(io.StringIO()):
                            r = ref_ds.ped_dataset.context_range(o, i)
                        res[i] = np.array(r)
                    except NotImplementedError:
                        pass
                out['%s|%s|%d' % (lname, mode, c)] = res
        out['layout_' + lname] = np.array(lens)
    # AUROC of the reference's save_roc_pr_curve_data on seeded synthetic scores / labels
    rng = np.random.RandomState(11)
    labels = (rng.rand(400) > 0.7)
    scores = rng.randn(400) + 1.2 * labels
    import io, contextlib, tempfile
    with contextlib.redirect_stdout(io.StringIO()):
        auc = ref_utils.save_roc_pr_curve_data(scores, labels, os.path.join(tempfile.mkdtemp(), 'r.npz'))
    path = os.path.join(HERE, 'index_paths.npz')
    np.savez_compressed(path, roc_scores=scores, roc_labels=labels, roc_auc=np.float64(auc), **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'index':
        index_fixture()
        sys.exit(0)
    common = dict(features_root=32, tot_raw_num=5, border_mode='predict', rawRange=None)
    unet_fixture('net4_flow_b2', SelfCompleteNet4, dict(common, tot_of_num=1, useFlow=True, padding=False), 2, 1)
    unet_fixture('net4_noflow_b4', SelfCompleteNet4, dict(common, tot_of_num=1, useFlow=False, padding=False), 4, 1)
    unet_fixture('net4_pad_b2', SelfCompleteNet4, dict(common, tot_of_num=1, useFlow=True, padding=True), 2, 1, lam=(1.0, 0.5))
    unet_fixture('full_b2', SelfCompleteNetFull, dict(common, tot_of_num=5, useFlow=True, padding=False), 2, 5)
    unet_fixture('net1raw1of_b2', SelfCompleteNet1raw1of,
                 dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False), 2, 1)
    dataset_fixture()
    index_fixture()

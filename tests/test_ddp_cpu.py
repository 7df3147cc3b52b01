"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes exercise shard bounds, the summing gradient
reduction + 1/world scale (== DataParallel's gradient of the global-batch mean loss), state broadcast and score gather.
The oracle UNet set plays the model: 2 ranks x B/2 cubes with all-reduced gradients must match 1 process x B cubes up to
the BatchNorm statistics being rank-local (checked with BatchNorm in eval mode so the comparison is exact)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vec_vad_b200 import ddp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 128, 129, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = ddp.shard_bounds(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [ddp.shard_bounds(n, r, world)[1] - ddp.shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) <= (n + world - 1) // world


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, local, w = ddp.init_from_env('gloo')
    assert (r, w) == (rank, world)
    from oracle import unet_oracle as orc
    kw = dict(features_root=16, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False)
    torch.manual_seed(100 + rank)                      # deliberately different initial weights per rank
    m = orc.CompletionNetOracle('net4', **kw).eval()   # eval: BatchNorm uses (identical) running stats -> exact sharding identity

    class Flat:                                        # the three tensors broadcast_state() touches
        pass
    flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    holder = Flat()
    holder._pflat, holder._sflat, holder._nbt = flat, torch.zeros(4), torch.zeros(2, dtype=torch.long)
    ddp.broadcast_state(holder, src=0)
    off = 0
    for p in m.parameters():
        p.data.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    raw_u8, flow = orc.synthetic_cubes(6, t_of=1, seed=5)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    b, e = ddp.shard_bounds(6, rank, world)
    mse = torch.nn.MSELoss()
    of_o, raw_o, of_t, raw_t = m(x[b:e], x_of[b:e])
    (mse(raw_t, raw_o) + mse(of_t, of_o)).backward()
    g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    scale = ddp.GradReducer(bucket_bytes=1 << 20)(g)
    g = g * scale
    scores = ((raw_t - raw_o.detach()) ** 2).sum(dim=(1, 2, 3))
    allscores = ddp.gather_scores(scores, 6, rank, world)
    if rank == 0:
        ret['grad'] = g.numpy()
        ret['scores'] = allscores.numpy()
        ret['w0'] = flat.numpy()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_global_batch_gradient():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    from oracle import unet_oracle as orc
    kw = dict(features_root=16, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False)
    torch.manual_seed(100)
    torch.set_num_threads(1)
    m = orc.CompletionNetOracle('net4', **kw).eval()
    np.testing.assert_array_equal(ret['w0'], torch.cat([p.data.reshape(-1) for p in m.parameters()]).numpy())   # rank 0's weights won
    raw_u8, flow = orc.synthetic_cubes(6, t_of=1, seed=5)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    mse = torch.nn.MSELoss()
    of_o, raw_o, of_t, raw_t = m(x, x_of)
    (mse(raw_t, raw_o) + mse(of_t, of_o)).backward()
    g = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).numpy()
    np.testing.assert_allclose(ret['grad'], g, rtol=1e-4, atol=1e-7)
    want = ((raw_t - raw_o.detach()) ** 2).sum(dim=(1, 2, 3)).numpy()
    np.testing.assert_allclose(ret['scores'], want, rtol=1e-5)


# ---------------------------------------------------------------------------------------------------------------------------
# Global batch plan: n not divisible by world * batch.  Every rank must run the same number of steps, the
# ranks' shares of every global batch must partition it, and the weighted gradient sum must equal the global-batch gradient.
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,batch,world', [(9, 4, 4), (10001, 128, 8), (129, 128, 2), (7, 4, 2), (5, 4, 4), (1, 4, 2), (256, 128, 2)])
def test_rank_batch_plan_same_steps_and_partition(n, batch, world):
    plans = [ddp.rank_batch_plan(n, batch, r, world, seed=1234, shuffle=True) for r in range(world)]
    steps = {len(p) for p in plans}
    assert steps == {(n + batch - 1) // batch}                       # identical step count on every rank
    g = torch.Generator().manual_seed(1234)
    order = torch.randperm(n, generator=g)
    seen = []
    for s in range(len(plans[0])):
        glob = order[s * batch:(s + 1) * batch]
        parts = [plans[r][s][0] for r in range(world)]
        assert all(plans[r][s][1] == glob.numel() for r in range(world))
        assert torch.equal(torch.cat(parts), glob)                   # the shares partition the global batch, in order
        seen.append(glob)
    assert sorted(torch.cat(seen).tolist()) == list(range(n))        # one epoch = every cube exactly once


def _worker_ragged(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    ddp.init_from_env('gloo')
    from oracle import unet_oracle as orc
    kw = dict(features_root=16, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False)
    torch.manual_seed(100)
    m = orc.CompletionNetOracle('net4', **kw).eval()
    raw_u8, flow = orc.synthetic_cubes(5, t_of=1, seed=6)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    seed = ddp.shared_seed()
    reducer = ddp.GradReducer()
    mse = torch.nn.MSELoss()
    out = []
    for idx, global_n in ddp.rank_batch_plan(5, 4, rank, world, seed=seed, shuffle=True):   # global batches of 4 and 1: rank 1's last share is empty
        reducer.set_batch(idx.numel(), global_n)
        m.zero_grad()
        if idx.numel():
            of_o, raw_o, of_t, raw_t = m(x[idx], x_of[idx])
            (mse(raw_t, raw_o) + mse(of_t, of_o)).backward()
            g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        else:
            g = torch.zeros(sum(p.numel() for p in m.parameters()))
        out.append((g * reducer(g)).numpy().copy())
    if rank == 0:
        ret['seed'] = seed
        ret['grads'] = out
    dist.destroy_process_group()


def test_ragged_global_batches_weighted_gradient_equals_global_batch_gradient():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_ragged, args=(world, port, ret), nprocs=world, join=True)
    from oracle import unet_oracle as orc
    kw = dict(features_root=16, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False)
    torch.manual_seed(100)
    torch.set_num_threads(1)
    m = orc.CompletionNetOracle('net4', **kw).eval()
    raw_u8, flow = orc.synthetic_cubes(5, t_of=1, seed=6)
    x, x_of = orc.cubes_to_tensors(raw_u8, flow)
    mse = torch.nn.MSELoss()
    plan = ddp.rank_batch_plan(5, 4, 0, 1, seed=ret['seed'], shuffle=True)      # world 1 = the global batches themselves
    assert len(plan) == len(ret['grads']) == 2
    for (idx, _), got in zip(plan, ret['grads']):
        m.zero_grad()
        of_o, raw_o, of_t, raw_t = m(x[idx], x_of[idx])
        (mse(raw_t, raw_o) + mse(of_t, of_o)).backward()
        want = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).numpy()
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------------------
# overlapped (phased) gradient exchange: GradReducer.reduce_phased against the plain whole-buffer sum
class _PhasedStandIn:
    """What reduce_phased needs from CompletionNet (unet.py grad_phase_views / wait_grad_phase), on a CPU buffer: 3 slots of 40
    floats, phases = [24, 40), [16, 24), [0, 16) of every slot."""
    def __init__(self, flat):
        self.flat_grads = flat
        self.waited = []

    def grad_phase_views(self):
        rng = [(24, 40), (16, 24), (0, 16)]
        return [[self.flat_grads[s * 40 + b:s * 40 + e] for s in range(3)] for (b, e) in rng]

    def wait_grad_phase(self, phase, stream=None):
        self.waited.append(phase)


def _worker_phased(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    ddp.init_from_env('gloo')
    g = torch.Generator().manual_seed(7 + rank)
    flat = torch.randn(120, generator=g)
    want = flat.clone()
    dist.all_reduce(want)
    out = {}
    for tag, local_n, global_n in (('even', 4, 8), ('ragged', 3 if rank == 0 else 2, 5)):
        red = ddp.GradReducer(overlap=True)
        red.set_batch(local_n, global_n)
        m = _PhasedStandIn(flat.clone())
        scale = red.reduce_phased(m)
        ref = flat.clone()
        scale_ref = ddp.GradReducer(overlap=False)
        scale_ref.set_batch(local_n, global_n)
        s2 = scale_ref(ref)
        out[tag] = (torch.equal(m.flat_grads, ref), scale, s2, m.waited)
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_phased_reduction_equals_whole_buffer_reduction():
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_phased, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for tag in ('even', 'ragged'):
        same, scale, scale_ref, waited = out[tag]
        assert same, tag
        assert scale == scale_ref
        assert waited == [0, 1, 2]


# ---------------------------------------------------------------------------------------------------------------------------
# sharded optimiser step: reduce-scatter + Adam on 1/world + all-gather == all-reduce + Adam everywhere
class _ShardStandIn:
    """What GradReducer.sharded_step needs from CompletionNet, on CPU buffers, with a plain-torch Adam update of a flat range."""
    def __init__(self, params, grads):
        self.flat_params, self.flat_grads = params, grads
        self.m, self.v, self.step = torch.zeros_like(params), torch.zeros_like(params), 1

    def _adam_flat(self, off, n, scale, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
        sl = slice(off, off + n)
        g = self.flat_grads[sl] * scale
        self.m[sl] = b1 * self.m[sl] + (1 - b1) * g
        self.v[sl] = b2 * self.v[sl] + (1 - b2) * g * g
        self.flat_params[sl] -= (lr / (1 - b1 ** self.step)) * self.m[sl] / (self.v[sl].sqrt() / (1 - b2 ** self.step) ** 0.5 + eps)


def _worker_sharded(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    ddp.init_from_env('gloo')
    g = torch.Generator().manual_seed(3)
    params = torch.randn(128, generator=g)                              # identical on every rank (broadcast_state in real runs)
    grads = torch.randn(128, generator=torch.Generator().manual_seed(50 + rank))
    out = {}
    for tag, local_n, global_n in (('even', 4, 8), ('ragged', 3 if rank == 0 else 2, 5)):
        a = _ShardStandIn(params.clone(), grads.clone())
        red = ddp.GradReducer(shard_optimizer=True)
        red.set_batch(local_n, global_n)
        red.sharded_step(a)
        b = _ShardStandIn(params.clone(), grads.clone())
        ref = ddp.GradReducer()
        ref.set_batch(local_n, global_n)
        b._adam_flat(0, 128, ref(b.flat_grads))
        gathered = [torch.empty(128) for _ in range(world)]
        dist.all_gather(gathered, a.flat_params)
        out[tag] = (float((a.flat_params - b.flat_params).abs().max()), all(torch.equal(gathered[0], t) for t in gathered))
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_optimizer_step_equals_replicated_step():
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for tag in ('even', 'ragged'):
        diff, identical = out[tag]
        assert diff < 1e-6, (tag, diff)          # same update as all-reduce + Adam on every rank ...
        assert identical, tag                    # ... and bit-identical parameters on all ranks afterwards

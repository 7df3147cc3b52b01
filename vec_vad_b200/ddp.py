"""Data-parallel plumbing for the completion-UNet set: one process per GPU, cubes sharded by batch, ONE summing
all-reduce of the flat gradient buffer per step over NCCL (NVLink 5 / NVSwitch), BatchNorm statistics rank-local.

Replaces the reference's single-process ``torch.nn.DataParallel`` (train.py:289,375), which scatters the batch along
dim 0, keeps per-replica BatchNorm statistics and reduce-adds the gradients onto GPU 0.  Same arithmetic here:
every rank runs the full UNet set on B/world cubes, the per-rank losses are means over the local shard, so the
global-batch gradient is the average of the rank gradients = sum-all-reduce followed by 1/world (folded into Adam's
``grad_scale``, include/vecvad.h vecvad_adam_step).  Rank 0's running statistics are the ones saved (DataParallel keeps
replica 0's).
"""
import contextlib
import os

import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """[begin, end) of rank's contiguous shard of n items, chunks of ceil(n/world) like DataParallel's scatter
    (torch.chunk): the LAST shards may be shorter or EMPTY (n=9, world=4 -> 3,3,3,0)."""
    chunk = (n + world - 1) // world
    b = min(n, rank * chunk)
    return b, min(n, b + chunk)


def rank_batch_plan(n, batch_size, rank, world, seed=None, shuffle=True):
    """This rank's share of every GLOBAL batch of one epoch over n cubes: [(index tensor (may be empty), global batch size)].

    One permutation per epoch from a seed every rank shares (``shared_seed``), cut into global batches of ``batch_size``
    (the last one ragged, like ``DataLoader(shuffle=True)``, train.py:373), each split over the ranks with ``shard_bounds``
    exactly as ``nn.DataParallel`` scatters a batch (train.py:375).  Every rank therefore runs the SAME number of steps
    (ceil(n / batch_size)) whatever n is, so the per-step gradient all-reduces always pair up; a rank whose share of a ragged
    last batch is empty still takes part with a zero gradient."""
    if shuffle:
        g = torch.Generator().manual_seed(int(seed))
        order = torch.randperm(n, generator=g)
    else:
        order = torch.arange(n)
    plan = []
    for i in range(0, n, batch_size):
        glob = order[i:i + batch_size]
        b, e = shard_bounds(glob.numel(), rank, world)
        plan.append((glob[b:e], int(glob.numel())))
    return plan


def shared_seed(group=None):
    """A fresh shuffle seed drawn on rank 0 (RandomSampler's own seeding) and broadcast, so all ranks cut the same batches."""
    seed = torch.empty((), dtype=torch.int64).random_()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = seed.reshape(1)
        if dist.get_backend(group) == 'nccl':
            t = t.cuda()
        dist.broadcast(t, src=0, group=group)
        seed = t.cpu()[0]
    return int(seed.item())


def init_from_env(backend=None):
    """torchrun / torch.distributed.run rendezvous (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def _all_reduce_views(views, group=None):
    """Sum every tensor of ``views`` over the ranks, asynchronously on the current stream, in ONE coalesced collective where the
    backend offers it (NCCL: one ncclGroupStart/End, one kernel) and one collective per view otherwise.  -> list of work handles."""
    try:
        from torch.distributed.distributed_c10d import _coalescing_manager
        with _coalescing_manager(group=group, async_ops=True) as cm:
            for v in views:
                dist.all_reduce(v, group=group)
        return [cm]
    except (ImportError, RuntimeError, TypeError, AttributeError):
        return [dist.all_reduce(v, group=group, async_op=True) for v in views]


class GradReducer:
    """Callable handed to ``CompletionNet.train_step(reduce_grads=...)``: sums the flat gradient buffer over the ranks
    in place on the current stream and returns the scale Adam must apply (1/world).

    overlap=True exchanges the three gradient phases while the backward still runs (``reduce_phased``).  Measured on 2 x B200
    (profiles/r02_ddp_overlap_n2.txt) it buys nothing at NCCL's default channel count -- the step keeps every SM busy, so the
    collective's SM time is paid either way (2.61 ms overlapped vs 2.60 ms after the backward, 2.46 ms on one GPU) -- and wins
    only when the collective is throttled to few CTAs (NCCL_MAX_CTAS <= 16); hence off by default."""

    def __init__(self, group=None, bucket_bytes=0, overlap=False, shard_optimizer=False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bucket_elems = bucket_bytes // 4
        self.pre_weight = None
        self.overlap = overlap
        # sharded optimiser step (``sharded_step``): same bytes on the wire as the all-reduce (reduce-scatter + all-gather), but every
        # rank runs Adam on 1/world of the parameters instead of all of them
        self.shard_optimizer = bool(shard_optimizer) and self.world > 1
        self._comm = None                 # side stream the phase exchanges are queued behind (created on first use)

    def set_batch(self, local_n, global_n):
        """Tell the reducer how this step's global batch was split.  The global-batch mean gradient is
        sum_r (local_n_r / global_n) * grad_r (each rank's loss is a mean over ITS cubes): with equal shares that is the plain
        sum times 1/world (folded into Adam); with a ragged split every rank weights its gradient before the sum."""
        self.pre_weight = None if global_n % self.world == 0 else float(local_n) / float(global_n)

    def __call__(self, flat):
        if self.world == 1:
            return 1.0
        pre = self.pre_weight
        if pre is not None:
            flat.mul_(pre)
        scale = 1.0 / self.world if pre is None else 1.0
        return self._sum(flat, scale)

    def sharded_step(self, model):
        """The optimiser step of a data-parallel rank without replicating Adam: sum-reduce-scatter of the flat gradient buffer (each
        rank ends up with the summed gradients of ITS contiguous 1/world of the buffer, in place), Adam on that shard with
        grad_scale 1/world, all-gather of the updated parameters (in place).  ``model``: CompletionNet (flat_params / flat_grads /
        _adam_flat); its ``_adam['step']`` is already incremented.  The flat buffers are a multiple of 64 floats per slot, so they
        split evenly over 2 / 4 / 8 ranks.  Every rank's parameters are bit-identical afterwards (they all receive the same shards)."""
        flat_g, flat_p = model.flat_grads, model.flat_params
        n = flat_g.numel()
        if n % self.world:
            raise RuntimeError('sharded_step: %d parameters do not split over %d ranks' % (n, self.world))
        shard = n // self.world
        lo = self.rank * shard
        pre = self.pre_weight
        if pre is not None:
            flat_g.mul_(pre)
        dist.reduce_scatter_tensor(flat_g[lo:lo + shard], flat_g, group=self.group)
        model._adam_flat(lo, shard, 1.0 / self.world if pre is None else 1.0)
        dist.all_gather_into_tensor(flat_p, flat_p[lo:lo + shard], group=self.group)

    def reduce_phased(self, model):
        """Called by ``CompletionNet.train_step`` right after the (asynchronous) backward was queued: exchange every gradient
        phase as soon as the backward has produced it, while the rest of the backward still runs.

        ``model.grad_phase_views()`` gives, per phase, one contiguous view of the flat gradient buffer per UNet slot;
        ``model.wait_grad_phase(p, stream)`` makes a stream wait for phase p.  Each phase's views are summed over the ranks in one
        coalesced NCCL call queued behind a side stream that waits for the phase; the caller's stream finally waits for all
        three.  Returns the scale Adam must apply, like ``__call__``.  Without overlap (``overlap=False``) or with one rank this
        is ``__call__`` on the flat buffer."""
        if self.world == 1:
            return 1.0
        if not self.overlap:
            return self(model.flat_grads)
        pre = self.pre_weight
        on_gpu = model.flat_grads.is_cuda          # (the CPU / gloo tests drive the same code with a stand-in model)
        if on_gpu:
            cur = torch.cuda.current_stream()
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=model.flat_grads.device)
        works = []
        for ph, views in enumerate(model.grad_phase_views()):
            with (torch.cuda.stream(self._comm) if on_gpu else contextlib.nullcontext()):
                model.wait_grad_phase(ph, self._comm)
                if pre is not None:
                    for v in views:
                        v.mul_(pre)
                works.extend(_all_reduce_views(views, self.group))
        for w in works:
            w.wait()                       # NCCL: the caller's stream waits for the collective, the host does not block
        if on_gpu:
            cur.wait_stream(self._comm)
        return 1.0 / self.world if pre is None else 1.0

    def _sum(self, flat, scale):
        if self.bucket_elems and flat.numel() > self.bucket_elems:
            # bucketed: lets NCCL pipeline the buckets; each is contiguous in the flat buffer
            works = [dist.all_reduce(flat[o:o + self.bucket_elems], group=self.group, async_op=True)
                     for o in range(0, flat.numel(), self.bucket_elems)]
            for w in works:
                w.wait()
        else:
            dist.all_reduce(flat, group=self.group)
        return scale


def broadcast_state(model, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters, statistics and step counters (DataParallel replicates the
    module from GPU 0 every step; here it is done once, the all-reduced gradients keep the replicas identical)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if hasattr(model, '_flush_nbt'):
        model._flush_nbt()                  # step counters still pending on the host go into the buffer that is broadcast
    for t in (model._pflat, model._sflat, model._nbt):
        dist.broadcast(t, src=src, group=group)


def gather_scores(local_scores, n_total, rank, world, group=None):
    """Concatenate per-rank score vectors (shards from ``shard_bounds``) in rank order on every rank."""
    if world == 1:
        return local_scores
    chunk = (n_total + world - 1) // world
    pad = torch.zeros(chunk, dtype=local_scores.dtype, device=local_scores.device)
    pad[:local_scores.numel()] = local_scores
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(world):
        b, e = shard_bounds(n_total, r, world)
        parts.append(out[r][:e - b])
    return torch.cat(parts)

#!/usr/bin/env python
"""bench.py -- STCs/sec of one train step of the completion-UNet set (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 30 --warmup 5            (ours; N>1 is launched by torch.distributed.run)
    python bench.py --impl reference --steps 3 --warmup 1     (the reference algorithm on the host cores)

A step = cube staging (uint8 cubes -> float tensors) + forward of every UNet + MSE losses + backward +
(N>1: one NCCL all-reduce of the flat gradient buffer) + Adam, for one batch of 128 synthetic 5x32x32x3 cubes
per GPU (BASELINE.json configs[1]: UCSDped2 5raw1of dual-UNet set, batch 128).
  value : device-timed (CUDA events on the launch stream), cubes already resident in HBM as uint8.
  e2e   : the same step through the public module API with HOST (pinned) cube buffers: H2D of the cubes and
          D2H of the two losses inside the timed region, every step.
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# DRAM traffic of the contraction kernel classes over ONE train step (batch 128, 5raw1of): dram__bytes_read.sum + dram__bytes_write.sum
# summed over the class's launches (k_igemm_flat + k_igemm_tc3, k_wgrad_flat + k_wgrad_tc2) in the per-launch ncu pass committed under
# profiles/ (NCU_SOURCE; tabulated by profiles/step_table.py): {class: {operand type: (bytes per step, launches per step)}}
NCU_DRAM_BYTES_PER_STEP = {'conv_dgrad_tcgen05': {'tf32': (1.876e9 + 0.604e9, 33), 'f16': (0.939e9 + 0.122e9, 33)},
                           'wgrad_tcgen05': {'tf32': (1.924e9 + 0.002e9, 17), 'f16': (0.992e9 + 0.002e9, 17)}}
NCU_SOURCE = 'profiles/r01_v5_step_metrics.csv (tf32), profiles/r02_f16_step_metrics.csv (f16)'

METRIC = 'STCs/sec (train step, device-timed)'       # BASELINE.json's metric; both arms print the same string

# algorithmic FLOPs of one train step per STC (fwd + dgrad + wgrad of every conv; SURVEY.md section 8a / BASELINE.md section 2)
FLOPS_PER_STC = {'net4': 5.524e9, 'full': 9.206e9, 'noflow': 4.604e9}
NET_KW = {
    'net4': ('net4', dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=True, padding=False), 1),
    'full': ('full', dict(features_root=32, tot_raw_num=5, tot_of_num=5, border_mode='predict', rawRange=None, useFlow=True, padding=False), 5),
    'noflow': ('net4', dict(features_root=32, tot_raw_num=5, tot_of_num=1, border_mode='predict', rawRange=None, useFlow=False, padding=False), 1),
}
WORKLOAD = {'net4': 'UCSDped2 raw+flow dual-UNet set (5raw1of, SelfCompleteNet4), synthetic 5x32x32x3 cubes',
            'full': 'avenue 5raw5of (SelfCompleteNetFull, context_of_num=4), synthetic 5x32x32x3 cubes',
            'noflow': 'UCSDped2 appearance-only UNet set (useFlow=False), synthetic 5x32x32x3 cubes'}


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def cpu_reference_steps(net, batch, steps, warmup, threads, pool=1):
    """The reference algorithm (oracle port of model/unet.py + train.py:383-402) on the host cores. -> (STC/s, s/step)"""
    import torch
    from oracle import unet_oracle as orc
    kind, kw, t_of = NET_KW[net]
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = orc.CompletionNetOracle(kind, **kw).train()
    opt = orc.make_adam(m)
    data = []
    for i in range(max(1, pool)):
        raw_u8, flow = orc.synthetic_cubes(batch, t_of=t_of, seed=1234 + i)
        data.append(orc.cubes_to_tensors(raw_u8, flow))
    for i in range(warmup):
        orc.train_step(m, opt, *data[i % len(data)])
    t0 = time.perf_counter()
    for i in range(steps):
        orc.train_step(m, opt, *data[(warmup + i) % len(data)])
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt


def config_of(net, batch, world, pool):
    """The workload description both arms print: identical key for key (the driver compares them)."""
    return {'workload': WORKLOAD[net], 'net': net, 'batch_per_gpu': batch, 'global_batch': world * batch, 'input_pool_batches': pool,
            'l2': 'per-step working set (activations + gradients, >3 GB at batch 128) exceeds the 126 MB L2; inputs rotate over %d batches' % pool}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.batch
    v, dt = cpu_reference_steps(args.net, batch, args.steps, args.warmup, threads, args.pool)
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'STC/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': config_of(args.net, batch, max(1, args.gpus), args.pool),      # the workload of the arm it stands beside (N x batch)
            'details': {'timing': 'host wall clock on rank 0 only (the CPU arm has no device): each step is one batch of %d cubes of the same '
                                  'workload, a bounded sample of the N-GPU global batch' % batch},
            'cpu_baseline': {'value': v, 'unit': 'STC/s', 'cores': threads, 'kind': 'port',
                             'sample': '%d train steps of batch %d (oracle port of model/unet.py + train.py:383-402, torch CPU fp32)' % (args.steps, batch)},
            'e2e': {'value': v, 'unit': 'STC/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def measure_net(args, net, steps, warmup, with_e2e, sample_clocks):
    """Device-timed (and optionally end-to-end) train steps of one UNet set on this rank's GPU.  Collective when WORLD_SIZE > 1."""
    import torch
    import torch.distributed as dist
    from vec_vad_b200 import _lib, unet as vu, vad_datasets as vd

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local)
    kind, kw, t_of = NET_KW[net]
    cls = {'net4': vu.SelfCompleteNet4, 'full': vu.SelfCompleteNetFull}[kind]
    torch.manual_seed(0)
    prec = 0 if args.simt else {'tf32': 1, 'f16': 2}[args.precision]
    model = cls(use_tensor_cores=prec, **kw).cuda().train()
    model.init_adam(lr=1e-3, eps=1e-7)
    B, P = args.batch, args.pool
    g = torch.Generator().manual_seed(1234 + rank)
    host_raw = [torch.randint(0, 256, (B, 5, 32, 32, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(P)]
    host_flow = [torch.randn((B, t_of, 32, 32, 2), generator=g).pin_memory() for _ in range(P)]
    dev_raw = [t.to(dev) for t in host_raw]
    dev_flow = [t.to(dev) for t in host_flow]
    losses = torch.zeros(2, device=dev)
    reduce = None
    if world > 1:
        # NCCL sum over NVLink/NVSwitch, 1/world folded into Adam: one all-reduce of the flat buffer after the backward; --overlap:
        # each gradient phase (decoder / deepest encoder block / rest) is exchanged on a side stream as soon as it is final
        from vec_vad_b200 import ddp
        reduce = ddp.GradReducer(overlap=args.overlap, shard_optimizer=not (args.no_shard_optimizer or args.overlap))

    def step_dev(i):
        x, x_of = vd.cubes_to_device_tensors(dev_raw[i % P], dev_flow[i % P])
        model.train_step(x, x_of, 1.0, 1.0, losses=losses, reduce_grads=reduce)

    feeder = vd.HostCubeFeeder(list(zip(host_raw, host_flow)), device=dev) if with_e2e else None

    def step_e2e(i):
        # every step: one H2D copy of a batch of pinned host cubes (queued one batch ahead on the feeder's copy stream, so it
        # overlaps the previous step) and one D2H read of the two losses (a host sync)
        x, x_of = feeder.next()
        return model.train_step(x, x_of, 1.0, 1.0, reduce_grads=reduce).cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.lib().vecvad_launch_count()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, (_lib.lib().vecvad_launch_count() - n0) // steps

    clk = ClockSampler(local)
    if rank == 0 and sample_clocks:
        clk.start()
    ms_dev, launches = timed(step_dev, steps, warmup)
    clocks = clk.stop() if (rank == 0 and sample_clocks) else None
    ms_e2e = timed(step_e2e, steps, max(3, warmup // 2))[0] if with_e2e else None
    final_loss = losses.cpu().tolist()
    # ---- per-kernel-class breakdown (separate, untimed pass; CUDA events on the launch stream around every launch)
    prof = None
    psteps = min(5, steps)
    if rank == 0:
        _lib.profile_begin()
    for i in range(psteps):            # every rank steps (the gradient all-reduce is collective); rank 0 records
        step_dev(i)
    if rank == 0:
        prof = {k: (ms / psteps, fl / psteps, la // psteps) for k, (ms, fl, la) in _lib.profile_end().items()}
    barrier()
    out = dict(ms_dev=ms_dev, launches=int(launches), ms_e2e=ms_e2e, final_loss=final_loss, prof=prof, clocks=clocks,
               h2d=int(host_raw[0].numel() + 4 * host_flow[0].numel()), world=world, batch=B)
    del model
    torch.cuda.empty_cache()
    return out


def operand_peak(args, pk):
    """Tensor-pipe peak of the operand type the tiles run in: fp16 operands run at the bf16 rate (MEASURED_PEAKS.json, sustained);
    tf32 at the tf32 cuBLAS rate measured the same way on this pool's B200 (profiles/r02_tf32_peak.json, scratch/measure_tf32_peak.py)."""
    peak16 = pk['bf16_tflops_sustained'] if 'bf16_tflops_sustained' in pk else pk['bf16_tflops']
    if args.simt or args.precision == 'f16':
        return peak16, peak16, 'bf16/fp16 rate'
    tf = os.path.join(REPO, 'profiles', 'r02_tf32_peak.json')
    tf32 = json.load(open(tf))['tf32_tflops_sustained'] if os.path.exists(tf) else peak16 / 2
    return peak16, tf32, 'tf32 rate measured with cuBLAS (profiles/r02_tf32_peak.json)'


def roofline_of(args, net, m, pk, src):
    prof = m['prof']
    world, B = m['world'], m['batch']
    value = world * B / (m['ms_dev'] * 1e-3)
    step_tf = value * FLOPS_PER_STC[net] / 1e12 / world           # per-GPU rate of the step's contractions over the WHOLE step
    peak16, peak_op, peak_how = operand_peak(args, pk)
    dom = max((k for k in prof if prof[k][1] > 0), key=lambda k: prof[k][0])
    dom_ms, dom_fl, dom_n = prof[dom]
    dom_tf = dom_fl / (dom_ms * 1e-3) / 1e12
    traffic = None
    if dom in NCU_DRAM_BYTES_PER_STEP and net == 'net4' and B == 128 and not args.simt and args.precision in NCU_DRAM_BYTES_PER_STEP[dom]:
        tot, n = NCU_DRAM_BYTES_PER_STEP[dom][args.precision]
        traffic = tot / n
    return {'bound': 'tensor', 'kernel': dom, 'achieved': dom_tf, 'peak': peak16, 'unit': 'TFLOP/s', 'frac': dom_tf / peak16,
            'traffic': traffic, 'traffic_note': 'mean DRAM bytes per launch of this kernel class (ncu, %s)' % NCU_SOURCE,
            'launches_per_step': int(dom_n), 'ms_per_step': dom_ms, 'share_of_step': dom_ms / m['ms_dev'],
            'peak_of_operand_type': peak_op, 'frac_of_operand_type_peak': dom_tf / peak_op, 'operand_peak_is': peak_how,
            'whole_step_tflops': step_tf, 'whole_step_frac': step_tf / peak16,
            'note': 'achieved = ALGORITHMIC conv FLOPs of the dominant kernel class (transposed convs counted at their 9 taps) / its '
                    'CUDA-event time (sum over its launches in one step, measured live in a separate pass); peak = %s sustained bf16 '
                    'cuBLAS (%s); the weight-gradient tiles run concurrently on a side stream, so class times overlap and include '
                    'their mutual contention' % (src, 'MEASURED_PEAKS.json' if src == 'measured' else 'B200_PROFILING.md fallback')}


def flow_secondary(pk):
    """BASELINE.json configs[4] (FlowNet2 correlation + warp, 1024x436 pairs) beside the headline: bench_flow.py's rows.  The warp
    kernels are judged against the HBM roofline over their algorithmic bytes; the correlation (59 FLOP/B) against the fp32 FMA rate
    (148 SMs x 128 lanes x 2 x SM clock: no fp32 peak is in MEASURED_PEAKS.json); then the whole FlowNet2 stack on a 512x384 pair."""
    import bench_flow
    fp32_peak = 148 * 128 * 2 * pk.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
    rows = []
    for batch in (1, 8):
        for r in bench_flow.run(batch, 10, with_reference=False):
            roof = {'bound': 'hbm', 'achieved': r['achieved_gbs'], 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': r['frac_of_hbm_peak'],
                    'traffic': None, 'algorithmic_bytes': r['algorithmic_bytes']}
            if r.get('tflops_fp32'):
                roof = {'bound': 'fp32', 'achieved': r['tflops_fp32'], 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': r['tflops_fp32'] / fp32_peak,
                        'traffic': None, 'algorithmic_bytes': r['algorithmic_bytes'], 'frac_of_hbm_peak': r['frac_of_hbm_peak'],
                        'peak_is': 'nominal fp32 FMA rate (SMs x lanes x 2 x max SM clock); a register-only FMA loop measures 72.5 (FFMA) / 74.0 (FFMA2) TFLOP/s on this pool, profiles/r02_ffma_rate.txt'}
            rows.append({'workload': 'FlowNet2 %s, 1024x436 synthetic pairs, batch %d (BASELINE.json configs[4])' % (r['op'], batch),
                         'metric': 'pairs/sec', 'value': r['pairs_per_s'], 'unit': 'pairs/s', 'us_per_launch': r['us'], 'roofline': roof})
    f = bench_flow.run_flownet2(1, 5)
    rows.append({'workload': 'FlowNet2 full stack (5 networks, 162.5 M parameters), one 512x384 pair (calc_optical_flow.py:49-57)', 'metric': 'pairs/sec',
                 'value': f['pairs_per_s'], 'unit': 'pairs/s', 'ms_per_pair': f['ms'],
                 'roofline': {'bound': 'fp32', 'achieved': f['tflops_fp32'], 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': f['tflops_fp32'] / fp32_peak,
                              'traffic': None, 'algorithmic_gflop': f['algorithmic_gflop'],
                              'peak_is': 'nominal fp32 FMA rate (SMs x lanes x 2 x max SM clock); a register-only FMA loop measures 72.5 (FFMA) / 74.0 (FFMA2) TFLOP/s on this pool, profiles/r02_ffma_rate.txt'}})
    return rows


def epoch_secondary(args):
    """SURVEY.md section 8d's epoch-shaped run: the real UCSDped2 training-box count (31 089 STCs) resident in HBM as uint8 cubes,
    one epoch of shuffled batches of 128 through DeviceCubeStore + train_step (243 steps, the last one ragged: 113 cubes)."""
    import torch
    from vec_vad_b200 import unet as vu, vad_datasets as vd
    n = 31089
    g = torch.Generator().manual_seed(7)
    raw = torch.randint(0, 256, (n, 5, 32, 32, 3), generator=g, dtype=torch.uint8)
    flow = torch.randn((n, 1, 32, 32, 2), generator=g)
    store = vd.DeviceCubeStore(raw, flow)
    kind, kw, _ = NET_KW['net4']
    torch.manual_seed(0)
    prec = 0 if args.simt else {'tf32': 1, 'f16': 2}[args.precision]
    model = vu.SelfCompleteNet4(use_tensor_cores=prec, **kw).cuda().train()
    model.init_adam(lr=1e-3, eps=1e-7)
    losses = torch.zeros(2, device='cuda')

    def epoch(seed):
        steps = 0
        for x, x_of in store.batches(args.batch, shuffle=True, generator=torch.Generator().manual_seed(seed)):
            model.train_step(x, x_of, 1.0, 1.0, losses=losses)
            steps += 1
        return steps
    epoch(0)                                             # warm-up epoch (workspace, allocator, both batch shapes)
    torch.cuda.synchronize()
    times = []
    for seed in (1, 2):                                  # two timed epochs; the faster one is the value (the host issues ~100 launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # per step and a busy box shows there first)
        e0.record()
        steps = epoch(seed)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    del model, store
    torch.cuda.empty_cache()
    return {'workload': 'UCSDped2-shaped epoch: %d synthetic STCs resident in HBM, shuffled batches of %d through DeviceCubeStore + train_step '
                        '(%d steps, ragged last batch)' % (n, args.batch, steps), 'metric': 'STCs/sec (one training epoch, device-timed)',
            'value': n / (ms * 1e-3), 'unit': 'STC/s', 'ms_per_epoch': ms, 'ms_per_epoch_all': times, 'steps': steps}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    m = measure_net(args, args.net, args.steps, args.warmup, with_e2e=True, sample_clocks=True)
    B, P = args.batch, args.pool
    if rank == 0:
        pk, src = peaks()
        value = world * B / (m['ms_dev'] * 1e-3)
        line = {'metric': METRIC, 'value': value, 'unit': 'STC/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': m['ms_dev'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': ('f32' if args.simt else args.precision), 'data': 'synthetic',
                'config': config_of(args.net, B, world, P),
                'details': {'parallelism': 'dp%d' % world,
                            'grad_exchange': ('none' if world == 1 else 'three gradient phases exchanged while the backward runs' if args.overlap
                                              else 'one all-reduce after backward, Adam on every rank' if args.no_shard_optimizer
                                              else 'reduce-scatter after backward, Adam on 1/N of the parameters, all-gather of the parameters'),
                            'operands': ('fp32 SIMT' if args.simt else 'tcgen05 kind::%s operands, fp32 accumulation; BatchNorm, losses, 1x1 output '
                                         'conv and Adam in fp32' % args.precision),
                            'final_losses': m['final_loss']},
                'clocks': m['clocks'], 'gpu_launches': m['launches'],
                'e2e': {'value': world * B / (m['ms_e2e'] * 1e-3), 'unit': 'STC/s', 'ms_per_step': m['ms_e2e'],
                        'h2d_bytes_per_step': m['h2d'], 'd2h_bytes_per_step': 8},
                'roofline': roofline_of(args, args.net, m, pk, src),
                'kernel_classes_ms_per_step': {k: round(v[0], 4) for k, v in m['prof'].items() if v[2] > 0}}
    # ---- secondary workloads (BASELINE.json configs[2] and configs[4]) so the driver's record carries them; 1 GPU only
    if world == 1 and not args.no_secondary:
        sec = []
        if args.net != 'full':
            m2 = measure_net(args, 'full', min(args.steps, 20), 5, with_e2e=False, sample_clocks=False)
            sec.append({'workload': WORKLOAD['full'] + ', batch %d per GPU (BASELINE.json configs[2] is 2 x this)' % B, 'metric': METRIC,
                        'value': B / (m2['ms_dev'] * 1e-3), 'unit': 'STC/s', 'ms_per_step': m2['ms_dev'], 'gpu_launches': m2['launches'],
                        'roofline': roofline_of(args, 'full', m2, pk, src),
                        'kernel_classes_ms_per_step': {k: round(v[0], 4) for k, v in m2['prof'].items() if v[2] > 0}})
        try:
            sec.append(epoch_secondary(args))
        except Exception as e:
            sec.append({'workload': 'UCSDped2-shaped epoch', 'error': repr(e)})
        try:
            sec += flow_secondary(pk)
        except Exception as e:                                  # the headline line must still print
            sec.append({'workload': 'FlowNet2 ops', 'error': repr(e)})
        line['secondary'] = sec
    if rank == 0:
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            v, dt = cpu_reference_steps(args.net, B, 4, 1, threads)
            line['cpu_baseline'] = {'value': v, 'unit': 'STC/s', 'cores': threads, 'kind': 'port',
                                    'sample': '4 train steps of batch %d after 1 warm-up (oracle port of model/unet.py + train.py:383-402, torch CPU fp32)' % B}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--net', default='net4', choices=sorted(NET_KW))
    ap.add_argument('--batch', type=int, default=128, help='cubes per GPU per step')
    ap.add_argument('--pool', type=int, default=8, help='distinct input batches rotated through')
    ap.add_argument('--simt', action='store_true', help='fp32 SIMT tiles instead of tcgen05 tiles')
    ap.add_argument('--precision', default='f16', choices=['tf32', 'f16'], help='operand type of the tcgen05 tiles (fp32 accumulation either way)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-secondary', action='store_true', help='skip the secondary workloads (5raw5of set, FlowNet2 ops)')
    ap.add_argument('--no-shard-optimizer', action='store_true', help='N>1: all-reduce + full Adam on every rank instead of reduce-scatter + '
                    'Adam on 1/N of the parameters + all-gather of the parameters')
    ap.add_argument('--overlap', action='store_true', help='N>1: exchange the three gradient phases while the backward runs instead of one '
                    'all-reduce after it (measured: no gain while the step saturates the SMs, profiles/r02_ddp_overlap_n2.txt)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

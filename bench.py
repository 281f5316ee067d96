#!/usr/bin/env python
"""Benchmark of the multislice hot path (BASELINE.json metric: multislice Gpixel*slice/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|headline|config1] [--impl reference]

One "step" = forward multislice + |psi| loss + adjoint gradient over one synthetic field per GPU.
Default workload (N=1) is BASELINE.json configs[1]: random delta/beta phantom 2048 x 2048 x 256,
forward + adjoint on one B200.  1 unit = one pixel advanced through one slice (forward + adjoint
counts the slice once).  With N>1 ranks (torchrun) every rank owns one such field (one projection
angle of the data-parallel reconstruction, weak scaling) and the mean of the object gradient over ranks
is formed every step, overlapped with the adjoint sweep (--exchange auto: copy engines over NVLink peer
memory between two GPUs, NCCL all-reduce for three or more).

Keys of the JSON line: see the contract in the task description; in short
  value        device-timed whole-job throughput, inputs resident in HBM
  e2e          the same step through the public API (FullfieldObjective.step) with this step's
               measured projections copied from pinned host memory and the loss read back
  roofline     dominant kernel (sweep_kernel, forward + adjoint instantiations, launch-weighted): algorithmic
               bytes / in-situ CUDA-event duration vs measured HBM peak; per direction under "kernels"
  cpu_baseline oracle port (NumPy complex128, the reference algorithm) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # name: (ny, nx, n_slice, description)
    'config2': (2048, 2048, 256, 'random delta/beta phantom 2048x2048x256, forward + adjoint gradient (BASELINE configs[1])'),
    'headline': (4096, 4096, 512, 'random delta/beta phantom 4096x4096x512, forward + adjoint gradient (north_star target size)'),
    'config1': (512, 512, 100, 'zone-plate-sized 512x512x100 field, forward + adjoint gradient (BASELINE configs[0] shape)'),
    'small': (256, 256, 16, 'tiny functional check'),
}
ENERGY_EV, PSIZE_CM = 5000, 1e-7
# algorithmic bytes per pixel*slice of each pass: complex64 field, fp32 (delta,beta)
PASS_BYTES = {'row_conv_transmit': 24, 'col_conv': 16, 'row_conv_adjoint': 40,
              # sweep kernels: one slice per launch; contract figures of SURVEY 8d (two-pass model): forward 40, adjoint 56
              'sweep_forward': 40, 'sweep_adjoint': 56}
STEP_BYTES = 96


def auto_sm_reserve(B, ny, nx, n_sm=148, max_reserve=24):
    """largest number of SMs (<= max_reserve) that can be left to NCCL without adding a round of tiles to the
    sweep kernels (8 lines per tile up to 2048-long lines, 4 for 4096)"""
    def rounds(lines, length, ctas):
        lpc = 4 if length >= 4096 else (8 if length >= 256 else 16)
        tiles = B * lines // lpc
        return -(-tiles // ctas)
    best = 0
    for r in range(0, max_reserve + 1):
        if rounds(ny, nx, n_sm - r) == rounds(ny, nx, n_sm) and rounds(nx, ny, n_sm - r) == rounds(nx, ny, n_sm):
            best = r
    return best


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smax)), 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (NumPy complex128) on a bounded sample
# ------------------------------------------------------------------------------------------------
def _cpu_sample(args):
    ny, nx, nz, seed = args
    from oracle import multislice_oracle as mo       # the one place bench.py executes oracle/: the measured CPU baseline
    gd, gb = mo.random_phantom((1, ny, nx, nz), seed=seed)
    gd = gd.astype(np.float64); gb = gb.astype(np.float64)
    one, zero = np.ones((ny, nx)), np.zeros((ny, nx))
    target = np.full((1, ny, nx), 0.98)
    t0 = time.perf_counter()
    mo.loss_and_grad(gd, gb, one, zero, ENERGY_EV, PSIZE_CM, target)
    return time.perf_counter() - t0


def cpu_baseline(ny, nx, procs, sample_slices):
    """forward + adjoint of the oracle on [1, ny, nx, sample_slices], `procs` independent processes
    (the reference's parallel model is one MPI rank per batch element; NumPy's FFT is single-threaded)."""
    import multiprocessing as mp
    work = [(ny, nx, sample_slices, 100 + i) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        times = [_cpu_sample(work[0])]
    else:
        with mp.get_context('fork').Pool(procs) as pool:
            times = pool.map(_cpu_sample, work)
    wall = max(times) if procs > 1 else times[0]
    units = procs * ny * nx * sample_slices
    return units / wall / 1e9, wall, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    ny, nx, nz, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    sample_slices = 4 if ny * nx >= 2048 * 2048 else min(nz, max(4, (2048 * 2048 * 4) // (ny * nx)))
    vals = []
    for i in range(args.warmup + args.steps):
        v, wall, _ = cpu_baseline(ny, nx, procs, sample_slices)
        if i >= args.warmup:
            vals.append((v, wall))
    v = float(np.mean([a for a, _ in vals]))
    ms = float(np.mean([b for _, b in vals])) * 1e3
    sample = '%d independent processes x forward+adjoint of [1,%d,%d,%d] (NumPy complex128 oracle port of npfuncs.py:16-63 + hand adjoint)' % (procs, ny, nx, sample_slices)
    line = {
        'impl': 'reference', 'metric': 'multislice Gpixel*slice/s (forward + adjoint)', 'value': v, 'unit': 'Gpixel*slice/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c128', 'data': 'synthetic',
        'config': {'workload': desc, 'ny': ny, 'nx': nx, 'n_slice': nz, 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': 'Gpixel*slice/s', 'cores': procs, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'Gpixel*slice/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from beyond_dof_b200 import capi
    from beyond_dof_b200.models import FullfieldObjective

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    ny, nx, nz, desc = WORKLOADS[args.workload]
    B = 1
    if args.shape:
        B, ny, nx, nz = (int(v) for v in args.shape.split(','))
        desc = 'custom shape %s' % args.shape
    sm_reserve = 0
    if world > 1:
        import datetime
        # The sweep kernels are persistent (one CTA per SM, all of its shared memory and registers), so the NCCL kernels
        # that reduce the gradient buckets take SMs away from them while both run.  Measured on 2 B200 (2048^2x256, 8.6 GB
        # gradient, 14 ms all-reduce at NVLink line rate vs 13.7 ms of adjoint sweep): reserving SMs for NCCL and capping
        # its grid (NCCL_MAX_CTAS) lost more than it saved (tools/exp11.sh, exp12.sh); the default leaves both alone.
        sm_reserve = max(0, args.sm_reserve)
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))
    units_per_step = B * ny * nx * nz * world

    # synthetic inputs, created once on the device (value arm) / in pinned host memory (e2e arm)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    db = torch.empty((nz, B, ny, nx, 2), dtype=torch.float32, device=dev)
    for z0 in range(0, nz, 16):                       # bounded temporaries
        blk = db[z0:z0 + 16]
        blk.copy_(torch.rand(blk.shape, device=dev, generator=g))
        blk[..., 0] *= 1e-5
        blk[..., 1] *= 1e-6
    probe = torch.ones((ny, nx), dtype=torch.complex64, device=dev)
    obj = FullfieldObjective(db, probe, ENERGY_EV, PSIZE_CM, in_place=args.in_place)
    # target: measured magnitudes of a perturbed object (well inside (0, 1]); synthetic
    target_host = (0.9 + 0.1 * torch.rand((B, ny, nx), generator=torch.Generator().manual_seed(4321 + rank))).pin_memory()
    target_dev = target_host.to(dev)

    if world > 1:
        # data-parallel exchange (Horovod allreduce in fullfield.py:412; comm.Allreduce in cnn fullfield.py:350):
        # mean of the object gradient over ranks, reduced in z-buckets on a communication stream while the
        # adjoint sweep is still producing the remaining slices
        # default: copy engines over NVLink peer memory (no communication kernels on the SMs the sweep kernels own);
        # --exchange nccl = bucketed NCCL all-reduce.  If the peer-memory set-up fails on any rank, all fall back to NCCL.
        from beyond_dof_b200.dist import pick_exchange
        exchange_used = pick_exchange() if args.exchange == 'auto' else args.exchange
        if args.buckets <= 0:
            args.buckets = 8 if exchange_used == 'nccl' else 16      # measured optima (2 x B200)
        if exchange_used in ('ce', 'hybrid'):
            ok, why = 1, ''
            try:
                obj.enable_data_parallel(n_buckets=args.buckets, exchange=exchange_used)
            except Exception as ex:               # noqa: BLE001
                ok, why = 0, '%s: %s' % (type(ex).__name__, ex)
            t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
            if int(t_ok.item()) == 0:
                print('[rank %d] copy-engine exchange unavailable (%s): falling back to NCCL' % (rank, why or 'another rank failed'), file=sys.stderr, flush=True)
                exchange_used = 'nccl (copy-engine set-up failed)'
                obj.enable_data_parallel(n_buckets=args.buckets, exchange='nccl')
        else:
            obj.enable_data_parallel(n_buckets=args.buckets, exchange='nccl')
        if sm_reserve:
            capi.check(capi.lib.bdof_set_sm_reserve(sm_reserve))

    if world > 1 and args.diag:
        # diagnostics: the plain all-reduce of the whole gradient, and the step without any exchange
        for _ in range(2):
            dist.all_reduce(obj.grad, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dist.all_reduce(obj.grad, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        t_ar = (time.perf_counter() - t0) / 3
        dp_state = obj._dp
        obj._dp = None
        for _ in range(2):
            obj.step_device(target_dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            obj.step_device(target_dev)
        torch.cuda.synchronize()
        t_step = (time.perf_counter() - t0) / 3
        obj._dp = dp_state
        print('[diag rank %d] all-reduce of %.2f GB alone: %.2f ms; step without exchange: %.2f ms' %
              (rank, obj.grad.numel() * 4 / 1e9, t_ar * 1e3, t_step * 1e3), file=sys.stderr, flush=True)

    def step_device():
        return obj.step_device(target_dev)

    def step_e2e():
        return obj.step(target_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss = step_device()
    ev1.record()
    barrier()
    launches = capi.launch_count() - l0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = units_per_step / (ms_step * 1e-3) / 1e9

    # ---- e2e: this step's projections from pinned host memory, loss read back, through the public API
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_e2e()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step / (t.item() / args.steps * 1e-3) / 1e9
    h2d = target_host.numel() * target_host.element_size()
    d2h = 8

    # ---- e2e with the OBJECT on the host too: the literal drop-in of the cnn_propagator driver's call
    #      loss_grad(obj_delta, obj_beta, this_ind_batch, this_prj_batch) (cnn_propagator/fullfield.py:329,346) with NumPy arrays in
    #      and NumPy gradients out -- H2D of delta/beta [Y,X,Z], layout conversion, rotation (theta = 0), forward, loss, adjoint,
    #      back-rotation, layout conversion, D2H of both gradients.  PCIe-bound by construction; reported beside `e2e`, whose
    #      object lives on the GPU as the tf.Variables of the TF driver do (tensorflow_recon/fullfield.py:243-303).
    e2e_host = None
    if world == 1 and B == 1 and not args.no_host_object and not args.in_place and ny * nx * nz * 8 <= 12e9:
        try:
            from beyond_dof_b200 import fullfield_loss_and_grad
            from beyond_dof_b200.propagation import clear_plan_cache
            rng = np.random.default_rng(77)
            od = rng.random((ny, nx, nz), dtype=np.float32); od *= np.float32(1e-5)
            ob = rng.random((ny, nx, nz), dtype=np.float32); ob *= np.float32(1e-6)
            prj_np = target_host.numpy()
            one, zero = np.ones((ny, nx), np.float32), np.zeros((ny, nx), np.float32)

            def host_call():
                loss_h, (g_d, g_b), _ = fullfield_loss_and_grad(od, ob, np.zeros(1), prj_np, one, zero, ENERGY_EV, PSIZE_CM,
                                                                 propagate_last=False)
                return float(loss_h), g_d.cpu().numpy(), g_b.cpu().numpy()
            host_call()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            host_call()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e_host = {'value': ny * nx * nz / dt / 1e9, 'unit': 'Gpixel*slice/s', 'ms_per_step': dt * 1e3,
                        'h2d_bytes_per_step': int(od.nbytes + ob.nbytes + prj_np.nbytes), 'd2h_bytes_per_step': int(od.nbytes + ob.nbytes + 8),
                        'api': 'beyond_dof_b200.fullfield_loss_and_grad(NumPy delta, beta [Y,X,Z], theta, NumPy projections) -> loss, NumPy gradients '
                               '(pageable host memory, wall clock)'}
            del od, ob
            clear_plan_cache()
            torch.cuda.empty_cache()
        except Exception as ex:            # noqa: BLE001  (an optional extra must not take the bench line down)
            e2e_host = {'error': '%s: %s' % (type(ex).__name__, ex)}

    # ---- per-kernel in-situ timing (one extra, untimed step with CUDA events around every pass)
    kern = {}
    roofline = None
    cpu = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        dp_state = getattr(obj, '_dp', None)
        obj._dp = None                       # rank-0-only step: no collective in here
        obj.plan.profile_begin()
        obj.step_device(target_dev)
        prof = obj.plan.profile_end()
        obj._dp = dp_state
        px = B * ny * nx
        tot = sum(ms for _, ms in prof.values())
        for name, (cnt, ms) in prof.items():
            avg_ms = ms / cnt
            bpp = PASS_BYTES.get(name)
            kern[name] = {'launches': cnt, 'avg_ms': avg_ms, 'share_of_kernel_time': ms / tot,
                          'alg_bytes_per_px': bpp, 'achieved_gbs': (px * bpp / (avg_ms * 1e-3) / 1e9) if bpp else None}
        # Dominant kernel = the sweep kernel (csrc/sweepfft.cuh): ONE template whose forward and adjoint instantiations alternate
        # launch by launch and share the step ~50/50.  Its roofline entry is launch-weighted over both directions: algorithmic bytes
        # of all its launches / their summed in-situ durations; the per-direction figures stay in "kernels".
        fam = [n for n in prof if n.startswith('sweep_')] or [max(prof.items(), key=lambda kv: kv[1][1])[0]]
        fam_bytes = sum(px * PASS_BYTES[n] * prof[n][0] for n in fam if PASS_BYTES.get(n))
        fam_ms = sum(prof[n][1] for n in fam)
        fam_launches = sum(prof[n][0] for n in fam)
        a = fam_bytes / (fam_ms * 1e-3) / 1e9 if fam_ms > 0 else None
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                tj = json.load(f).get(args.workload, {})
            if all(n in tj for n in fam):
                traffic = sum(tj[n] * prof[n][0] for n in fam) / fam_launches      # ncu DRAM bytes per launch, launch-weighted
        except Exception:
            pass
        roofline = {'bound': 'hbm', 'kernel': 'sweep_kernel (%s; %d launches, %.1f %% of the kernel time of a step)' % (' + '.join(fam), fam_launches, 100 * fam_ms / tot),
                    'achieved': a, 'peak': peak, 'unit': 'GB/s', 'frac': a / peak if a else None,
                    'traffic': traffic, 'alg_bytes_per_launch': fam_bytes / fam_launches, 'avg_launch_ms': fam_ms / fam_launches,
                    'peak_source': peak_src,
                    'whole_step': {'alg_bytes_per_px_slice': STEP_BYTES, 'achieved': value * STEP_BYTES, 'frac': value * STEP_BYTES / peak}}
        # ---- CPU baseline: oracle port, 1 process (scalar port), bounded sample
        if not args.no_cpu:
            ss = 4 if ny * nx >= 2048 * 2048 else min(nz, max(4, (2048 * 2048 * 4) // (ny * nx)))
            v, wall, _ = cpu_baseline(ny, nx, 1, ss)
            cpu = {'value': v, 'unit': 'Gpixel*slice/s', 'cores': 1, 'kind': 'port',
                   'sample': 'forward+adjoint of [1,%d,%d,%d], NumPy complex128 oracle (%.1f s)' % (ny, nx, ss, wall)}
        line = {
            'metric': 'multislice Gpixel*slice/s (forward + adjoint)', 'value': value, 'unit': 'Gpixel*slice/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': desc, 'ny': ny, 'nx': nx, 'n_slice': nz, 'batch_per_gpu': B, 'semantics': 'numpy (last slice modulates only)',
                       'l2': 'inputs larger than L2 (%.1f GB of delta/beta + %.1f GB slice store per GPU streamed every step)' % (db.numel() * 4 / 1e9, db.numel() * 4 / 1e9),
                       'parallelism': ('dp%d: one field per GPU; mean of the object gradient (%.1f GB) over ranks every step in %d z-buckets overlapped with the adjoint sweep; exchange: %s'
                                       % (world, db.numel() * 4 / 1e9, args.buckets,
                                          'copy engines over NVLink peer memory (push partial shards, owner sums, gather)' if exchange_used == 'ce' else 'NCCL reduce-scatter (AVG) + copy-engine all-gather over NVLink peer memory' if exchange_used == 'hybrid' else 'NCCL all-reduce (AVG) on a communication stream' if exchange_used == 'nccl' else exchange_used)) if world > 1 else 'single GPU'},
            'e2e': {'value': e2e_value, 'unit': 'Gpixel*slice/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'api': 'beyond_dof_b200.models.FullfieldObjective.step(projection magnitudes in pinned host memory) -> loss'},
            'e2e_host_object': e2e_host,
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'kernels': kern, 'cpu_baseline': cpu,
            'loss': float(loss.item()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_config4(args):
    """BASELINE configs[3]: full-field tomographic reconstruction, 256^3 object, angles sharded over the ranks,
    minibatch of 10 angles per rank and update (reconstruct_fullfield.py:30,60), Adam, NCCL all-reduce of the object
    gradient.  One step = one optimiser update: rotate x10 -> multislice forward -> loss -> adjoint -> back-rotate x10
    -> all-reduce -> Adam, with this step's measured projections copied from pinned host memory."""
    import torch
    import torch.distributed as dist
    from beyond_dof_b200 import capi
    from beyond_dof_b200.models import TomographyObjective

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))
    n, mb, n_theta = 256, 10, 180
    g = torch.Generator(device=dev).manual_seed(1234)
    obj = torch.rand((n, n, n, 2), device=dev, generator=g) * torch.tensor([8.7e-7, 5.1e-8], device=dev)     # fullfield.py:273-274 scale
    probe = torch.ones((n, n), dtype=torch.complex64, device=dev)
    tomo = TomographyObjective(obj, probe, ENERGY_EV, PSIZE_CM, minibatch_size=mb, free_prop_cm=1e-4, propagate_last=True, step_size=1e-7)
    if world > 1:
        tomo.enable_data_parallel(exchange='nccl' if args.exchange in ('auto', 'hybrid') else args.exchange)
    thetas = np.linspace(0, np.pi, n_theta)
    mine = thetas[rank::world]
    tomo.prepare(mine)
    prj_host = (0.9 + 0.1 * torch.rand((mb, n, n), generator=torch.Generator().manual_seed(4321 + rank))).pin_memory()

    def batch(i):
        return np.take(mine, np.arange(i * mb, (i + 1) * mb), mode='wrap')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for i in range(args.warmup):
        tomo.step(batch(i), prj_host)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = tomo.step(batch(args.warmup + i), prj_host)
    ev1.record()
    barrier()
    launches = capi.launch_count() - l0
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    units = mb * n * n * n * world
    value = units / (ms_step * 1e-3) / 1e9
    if rank == 0:
        clocks = sampler.stop()
        line = {
            'metric': 'multislice Gpixel*slice/s (forward + adjoint)', 'value': value, 'unit': 'Gpixel*slice/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': 'full-field tomography 256^3, 180 angles sharded over ranks, minibatch 10 angles per rank and update '
                                   '(BASELINE configs[3]); rotate + multislice forward/adjoint + back-rotate + all-reduce + Adam',
                       'ny': n, 'nx': n, 'n_slice': n, 'batch_per_gpu': mb, 'semantics': 'tf (every slice propagates), free_prop_cm 1e-4',
                       'l2': 'every step streams the %.1f GB rotated minibatch object and its slice store' % (mb * n ** 3 * 8 / 1e9),
                       'parallelism': 'dp%d' % world},
            'recon_iter_per_s': 1e3 / ms_step, 'epoch_s': (n_theta / (mb * world)) * ms_step * 1e-3,
            'e2e': {'value': value, 'unit': 'Gpixel*slice/s', 'h2d_bytes_per_step': prj_host.numel() * 4, 'd2h_bytes_per_step': 8,
                    'api': 'beyond_dof_b200.models.TomographyObjective.step(theta_batch, projections in pinned host memory) -> loss '
                           '(the timed region IS the public call: H2D copy, update, loss read-back)'},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': None, 'cpu_baseline': None, 'loss': float(loss),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='config2', choices=sorted(WORKLOADS) + ['config4'])
    ap.add_argument('--impl', default='bdof', choices=['bdof', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    ap.add_argument('--no-host-object', action='store_true', help='skip the extra end-to-end leg with delta/beta and the gradients in host memory')
    ap.add_argument('--sm-reserve', type=int, default=0, help='SMs left free for NCCL while the sweep runs (N > 1)')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'ce', 'nccl', 'hybrid'], help='N > 1: gradient exchange (copy engines over peer memory, or NCCL all-reduce)')
    ap.add_argument('--buckets', type=int, default=0, help='z-buckets of the gradient all-reduce (N > 1)')
    ap.add_argument('--shape', default=None, help='experiment: B,NY,NX,NZ overrides the workload shape')
    ap.add_argument('--diag', action='store_true', help='N > 1: print the plain all-reduce time and the step time without exchange')
    ap.add_argument('--in-place', action='store_true', help='adjoint overwrites delta/beta with the gradient (needed for the 4096^2x512 size)')
    args = ap.parse_args()
    if args.workload == 'config4':
        if args.impl == 'reference':
            print(json.dumps({'impl': 'reference', 'unavailable': 'the reference arm is defined on the default workload (config2)'}))
            return 0
        return run_config4(args)
    if args.impl == 'reference':
        return run_reference(args)
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())

#!/usr/bin/env python
"""Benchmark of the multislice hot path (BASELINE.json metric: multislice Gpixel*slice/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|headline|config1] [--impl reference]

One "step" = one minibatch of the data-parallel reconstruction on every GPU: K = 10 fields (projection angles; minibatch_size of
reconstruct_fullfield.py:30,60), each forward multislice + |psi| loss + adjoint gradient, the K gradients summed (fused into the
adjoint kernels' stores), then -- with N > 1 ranks -- ONE exchange: the mean of the object gradient over ranks.
Default workload is BASELINE.json configs[1]: random delta/beta phantom 2048 x 2048 x 256, forward + adjoint.  1 unit = one
pixel advanced through one slice (forward + adjoint counts the slice once).  Weak scaling: every rank evaluates its own K
fields; the exchange (8.6 GB) is overlapped with the last adjoint sweep (--exchange auto: copy engines over NVLink peer
memory between two GPUs, NCCL all-reduce for three or more).  --fields-per-exchange 1 is the harshest ratio
(one field per exchange); the other BASELINE configs: --workload config3 | config4 | config5.

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
def _make_db(torch, dev, nz, B, ny, nx, seed):
    """synthetic (delta, beta) in the native slice-major layout, config-2 recipe (delta <= 1e-5, beta <= 1e-6)"""
    g = torch.Generator(device=dev).manual_seed(seed)
    db = torch.empty((nz, B, ny, nx, 2), dtype=torch.float32, device=dev)
    for z0 in range(0, nz, 16):                       # bounded temporaries
        blk = db[z0:z0 + 16]
        blk.copy_(torch.rand(blk.shape, device=dev, generator=g))
        blk[..., 0] *= 1e-5
        blk[..., 1] *= 1e-6
    return db


def _time_steps(torch, dist, world, dev, fn, steps):
    """barrier + synchronize, `steps` calls of fn timed by CUDA events on the current stream, max over ranks -> ms per step"""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    out = None
    for _ in range(steps):
        out = fn()
    ev1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / steps, out


def _traffic(workload_key):
    """ncu DRAM bytes per sweep launch (dram__bytes_read.sum + dram__bytes_write.sum, `ncu --set full`), transcribed from the
    committed summary named in the entry: NOT measured by this run."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f).get(workload_key)
    except Exception:
        return None


def _sweep_roofline(obj, px, nz, ms_step, value_per_gpu, peak, peak_src, traffic_key):
    """Roofline record of the dominant kernel family (sweep_kernel: one launch per slice and direction).  Durations come from
    ONE event pair around the whole forward launch sequence and one around the adjoint's (bdof_plan_last_times), inside a
    normal step: the launches overlap exactly as in production (programmatic dependent launch), the few small kernels of each
    call (probe broadcast, phase scale) are counted as sweep time, and forward + adjoint <= ms_per_step by construction."""
    lt = obj.plan.last_times()
    (f_ms, f_n), (a_ms, a_n) = lt['forward'], lt['adjoint']
    n_launch = 2 * nz
    avg_ms = (f_ms + a_ms) / n_launch
    alg = px * STEP_BYTES / 2.0                                   # (40 + 56) / 2 bytes per pixel per launch
    ach = alg / (avg_ms * 1e-3) / 1e9
    tr = _traffic(traffic_key)
    return {'bound': 'hbm',
            'kernel': 'sweep_kernel (forward + adjoint instantiations, %d launches per step)' % n_launch,
            'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
            'traffic': (tr or {}).get('bytes_per_launch'), 'traffic_source': (tr or {}).get('source'),
            'alg_bytes_per_launch': alg, 'avg_launch_ms': avg_ms,
            'forward_ms': f_ms, 'adjoint_ms': a_ms, 'kernel_ms_per_step': f_ms + a_ms, 'ms_per_step': ms_step,
            'consistent': bool(f_ms + a_ms <= ms_step * 1.02),
            'by_direction': {'forward': {'alg_bytes_per_px': 40, 'achieved_gbs': px * 40 * nz / (f_ms * 1e-3) / 1e9},
                             'adjoint': {'alg_bytes_per_px': 56, 'achieved_gbs': px * 56 * nz / (a_ms * 1e-3) / 1e9}},
            'peak_source': peak_src,
            'whole_step': {'alg_bytes_per_px_slice': STEP_BYTES, 'achieved_per_gpu': value_per_gpu * STEP_BYTES,
                           'frac': value_per_gpu * STEP_BYTES / peak},
            'note': 'the contract figure stays 96 B per pixel*slice; when K > 1 fields are summed per step the adjoint of every field but the '
                    'first also reads the running sum (+8 B), which is not counted'}


def cufft_comparison(torch, dev, ny, nx, n_sample, steps=3):
    """COMPARISON ONLY (north_star: "cuFFT timed only as a comparison"; never on the product path): the reference's unfused
    TF loop (tensorflow_recon/util.py:464-483: exp, multiply, fft2, fftshift, multiply by H, ifftshift, ifft2) restated op by
    op on torch (cuFFT + elementwise kernels), complex64, slice-major inputs (kinder than the reference's z-fastest layout),
    forward only and forward + autograd backward (what the TF driver's minimize() runs), on n_sample slices of this field."""
    import math
    k = 2 * math.pi * 1.0 / (1240. / ENERGY_EV)
    g = torch.Generator(device=dev).manual_seed(5)
    delta = torch.rand((n_sample, 1, ny, nx), device=dev, generator=g) * 1e-5
    beta = torch.rand((n_sample, 1, ny, nx), device=dev, generator=g) * 1e-6
    fy = torch.linspace(-0.5, 0.5, ny, device=dev, dtype=torch.float64)
    fx = torch.linspace(-0.5, 0.5, nx, device=dev, dtype=torch.float64)
    lam = 1240. / ENERGY_EV
    H = torch.exp(-1j * math.pi * lam * 1.0 * (fy[:, None] ** 2 + fx[None, :] ** 2)).to(torch.complex64)       # centred, as get_kernel
    target = torch.full((1, ny, nx), 0.95, device=dev)

    def forward(d, b):
        psi = torch.ones((1, ny, nx), dtype=torch.complex64, device=dev)
        for i in range(n_sample):
            c = torch.exp(torch.complex(-k * b[i], k * d[i]))
            psi = psi * c
            psi = torch.fft.ifft2(torch.fft.ifftshift(torch.fft.fftshift(torch.fft.fft2(psi), dim=(1, 2)) * H, dim=(1, 2)))
        return psi

    def fwd_only():
        with torch.no_grad():
            return forward(delta, beta)

    def fwd_bwd():
        d = delta.clone().requires_grad_(True)
        b = beta.clone().requires_grad_(True)
        loss = ((forward(d, b).abs() - target) ** 2).mean()
        loss.backward()
        return loss

    res = {}
    for name, fn in (('forward', fwd_only), ('forward_backward', fwd_bwd)):
        try:
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[name] = {'value': ny * nx * n_sample / (ms * 1e-3) / 1e9, 'unit': 'Gpixel*slice/s', 'ms': ms}
        except Exception as ex:            # noqa: BLE001
            res[name] = {'error': '%s: %s' % (type(ex).__name__, str(ex)[:200])}
    res['what'] = ('reference TF loop (util.py:464-483) restated on torch.fft/cuFFT, unfused, complex64, [1,%d,%d] x %d slices; '
                   'comparison only, not on the product path' % (ny, nx, n_sample))
    del delta, beta
    torch.cuda.empty_cache()
    return res


def host_object_leg(torch, ny, nx, nz, target_host):
    """e2e with the OBJECT on the host too: the literal drop-in of the cnn_propagator driver's call
    loss_grad(obj_delta, obj_beta, this_ind_batch, this_prj_batch) (cnn_propagator/fullfield.py:329,346) with NumPy arrays in
    and NumPy gradients out.  PCIe-bound by construction; reported beside `e2e`, whose object lives on the GPU as the
    tf.Variables of the TF driver do (tensorflow_recon/fullfield.py:243-303)."""
    try:
        from beyond_dof_b200.models import fullfield_loss_and_grad_host
        rng = np.random.default_rng(77)
        od = rng.random((ny, nx, nz), dtype=np.float32); od *= np.float32(1e-5)
        ob = rng.random((ny, nx, nz), dtype=np.float32); ob *= np.float32(1e-6)
        prj_np = (target_host[0] if target_host.dim() == 4 else target_host).numpy()       # one field's projection
        one, zero = np.ones((ny, nx), np.float32), np.zeros((ny, nx), np.float32)
        g_d, g_b = np.empty_like(od), np.empty_like(ob)

        def host_call():
            return fullfield_loss_and_grad_host(od, ob, prj_np, one, zero, ENERGY_EV, PSIZE_CM, propagate_last=False,
                                                out=(g_d, g_b))
        host_call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {'value': ny * nx * nz / dt / 1e9, 'unit': 'Gpixel*slice/s', 'ms_per_step': dt * 1e3,
                'h2d_bytes_per_step': int(od.nbytes + ob.nbytes + prj_np.nbytes), 'd2h_bytes_per_step': int(od.nbytes + ob.nbytes + 8),
                'api': 'beyond_dof_b200.models.fullfield_loss_and_grad_host(NumPy delta, beta [Y,X,Z], NumPy projections) -> loss, NumPy '
                       'gradients (pageable host arrays in and out, chunked through pinned staging; wall clock)'}
    except Exception as ex:            # noqa: BLE001  (an optional extra must not take the bench line down)
        return {'error': '%s: %s' % (type(ex).__name__, str(ex)[:300])}


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
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))
    peak, peak_src = measured_peaks()
    K = max(1, args.fields_per_exchange)

    def build(ny, nx, nz, B, in_place):
        db = _make_db(torch, dev, nz, B, ny, nx, 1234 + rank)
        probe = torch.ones((ny, nx), dtype=torch.complex64, device=dev)
        obj = FullfieldObjective(db, probe, ENERGY_EV, PSIZE_CM, in_place=in_place)
        shape = (B, ny, nx) if (K == 1 or in_place) else (K, B, ny, nx)             # one measured projection per field
        th = (0.9 + 0.1 * torch.rand(shape, generator=torch.Generator().manual_seed(4321 + rank))).pin_memory()
        return obj, th, th.to(dev)

    def measure(obj, target_host, target_dev, ny, nx, nz, B, steps, warmup, traffic_key, K=K):
        """value (device-resident inputs), e2e (public API, pinned host projections in, loss out), roofline"""
        units = B * ny * nx * nz * world * K

        def step_device():
            return obj.step_device(target_dev, accumulate=K)

        def step_e2e():
            return obj.step(target_host, accumulate=K)
        for _ in range(warmup):
            step_device()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        l0 = capi.launch_count()
        ms_step, loss = _time_steps(torch, dist, world, dev, step_device, steps)
        launches = capi.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        value = units / (ms_step * 1e-3) / 1e9
        roof = _sweep_roofline(obj, B * ny * nx, nz, ms_step / K, value / world, peak, peak_src, traffic_key) if rank == 0 else None
        for _ in range(max(1, warmup // 2)):
            step_e2e()
        ms_e2e, _ = _time_steps(torch, dist, world, dev, step_e2e, steps)
        e2e = {'value': units / (ms_e2e * 1e-3) / 1e9, 'unit': 'Gpixel*slice/s',
               'h2d_bytes_per_step': target_host.numel() * target_host.element_size(), 'd2h_bytes_per_step': 8,
               'api': 'beyond_dof_b200.models.FullfieldObjective.step(projection magnitudes in pinned host memory) -> loss'}
        return {'value': value, 'ms_per_step': ms_step, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
                'roofline': roof, 'loss': float(loss.item())}

    obj, target_host, target_dev = build(ny, nx, nz, B, args.in_place)
    exchange_used = None
    if world > 1:
        # data-parallel exchange (Horovod allreduce in fullfield.py:412; comm.Allreduce in cnn fullfield.py:350): mean of the
        # object gradient over ranks, z-bucket by z-bucket while the adjoint sweep is still producing the remaining slices
        from beyond_dof_b200.dist import pick_exchange
        exchange_used = pick_exchange() if args.exchange == 'auto' else args.exchange
        if args.buckets <= 0:
            args.buckets = 8 if exchange_used == 'nccl' else 16      # measured optima (2 x B200)
        ok, why = 1, ''
        try:
            obj.enable_data_parallel(n_buckets=args.buckets, exchange=exchange_used, sm_reserve=args.sm_reserve)
        except Exception as ex:               # noqa: BLE001
            ok, why = 0, '%s: %s' % (type(ex).__name__, ex)
        t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            print('[rank %d] %s exchange unavailable (%s): falling back to NCCL' % (rank, exchange_used, why or 'another rank failed'), file=sys.stderr, flush=True)
            exchange_used = 'nccl (set-up of the requested exchange failed)'
            obj.enable_data_parallel(n_buckets=args.buckets, exchange='nccl', sm_reserve=args.sm_reserve)
        exchange_used = getattr(obj, 'exchange_name', exchange_used)

    main = measure(obj, target_host, target_dev, ny, nx, nz, B, args.steps, args.warmup, '%dx%d' % (ny, nx))

    extras = {}
    if rank == 0 and world == 1:
        if not args.no_host_object and B == 1 and not args.in_place and ny * nx * nz * 8 <= 12e9:
            extras['e2e_host_object'] = host_object_leg(torch, ny, nx, nz, target_host)
        if not args.no_cufft:
            extras['cufft_comparison'] = cufft_comparison(torch, dev, ny, nx, 32 if ny * nx <= 2048 * 2048 else 8)
        if not args.no_cpu:
            ss = 4 if ny * nx >= 2048 * 2048 else min(nz, max(4, (2048 * 2048 * 4) // (ny * nx)))
            v, wall, _ = cpu_baseline(ny, nx, 1, ss)
            extras['cpu_baseline'] = {'value': v, 'unit': 'Gpixel*slice/s', 'cores': 1, 'kind': 'port',
                                      'sample': 'forward+adjoint of [1,%d,%d,%d], NumPy complex128 oracle (%.1f s)' % (ny, nx, ss, wall)}
    # ---- the north_star target size on the same box, in the same run: 4096^2 x 512, gradient in place (137 GB)
    headline = None
    if world == 1 and args.workload == 'config2' and not args.shape and not args.no_headline:
        hy, hx, hz, hdesc = WORKLOADS['headline']
        try:
            free_b, total_b = torch.cuda.mem_get_info()
            del obj, target_dev
            torch.cuda.empty_cache()
            free_b, total_b = torch.cuda.mem_get_info()
            need = hy * hx * hz * 16 + 3 * hy * hx * 8 + (2 << 30)
            if free_b < need:
                headline = {'skipped': 'needs %.0f GB of device memory, %.0f GB free' % (need / 1e9, free_b / 1e9)}
            else:
                hobj, hth, htd = build(hy, hx, hz, 1, True)
                h = measure(hobj, hth, htd, hy, hx, hz, 1, max(2, min(args.steps, 5)), 3, '%dx%d' % (hy, hx), K=1)
                headline = {'workload': hdesc + ', gradient written in place over delta/beta', 'ny': hy, 'nx': hx, 'n_slice': hz,
                            'value': h['value'], 'unit': 'Gpixel*slice/s', 'ms_per_step': h['ms_per_step'], 'e2e': h['e2e'],
                            'roofline': h['roofline'], 'gpu_launches': h['gpu_launches'], 'clocks': h['clocks'], 'loss': h['loss'],
                            'steps': max(2, min(args.steps, 5)), 'warmup': 3}
                del hobj, htd
                torch.cuda.empty_cache()
        except Exception as ex:            # noqa: BLE001
            headline = {'error': '%s: %s' % (type(ex).__name__, str(ex)[:300])}

    if rank == 0:
        gb = nz * B * ny * nx * 8 / 1e9
        if world > 1:
            par = ('dp%d: %d field(s) per GPU and exchange; mean of the object gradient (%.1f GB) over ranks every step in %d z-buckets '
                   'overlapped with the adjoint sweep; exchange: %s' % (world, K, gb, args.buckets, exchange_used))
        else:
            par = 'single GPU' + ('' if K == 1 else ', %d fields accumulated per step' % K)
        line = {
            'metric': 'multislice Gpixel*slice/s (forward + adjoint)', 'value': main['value'], 'unit': 'Gpixel*slice/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': desc, 'ny': ny, 'nx': nx, 'n_slice': nz, 'batch_per_gpu': B, 'fields_per_exchange': K,
                       'semantics': 'numpy (last slice modulates only)',
                       'l2': 'inputs larger than L2 (%.1f GB of delta/beta + %.1f GB slice store per GPU streamed every step)' % (gb, gb),
                       'parallelism': par},
            'e2e': main['e2e'], 'e2e_host_object': extras.get('e2e_host_object'),
            'gpu_launches': main['gpu_launches'], 'clocks': main['clocks'], 'roofline': main['roofline'],
            'headline': headline, 'cufft_comparison': extras.get('cufft_comparison'), 'cpu_baseline': extras.get('cpu_baseline'),
            'loss': main['loss'],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_config3(args):
    """BASELINE configs[2]: ptychography forward model + loss gradient, 256 x 256 object x 128 slices, 1024 probe positions
    (32 x 32 raster, pos = 32 + 6 j) batch-sharded over the ranks (contiguous blocks, cnn_propagator/ptychography.py:
    292-297), 64 x 64 Gaussian probe (reconstruct_ptycho.py:92-94), far-field detector, TF semantics.  One step = one update on
    ALL 1024 positions: window cut -> multislice forward (resident kernels) -> far field -> loss -> adjoint -> deterministic window
    accumulation -> NCCL all-reduce of the 67 MB object gradient -> Adam -> clip, with this rank's measured diffraction
    magnitudes copied from pinned host memory every step and the loss read back.  Total work is fixed: strong scaling."""
    import torch
    import torch.distributed as dist
    from beyond_dof_b200 import capi
    from beyond_dof_b200.models import PtychographyObjective
    from beyond_dof_b200.dist import shard_contiguous

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))
    n, nz, ps, n_pos = 256, 128, 64, 1024
    g = torch.Generator(device=dev).manual_seed(1234)
    obj = torch.rand((nz, n, n, 2), device=dev, generator=g) * torch.tensor([1e-5, 1e-6], device=dev)
    yy = torch.arange(ps, dtype=torch.float64) - (ps - 1) / 2
    r2 = yy[:, None] ** 2 + yy[None, :] ** 2
    mag, phase = torch.exp(-r2 / (2 * 6.0 ** 2)), 0.5 * torch.exp(-r2 / (2 * 6.0 ** 2))
    probe = torch.complex(mag * torch.cos(phase), mag * torch.sin(phase)).to(torch.complex64).to(dev)
    jj, ii = np.meshgrid(np.arange(32), np.arange(32))
    pos_all = np.stack([32 + 6 * ii.ravel(), 32 + 6 * jj.ravel()], 1)
    mine = pos_all[shard_contiguous(n_pos, rank, world)]
    pty = PtychographyObjective(obj, probe, (ps, ps), ENERGY_EV, PSIZE_CM, n_pos_per_step=len(mine), n_pos_total=n_pos, step_size=1e-7)
    if world > 1:
        pty.enable_data_parallel()
    prj_host = (0.05 + torch.rand((len(mine), ps, ps), generator=torch.Generator().manual_seed(4321 + rank))).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        pty.step(mine, prj_host)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss = pty.step(mine, prj_host)
    ev1.record()
    barrier()
    launches = capi.launch_count() - l0
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    units = n_pos * ps * ps * nz
    value = units / (ms_step * 1e-3) / 1e9
    if rank == 0:
        clocks = sampler.stop()
        lt = pty.plan.last_times()
        peak, peak_src = measured_peaks()
        ms_k = lt['forward'][0] + lt['adjoint'][0]
        px = len(mine) * ps * ps
        # resident kernels: psi never leaves the SM; per pixel*slice the forward reads (delta, beta) 8 B and writes the stored psi 8 B
        # and the stash 8 B, the adjoint reads stash 8 B + psi 8 B and writes the gradient 8 B -> 48 B
        alg = px * nz * 48.0
        line = {
            'metric': 'multislice Gpixel*slice/s (forward + adjoint)', 'value': value, 'unit': 'Gpixel*slice/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': 'ptychography 256x256x128 object, 1024 probe positions (64x64 Gaussian probe, far field) sharded over the ranks '
                                   '(BASELINE configs[2]); window cut + multislice forward/adjoint + window accumulation + all-reduce + Adam',
                       'ny': ps, 'nx': ps, 'n_slice': nz, 'batch_per_gpu': len(mine), 'semantics': 'tf (every slice propagates), far field',
                       'l2': 'every step streams %.2f GB of windows, stored slices and gradients per GPU' % (px * nz * 24 / 1e9),
                       'parallelism': 'dp%d over scan positions, NCCL all-reduce of the %.0f MB object gradient per step' % (world, obj.numel() * 4 / 1e6)},
            'recon_iter_per_s': 1e3 / ms_step,
            'e2e': {'value': value, 'unit': 'Gpixel*slice/s', 'h2d_bytes_per_step': prj_host.numel() * 4 + len(mine) * 8, 'd2h_bytes_per_step': 8,
                    'api': 'beyond_dof_b200.models.PtychographyObjective.step(positions, diffraction magnitudes in pinned host memory) -> loss '
                           '(the timed region IS the public call: H2D copies, update, loss read-back)'},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'resident_forward_kernel + resident_adjoint_kernel (one launch each per step)',
                         'achieved': alg / (ms_k * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s', 'frac': alg / (ms_k * 1e-3) / 1e9 / peak,
                         'traffic': None, 'alg_bytes_per_launch': alg / 2, 'avg_launch_ms': ms_k / 2, 'forward_ms': lt['forward'][0],
                         'adjoint_ms': lt['adjoint'][0], 'peak_source': peak_src,
                         'note': 'forward_ms / adjoint_ms are the whole bdof_forward / bdof_adjoint calls (resident kernel + far-field passes); '
                                 'the field is on chip, so this kernel is bound by fp32 issue and shared memory, not by HBM (SURVEY 8d)'},
            'cpu_baseline': None, 'loss': float(loss),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_config5(args):
    """BASELINE configs[4]: tiling-based forward multislice of ONE 16384 x 16384 field through 1000 slices of an axially repeating
    zone plate, split over the ranks in a 1x1 / 1x2 / 2x2 / 2x4 block grid; per slice every rank steps its
    local FFT windows and stores its border strips into the neighbours' aprons over NVLink peer memory (tiling.TiledMultislice,
    csrc/tilehalo.cu).  One step = the whole propagation; total work is fixed: strong scaling.  e2e = the same call plus the
    read-back of the exit intensity's checksum (the field itself stays distributed on the GPUs, as a simulation would keep it)."""
    import torch
    import torch.distributed as dist
    from beyond_dof_b200 import capi
    from beyond_dof_b200.tiling import TiledMultislice

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    grid = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]
    n = args.tile_field
    nz = args.tile_slices
    halo = args.halo
    lmbda = 1240. / ENERGY_EV
    f_nm = (n / 2 - 192) * 4.0 / lmbda * 1.0            # outer zone width 4 px at r_N = n/2 - 192 px (SURVEY 8d recipe, 1 nm voxels)
    n_zones = int(((n / 2 - 192) ** 2) / (lmbda * f_nm))

    def db_block(y0, x0, h, w):
        ys = (torch.arange(y0, y0 + h, device=dev) % n).double() - (n - 1) / 2
        xs = (torch.arange(x0, x0 + w, device=dev) % n).double() - (n - 1) / 2
        out = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
        for r0 in range(0, h, 1024):                  # bounded temporaries
            r2 = ys[r0:r0 + 1024, None] ** 2 + xs[None, :] ** 2
            zone = torch.floor(r2 / (lmbda * f_nm)).long()
            mask = ((zone % 2) == 1) & (zone <= n_zones)
            out[r0:r0 + 1024, :, 0] = mask * 1.0e-4
            out[r0:r0 + 1024, :, 1] = mask * 1.0e-5
        return out
    lengths = tuple(int(v) for v in args.tile_lengths.split(',')) if args.tile_lengths else None
    tm = TiledMultislice(n, n, grid, halo, ENERGY_EV, PSIZE_CM, nz, db_block, **({'lengths': lengths} if lengths else {}))

    def step():
        out = tm.run()
        return (out.abs() ** 2).sum(dtype=torch.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        chk = step()
    ev1.record()
    barrier()
    launches = capi.launch_count() - l0
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    ms_step = t.item() / args.steps
    # e2e: the public call + the scalar read-back
    host = torch.empty((), dtype=torch.float64).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host.copy_(step(), non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    units = float(n) * n * nz
    value = units / (ms_step * 1e-3) / 1e9
    if rank == 0:
        clocks = sampler.stop()
        peak, peak_src = measured_peaks()
        line = {
            'metric': 'multislice Gpixel*slice/s (forward, tiled)', 'value': value, 'unit': 'Gpixel*slice/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': 'tiling-based forward multislice, %dx%d field x %d slices of an axially repeating zone plate, %dx%d block grid '
                                   'with per-slice NVLink halo exchange (BASELINE configs[4])' % (n, n, nz, grid[0], grid[1]),
                       'ny': n, 'nx': n, 'n_slice': nz, 'halo': halo, 'window': [tm.ly, tm.lx], 'windows_per_gpu': tm.n_tiles, 'apron': tm.apron,
                       'redundant_compute': tm.redundancy, 'semantics': 'numpy (last slice modulates only)',
                       'l2': 'every slice streams the %.1f GB of this GPU\'s windows' % (tm.n_tiles * tm.ly * tm.lx * 8 / 1e9),
                       'parallelism': 'one field over %d GPU(s): %dx%d blocks, border strips stored into the neighbours\' aprons over NVLink peer '
                                      'memory by one kernel per slice, stream-memop flags' % (world, grid[0], grid[1])},
            'e2e': {'value': units / (ms_e2e * 1e-3) / 1e9, 'unit': 'Gpixel*slice/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 8,
                    'api': 'beyond_dof_b200.tiling.TiledMultislice.run() + read-back of the exit intensity sum (wall clock); the object is analytic '
                           'and lives on the GPUs, nothing is uploaded per step'},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'line_kernel row pass with transmission + column pass (per window batch)',
                         'achieved': value / world * tm.redundancy * 40.0, 'peak': peak, 'unit': 'GB/s',
                         'frac': value / world * tm.redundancy * 40.0 / peak, 'traffic': None, 'peak_source': peak_src,
                         'note': 'whole-step figure per GPU: forward contract 40 B per TRANSFORMED pixel*slice (windows incl. halo), cut/paste '
                                 'copies and the exchange counted as overhead'},
            'cpu_baseline': None, 'exit_intensity_sum': float(chk.item()),
        }
        print(json.dumps(line), flush=True)
    tm.close()
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
        # the multislice kernels of the last update (one event pair around each direction).  The cluster-resident kernels move
        # 24 + 24 B per pixel*slice, the sweep kernels 96: the fraction says how far from an HBM bound they are
        lt = tomo.plan.last_times()
        peak, peak_src = measured_peaks()
        cluster = tomo.plan.is_cluster_resident()
        kernel_ms = lt['forward'][0] + lt['adjoint'][0]
        alg_bytes = (48.0 if cluster else 96.0) * mb * n * n * n
        roofline = {'bound': 'hbm',
                    'kernel': 'cluster_forward_kernel + cluster_adjoint_kernel (one launch each per update)' if cluster else 'sweep kernels',
                    'achieved': alg_bytes / (kernel_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                    'frac': alg_bytes / (kernel_ms * 1e-3) / 1e9 / peak, 'traffic': None, 'alg_bytes_per_px_slice': 48 if cluster else 96,
                    'forward_ms': lt['forward'][0], 'adjoint_ms': lt['adjoint'][0], 'kernel_ms_per_step': kernel_ms, 'ms_per_step': ms_step,
                    'peak_source': peak_src,
                    'note': 'not HBM-bound: ten 8-CTA clusters occupy 80 of 148 SMs and are bound by fp32 issue there '
                            '(profiles/r02_cluster_kernels_256.txt); the rest of the step is the rotation stage, Adam and the exchange'}
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
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': None, 'loss': float(loss),
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
    ap.add_argument('--workload', default='config2', choices=sorted(WORKLOADS) + ['config3', 'config4', 'config5'])
    ap.add_argument('--impl', default='bdof', choices=['bdof', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    ap.add_argument('--no-host-object', action='store_true', help='skip the extra end-to-end leg with delta/beta and the gradients in host memory')
    ap.add_argument('--halo', type=int, default=8, help='config5: halo of the local FFT windows (the reference kernel_size 17 = 8 px)')
    ap.add_argument('--tile-field', type=int, default=16384, help='config5: side of the global field')
    ap.add_argument('--tile-lengths', default=None, help='config5: comma-separated FFT window lengths to choose from (default: all powers of two 256..8192)')
    ap.add_argument('--tile-slices', type=int, default=1000, help='config5: slices')
    ap.add_argument('--no-cufft', action='store_true', help='skip the cuFFT comparison leg (unfused reference loop on torch.fft)')
    ap.add_argument('--no-headline', action='store_true', help='skip the 4096^2 x 512 headline record of the default run')
    ap.add_argument('--fields-per-exchange', type=int, default=10, help='fields (projection angles) each rank evaluates and sums per step / gradient '
                    'exchange: the minibatch_size = 10 of the reference drivers (reconstruct_fullfield.py:30,60); the same at every N')
    ap.add_argument('--sm-reserve', type=int, default=0, help='SMs left free for NCCL while the sweep runs (N > 1); -1 = as many as cost no extra round of tiles')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'ce', 'nccl', 'hybrid'], help='N > 1: gradient exchange (copy engines over peer memory, or NCCL all-reduce)')
    ap.add_argument('--buckets', type=int, default=0, help='z-buckets of the gradient all-reduce (N > 1)')
    ap.add_argument('--shape', default=None, help='experiment: B,NY,NX,NZ overrides the workload shape')
    ap.add_argument('--diag', action='store_true', help='N > 1: print the plain all-reduce time and the step time without exchange')
    ap.add_argument('--in-place', action='store_true', help='adjoint overwrites delta/beta with the gradient (needed for the 4096^2x512 size)')
    args = ap.parse_args()
    if args.workload in ('config3', 'config4', 'config5'):
        if args.impl == 'reference':
            print(json.dumps({'impl': 'reference', 'unavailable': 'the reference arm is defined on the default workload (config2)'}))
            return 0
        return {'config3': run_config3, 'config4': run_config4, 'config5': run_config5}[args.workload](args)
    if args.impl == 'reference':
        return run_reference(args)
    return run_gpu(args)


if __name__ == '__main__':
    sys.exit(main())

#!/usr/bin/env python
"""bench.py — quantised-training throughput of the DFXP hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one full training step (forward, loss, backward with quantised gradients, momentum SGD,
range controller) of the workload below on synthetic data.  N > 1 is launched by torchrun, one rank
per GPU, weak scaling (fixed per-GPU batch), gradient + overflow-counter all-reduce over NCCL.

Workload at N=1 (BASELINE.json configs[1]): CIFAR10_Resnet20 (models.py:453), batch 256 per GPU,
8-bit dynamic fixed point for W/A/G, stochastic rounding everywhere (the reference's behaviour).
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class, from the committed ncu captures
# (profiles/): (workload, C-ABI entry) -> bytes per launch (average over the class), or absent = null
TRAFFIC = {
    # profiles/r1_conv_ldg_halo_step_ncu_full_summary.csv (ncu --set full inside bench.py, halo-patch loader): the stage-1 launches
    # (conv_ldg_kernel<16, 1, 1, 0>) read 4.25-4.32 MB from DRAM and write ~0 (their 4-17 MB outputs stay in the 126 MB L2)
    # vs 8.4-21 MB algorithmic
    ('resnet20', 'lbt_conv_i8_fprop'): 4.3e6,
}

WORKLOADS = {
    # name: (model ctor name, image, classes, default batch per GPU, bits, grad_bits)
    'resnet20': ('CIFAR10_Resnet20', 32, 10, 256, 8, None),
    'cifar10': ('CIFAR10_Model', 32, 10, 128, 8, None),
    'resnet18': ('Resnet18', 224, 1000, 256, 8, None),
    'resnet50': ('Resnet50', 224, 1000, 128, 8, 16),          # BASELINE config 5: 8-bit W/A + 16-bit G
    'resnet50_g8': ('Resnet50', 224, 1000, 128, 8, None),
}


_REAL_STDOUT = None


def guard_stdout():
    """Everything libraries print to fd 1 during the run (NCCL's version banner, for one) goes to stderr; the ONE JSON line is
    written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='resnet20', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (0 = the workload default)')
    ap.add_argument('--no-graph', action='store_true', help='do not capture the step in a CUDA graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-batch', type=int, default=0, help='batch of the CPU baseline sample (0 = same as GPU)')
    ap.add_argument('--breakdown', action='store_true', help='also print the per-kernel time table to stderr')
    ap.add_argument('--dump-launches', default='', help='write the per-launch device times of one step (CSV) to this file')
    return ap.parse_args()


def describe(a, batch):
    name, image, classes, _, bits, gbits = WORKLOADS[a.workload]
    return {
        'workload': '%s %dx%dx3 batch %d/GPU, %d-bit dfxp W/A/G%s, stochastic rounding (Philox), SGD momentum 0.9 wd 2e-4'
                    % (name, image, image, batch, bits, '' if not gbits else ' (%d-bit G)' % gbits),
        'global_batch': batch * a.gpus,
        'parallelism': 'dp%d' % a.gpus,
        'l2': 'per-step working set (activations + gradients, >1 GB) exceeds the 126 MB L2; no explicit flush',
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path, all host threads
# ------------------------------------------------------------------------------------------------


def cpu_steps(a, batch, steps, warmup):
    import numpy as np
    import torch
    from oracle import dfxp as O
    name, image, classes, _, bits, gbits = WORKLOADS[a.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = dict(weight_decay=2e-4, noise=O.PhiloxNoise(0), seed=0, grad_bits=gbits)
    if name.startswith('Resnet'):
        kw.update(image=image, num_classes=classes)
    model = getattr(O, name)(bits, **kw)
    rng = np.random.default_rng(0)
    X = torch.from_numpy((rng.standard_normal((batch, image, image, 3)) * 0.5).astype(np.float32))
    y = torch.from_numpy(rng.integers(0, classes, batch))
    for _ in range(warmup):
        model.train_step(X, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.train_step(X, y)
    dt = (time.perf_counter() - t0) / max(1, steps)
    return batch / dt, dt, cores


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = a.batch or WORKLOADS[a.workload][3]
    cpu_batch = a.cpu_batch or (batch if a.workload in ('resnet20', 'cifar10') else 16)
    ips, dt, cores = cpu_steps(a, cpu_batch, a.steps, a.warmup)
    cfg = describe(a, batch)
    sample = '%d timed steps of the oracle port (torch-CPU fp32 restatement of dynamic_fixed_point.py) at batch %d' % (a.steps, cpu_batch)
    line = {
        'impl': 'reference', 'metric': 'quantized train imgs/sec', 'value': ips, 'unit': 'imgs/s', 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32 (fake-quant on CPU)', 'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {'value': ips, 'unit': 'imgs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': ips, 'unit': 'imgs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'TensorFlow 1.x reference cannot run in this image (SURVEY.md F10); this is the CPU oracle port',
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------


class Clocks:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------


def run_native(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lbt_b200 import _lib, models as M
    from lbt_b200.trainer import Trainer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a GPU: the DFXP path has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'        # NCCL's "NCCL version ..." banner goes to stdout: keep it to ONE JSON line
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert world == a.gpus, 'launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)' % (a.gpus, world)
    dev = torch.device('cuda', local)

    name, image, classes, dbatch, bits, gbits = WORKLOADS[a.workload]
    batch = a.batch or dbatch
    kw = dict(weight_decay=2e-4, grad_bits=gbits, seed=0)
    if name.startswith('Resnet'):
        kw.update(image=image, num_classes=classes)
    torch.manual_seed(0)
    model = getattr(M, name)(bits, **kw).to(dev)
    trainer = Trainer(model, lr=1e-2, momentum=0.9)

    # synthetic data: a small pool of pinned host batches (NHWC fp32, like the reference's feed_dict)
    rng = np.random.default_rng(1234 + rank)
    pool = 4
    hostX = [torch.from_numpy((rng.standard_normal((batch, image, image, 3)) * 0.5).astype(np.float32)).pin_memory()
             for _ in range(pool)]
    hosty = [torch.from_numpy(rng.integers(0, classes, batch)).pin_memory() for _ in range(pool)]
    devX = [x.to(dev) for x in hostX]
    devy = [y.to(dev) for y in hosty]
    Xs = torch.empty(batch, image, image, 3, device=dev)          # static step input (NHWC memory)
    ys = torch.empty(batch, dtype=torch.int64, device=dev)
    Xs.copy_(devX[0])
    ys.copy_(devy[0])
    X_view = Xs.permute(0, 3, 1, 2)                               # logical NCHW, channels_last storage
    h2d = Xs.numel() * 4 + ys.numel() * 8

    def eager_step():
        return trainer.step(X_view, ys)

    # ---- warm-up (eager) + per-step launch count -------------------------------------------------
    for _ in range(2):
        loss = eager_step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    loss = eager_step()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0

    # ---- capture the whole step in a CUDA graph --------------------------------------------------
    graph = None
    loss_static = None
    if not a.no_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    eager_step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss_static = eager_step()
            torch.cuda.synchronize()
        except Exception as e:  # report, fall back to eager launches (still the CUDA path)
            if rank == 0:
                print('[bench] CUDA graph capture failed (%s: %s); timing eager launches' % (type(e).__name__, e), file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
            return loss_static
        return eager_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- timed region 1: inputs resident in HBM ---------------------------------------------------
    for i in range(a.warmup):
        Xs.copy_(devX[i % pool]); ys.copy_(devy[i % pool])
        step()
    clocks = Clocks(local)
    barrier()
    if rank == 0:
        clocks.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        Xs.copy_(devX[i % pool]); ys.copy_(devy[i % pool])
        loss = step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    ms_step = ms_total / a.steps
    value = batch * world / (ms_step * 1e-3)
    final_loss = float(loss)

    # ---- timed region 2: end to end through the public API with HOST buffers ----------------------
    # every step: H2D copy of ITS inputs from pinned host memory (prefetched on a copy stream while the previous step runs,
    # lbt_b200.trainer.HostFeeder) and a D2H read of ITS loss (read by the host one step later) — all inside the timed region
    from lbt_b200.trainer import HostFeeder
    feeder = HostFeeder(step, Xs, ys)
    feeder.prefetch(hostX[0], hosty[0])
    for i in range(min(3, a.warmup)):
        feeder.prefetch(hostX[(i + 1) % pool], hosty[(i + 1) % pool])
        feeder.step()
    feeder.drain()
    # one batch is in flight from the warm-up; the timed loop copies exactly one batch per step
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    losses = []
    for i in range(a.steps):
        feeder.prefetch(hostX[(i + 1) % pool], hosty[(i + 1) % pool])      # H2D of a step's inputs from pinned memory
        losses.append(feeder.step())                                        # D2H of the previous step's loss
    losses.append(feeder.drain())                                           # ... and of the last one
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    lv = losses[-1]
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0)) / a.steps
    e2e_value = batch * world / (e2e_ms * 1e-3)

    # ---- live per-kernel timing: the step is captured once more with an event pair around every C-ABI launch
    # (external event nodes), so each replay times every kernel back to back on the device, without host gaps ----
    nprof = 3
    prof = None
    eager_ms = None
    try:
        prof = _lib.Profiler(external=True)
        s2 = torch.cuda.Stream()
        s2.wait_stream(torch.cuda.current_stream())
        pg = torch.cuda.CUDAGraph()
        _lib.profiler = prof
        with torch.cuda.graph(pg, stream=s2):
            eager_step()
            cal = [(prof.event(), prof.event()) for _ in range(4)]   # empty pairs: the cost of the event nodes themselves
            for ea, eb in cal:
                ea.record()
                eb.record()
        _lib.profiler = None
        torch.cuda.synchronize()
        acc = {}
        pg.replay()
        torch.cuda.synchronize()
        for _ in range(nprof):
            pg.replay()
            torch.cuda.synchronize()
            overhead = min(ea.elapsed_time(eb) for ea, eb in cal)
            for name, ea, eb, meta in prof.records:
                d = acc.setdefault(name, dict(launches=0, ms=0.0, bytes=0, ops=0))
                d['launches'] += 1
                d['ms'] += max(0.0, ea.elapsed_time(eb) - overhead)
                d['bytes'] += (meta or {}).get('bytes', 0)
                d['ops'] += (meta or {}).get('ops', 0)
        if a.dump_launches and rank == 0:
            with open(a.dump_launches, 'w') as f:
                f.write('# one %s step, per-launch device time from event nodes inside the replayed CUDA graph (event overhead %.2f us subtracted)\n'
                        % (a.workload, overhead * 1e3))
                f.write('index,entry,us,algorithmic_bytes,ops\n')
                for i, (name, ea, eb, meta) in enumerate(prof.records):
                    f.write('%d,%s,%.2f,%d,%d\n' % (i, name, max(0.0, ea.elapsed_time(eb) - overhead) * 1e3,
                                                   (meta or {}).get('bytes', 0), (meta or {}).get('ops', 0)))
        summ, timing_mode = acc, 'per-launch CUDA events inside a replayed graph of the step'
    except Exception as e:      # fall back to eager bracketing (includes host launch gaps on tiny kernels)
        _lib.profiler = None
        if rank == 0:
            print('[bench] graph-event profiling unavailable (%s: %s); eager events' % (type(e).__name__, e), file=sys.stderr)
        torch.cuda.synchronize()
        prof = _lib.Profiler()
        _lib.profiler = prof
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nprof):
            eager_step()
        e1.record()
        _lib.profiler = None
        summ, timing_mode = prof.summary(), 'per-launch CUDA events, eager launches'
        eager_ms = e0.elapsed_time(e1) / nprof
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_kind = json.load(open(peaks_path))['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_kind = 6650.0, 'fallback (B200_PROFILING.md)'
    ours_ms = sum(d['ms'] for d in summ.values()) / nprof
    breakdown = {}
    for k, d in sorted(summ.items()):
        row = {'launches': d['launches'] // nprof, 'ms_per_step': d['ms'] / nprof}
        if d['bytes'] and d['ms'] > 0:
            row['gbs'] = (d['bytes'] / 1e9) / (d['ms'] * 1e-3)
            row['frac_hbm'] = row['gbs'] / peak
        if d['ops'] and d['ms'] > 0:
            row['tops'] = (d['ops'] / 1e12) / (d['ms'] * 1e-3)
        breakdown[k] = row
    breakdown['other (torch pooling / loss / adds, gaps)'] = {'ms_per_step': max(0.0, ms_step - ours_ms)}
    # the dominant kernel class of the step (largest share of device time among the classes with a byte model)
    cands = [(d['ms'], k) for k, d in summ.items() if d['bytes'] and d['ms'] > 0]
    top = max(cands)[1] if cands else None
    q = summ.get(top, dict(launches=0, ms=0.0, bytes=0, ops=0))
    achieved = (q['bytes'] / 1e9) / (q['ms'] * 1e-3) if q['ms'] > 0 else 0.0
    gemm = summ.get('lbt_gemm_i8')
    gemm_tops = (gemm['ops'] / 1e12) / (gemm['ms'] * 1e-3) if gemm and gemm['ms'] > 0 else None

    if rank != 0:
        _finish(world)
        return

    # ---- CPU baseline (the oracle port on this box's host cores), N=1 only -------------------------
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cpu_batch = a.cpu_batch or (batch if a.workload in ('resnet20', 'cifar10') else 16)
        ips, dt, cores = cpu_steps(a, cpu_batch, 2, 1)
        cpu = {'value': ips, 'unit': 'imgs/s', 'cores': cores, 'kind': 'port',
               'sample': '2 timed training steps (after 1 warm-up) of the oracle port at batch %d, %.2f s/step' % (cpu_batch, dt)}

    line = {
        'metric': 'quantized train imgs/sec', 'value': value, 'unit': 'imgs/s', 'n_gpus': world, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 's8/u8 mantissas, s32 accumulate (8-bit dfxp); fp32 master weights', 'data': 'synthetic',
        'config': dict(describe(a, batch), exchange={
            'fused': 'lbt_dp_step: one kernel over NVLink peer memory (reduce-scatter by peer loads + SGD on the owned slice + '
                     'all-gather of the weights by peer stores + counter sum + range controller)',
            'nccl': 'NCCL all-reduce of gradients and counters + lbt_sgd_momentum + lbt_update_ranges',
            'unfused': 'single GPU: lbt_sgd_momentum + lbt_update_ranges'}[trainer.dp_mode] if world > 1 or trainer.dp_mode != 'fused'
            else 'single GPU: lbt_dp_step (SGD + range controller + step counter in one launch)'),
        'dp_error': trainer.dp.error() if trainer.dp is not None else None,
        'clocks': clk,
        'e2e': {'value': e2e_value, 'unit': 'imgs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                'ms_per_step': e2e_ms, 'wall_ms_per_step': wall / a.steps * 1e3,
                'how': 'HostFeeder: per-step H2D of the inputs prefetched on a copy stream, per-step D2H loss read one step later'},
        'gpu_launches': launches_per_step * a.steps,
        'launches_per_step': launches_per_step,
        'cuda_graph': graph is not None,
        'roofline': {'bound': 'hbm', 'kernel': top, 'achieved': achieved,
                     'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak if peak else None,
                     'traffic': TRAFFIC.get((a.workload, top)),
                     'peak_kind': peak_kind, 'launches_per_step': q['launches'] // nprof,
                     'avg_launch_us': q['ms'] / max(1, q['launches']) * 1e3,
                     'share_of_step': (q['ms'] / nprof) / ms_step if ms_step else None,
                     'algorithmic_bytes_per_launch': q['bytes'] / max(1, q['launches']),
                     'timing': timing_mode,
                     'note': 'tensors of this workload are 1-17 MB: the kernels are launch/latency bound, far below the HBM roofline'
                             if a.workload in ('resnet20', 'cifar10') else None},
        'gemm_tops_in_step': gemm_tops,
        'breakdown_ms': breakdown,
        'kernel_ms_per_step': ours_ms,
        'loss': final_loss,
        'cpu_baseline': cpu,
    }
    if a.breakdown:
        for k, d in breakdown.items():
            print('[breakdown] %-28s %s' % (k, d), file=sys.stderr)
    emit(line)
    _finish(world)


def _finish(world):
    """End of a rank.  With several ranks the NCCL communicator is referenced by the captured CUDA graph, and tearing
    the process group down under it has been seen to block forever (the measurement itself is complete): skip the
    teardown and leave with a hard exit once everything is flushed."""
    if world <= 1:
        return
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == '__main__':
    args = parse()
    guard_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)

#!/usr/bin/env python
"""bench.py -- Markov-GP inference-iteration throughput (time-steps/s) on B200.

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): MarkovVariationalGP, Matern-5/2 (d = 3),
Bernoulli-probit likelihood with 20-point Gauss-Hermite sites, N = 10^7 time steps, the temporally
parallel (scan) filter/smoother, fp64.  One "step" = one train_op-equivalent iteration
(SURVEY 3.1): model.inference(lr=1) [filter, smoother, site update, filter, smoother] followed by
model.energy() [expected log-lik, filter log-lik, expected pseudo log-lik].  value = N / time.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # CPU arm: plain-C port of the reference algorithm

With N > 1 (torchrun) the time axis is sharded: every rank holds --n-local steps (weak scaling) and
the ranks exchange the O(d^2) scan carries with NCCL all-gathers (bayesnewton_b200/distributed.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'markov_gp_inference_iter_time_steps_per_sec'
UNIT = 'time-steps/s'

# algorithmic bytes per time step of each kernel at d = 3, D = 1, fp64 (the
# reference-interface layouts of SURVEY 8d: full matrices, no As/Qs arrays, no gains)
ALGO_BYTES = {
    'up_reduce': 24,          # dt, pseudo_y, pseudo_var
    'up_filter': 24 + 96,     # dt, pseudo_y, pseudo_var in; m[3], P[3,3] out (kept packed, 72 B, in scratch)  (= F)
    'up_smooth': 8 + 96 + 16,  # dt', fm, fP in; H sm, H sP H^T out                                             (= S)
    'kf_reduce': 24, 'kf_apply': 24 + 96, 'kf_apply_ell': 24, 'rts_reduce': 8 + 96, 'rts_apply': 8 + 96 + 16,
    'site_update': 72,        # y, m, v, nat1, nat2 in; nat1, nat2, mean, cov out   (= U)
    'expected_density': 24,   # y, m, v                                   (= V)
    'gaussian_ell': 32,       # pseudo_y, m, v, pseudo_var                (= X)
    'energy_terms': 24 + 33,  # y, m, v + pseudo_y, pseudo_var (+mask): V and X in one pass
}


# fp64 instructions (DFMA + DMUL + DADD) per time step, from ncu's sass thread-instruction counters of the same
# kernels (profiles/r1g_ncu_up_kernels.csv: op counts x cycles / N); the fp64 roofline uses them
FP64_OPS = {'up_reduce': 179, 'up_filter': 163, 'up_smooth': 260}

# dram__bytes_read.sum + dram__bytes_write.sum per time step of one launch, from the ncu --set full captures
# under profiles/ (N = 1e7); reported as `traffic` (scaled by the steps a launch processes)
# dram__bytes_read.sum + dram__bytes_write.sum per time step of one launch at N = 1e7 (profiles/r2l_ncu_full_summary_c2.csv)
NCU_DRAM_BYTES = {'up_reduce': 26.0, 'up_filter': 97.1, 'up_smooth': 96.9, 'site_update': 67.3, 'energy_terms': 40.7}


def bench_inputs(N, seed=0):
    from _data import bench_inputs as f
    return f(N, seed)


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)"""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc, self.mark_ = index, [], None, 0

    def mark(self):
        """samples before this point (start-up of nvidia-smi, warm-up) are not counted"""
        self.mark_ = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines[self.mark_:]:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------ CPU arm
def run_cpu(sample_n, steps, warmup, threads=None):
    """the reference algorithm on the host: plain-C port (oracle/c/markov_c.c), OpenMP over the vmapped loops"""
    from oracle import cport
    if threads:
        os.environ['OMP_NUM_THREADS'] = str(threads)
    cores = int(os.environ.get('OMP_NUM_THREADS', os.cpu_count() or 1))
    t, dt, y = bench_inputs(sample_n)
    m = cport.ViModel(3, 1.0, 1.0, 2, 0.0, dt, y)
    for _ in range(warmup):
        m.iteration(1.0)
    t0 = time.perf_counter()
    for _ in range(steps):
        E = m.iteration(1.0)
    el = (time.perf_counter() - t0) / steps
    return {'value': sample_n / el, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'N=%d steps of the same workload, %d iteration(s); sequential filter/smoother (reference CPU '
                      'default parallel=False) with As/Qs materialised, site loops OpenMP x%d' % (sample_n, steps, cores),
            'ms_per_step': el * 1e3, 'energy': E}


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.cpu_sample
    # all the host threads the box has (torchrun exports OMP_NUM_THREADS=1 to its children)
    r = run_cpu(n, max(1, min(args.steps, 3)), min(args.warmup, 1), threads=os.cpu_count() or 1)
    line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, 1),
            'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def workload_config(args, world):
    return {'workload': 'C2: MarkovVariationalGP Matern52 (d=3) Bernoulli-probit GH-20, scan form, '
                        'N=%d time steps per GPU' % args.n_local,
            'n_time_steps_total': args.n_local * world, 'parallelism': 'time-shard x%d' % world,
            'iteration': 'inference(lr=1) [F,S,U,F,S] + energy() [V,L,X]',
            'l2': 'inputs larger than L2 (>= 2 GB working set vs 126 MB L2); no flush needed'}


# ------------------------------------------------------------------------------------------ GPU arm
def main_gpu(args):
    import torch
    import torch.distributed as dist
    import bayesnewton_b200 as bn
    from bayesnewton_b200 import _lib, distributed

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    NL = args.n_local

    # synthetic inputs: the global series is generated shard by shard with per-rank seeds
    t, dt, y = bench_inputs(NL, seed=100 * rank)
    if rank > 0:
        dt[0] = 0.1 + 0.2 * np.random.default_rng(7 + rank).random()
    nxt = torch.tensor([dt[0]], dtype=torch.float64, device=dev)
    if world > 1:  # dt of the right neighbour's first step closes this shard's smoother
        allfirst = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allfirst, nxt)
        dt_next = float(allfirst[rank + 1]) if rank + 1 < world else 0.0
    else:
        dt_next = 0.0
    dt_pin = torch.from_numpy(dt).pin_memory()
    y_pin = torch.from_numpy(y).pin_memory()

    kern = bn.kernels.Matern52(variance=1.0, lengthscale=1.0)
    lik = bn.likelihoods.Bernoulli(link='probit')
    model = distributed.TimeShardedMarkovGP(kern, lik, dt_pin, y_pin, dt_next, _lib.BN_METHOD_VI, rank, world)
    L = _lib.lib()

    def step():
        model.inference(lr=1.0)
        return model.energy()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs ~0.1 s to come up: start it before the warm-up, count only what it samples from the first
    # timed region on (the device-timed steps, then the end-to-end and with-gradient steps: all under the same load)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    # ---- timed region: K steps, device-resident inputs, CUDA events on the launching stream
    sync_all()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        E = step()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    # ---- the same K steps once more with every kernel launch bracketed by CUDA events on its stream (the per-kernel
    # durations behind the roofline); kept out of the region above because the 60 event pairs per step cost ~5 %
    L.bn_timing_enable(1)
    for _ in range(args.steps):
        E = step()
    sync_all()
    import ctypes
    cbuf = ctypes.create_string_buffer(8192)
    L.bn_timing_report(cbuf, 8192)
    L.bn_timing_enable(0)
    kt = {}
    for ln in cbuf.value.decode().splitlines():
        name, cnt, tot = ln.split()
        kt[name] = (int(cnt), float(tot))
    energy = float(E)

    # ---- end to end through the public API with HOST buffers: every step, that step's inputs (dt, Y) come from
    # pinned host memory and the result (energy, a double) goes back to the host.  The copy of step i+1's inputs
    # runs on a second stream into the other device buffer while step i computes (double buffering).
    copy_stream = torch.cuda.Stream()
    bufs = [(model.shard.dt, model.Y), (torch.empty_like(model.shard.dt), torch.empty_like(model.Y))]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def start_copy(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[b])      # the step that last used this buffer has finished
            bufs[b][0].copy_(dt_pin, non_blocking=True)
            bufs[b][1].copy_(y_pin, non_blocking=True)
            ready[b].record(copy_stream)

    # the step's result goes back through a pinned double buffer: the copy of step i's energy is queued behind step
    # i's kernels and read on the host while step i+1 is already running (one D2H read per step, never skipped)
    e_pin = torch.zeros(2, dtype=torch.float64).pin_memory()
    e_done = [torch.cuda.Event(), torch.cuda.Event()]
    energies = []

    def e2e_step(i):
        b = i % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[b])
        model.shard.dt, model.Y = bufs[b]
        start_copy(i + 1)                        # H2D of the next step's inputs overlaps this step's kernels
        model.inference(lr=1.0)
        e = model.energy()
        free[b].record(cur)
        e_pin[b:b + 1].copy_(e.reshape(1), non_blocking=True)   # D2H read of this step's result
        e_done[b].record(cur)
        if i > 0:                                # the previous step's result, on the host
            e_done[b ^ 1].synchronize()
            energies.append(float(e_pin[b ^ 1]))

    for b in range(2):
        free[b].record(torch.cuda.current_stream())
    start_copy(0)
    e2e_step(0)
    sync_all()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        e2e_step(i + 1)
    f1.record()
    e_done[args.steps % 2].synchronize()         # the last step's result
    energies.append(float(e_pin[args.steps % 2]))
    sync_all()
    copy_stream.synchronize()
    ms2 = torch.tensor([max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2) / args.steps

    # ---- the same iteration with the hyper-gradient pass G ("reported with and without"): the closing
    # posterior update of inference() accumulates d log-lik / d (variance, lengthscale) inside its smoother sweep
    model.shard.dt, model.Y = bufs[0]

    def step_grad():
        model.inference(lr=1.0, want_grad=True)
        return model.energy_and_grad()

    for _ in range(2):
        step_grad()
    sync_all()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        Eg, dEg = step_grad()
    g1.record()
    sync_all()
    ms3 = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
    grad_ms = float(ms3) / args.steps
    clocks = sampler.stop()
    cbuf2 = ctypes.create_string_buffer(8192)

    # fp64 FMA peak of this device, measured now (the second roofline: at d = 3 the path is fp64-pipe bound)
    import ctypes as _C
    scratch = torch.empty(8 * 148 * 256 * 2, dtype=torch.float64, device=dev)
    dfma = _C.c_double(0.0)
    L.bn_measure_dfma_peak(scratch.data_ptr(), scratch.numel(), _C.byref(dfma))
    dfma_peak = dfma.value

    if rank == 0:
        total_steps = NL * world
        peak, peak_src = measured_peaks()
        # dominant kernel by device time inside the timed region
        dom = max(kt, key=lambda k: kt[k][1]) if kt else None
        roof = None
        if dom:
            cnt, tot = kt[dom]
            avg_ms = tot / cnt
            ab = ALGO_BYTES.get(dom, 0)
            if dom == 'kf_apply':  # 2 of 3 launches per step write the states, 1 is log-likelihood only
                ab = (2 * ALGO_BYTES['kf_apply'] + ALGO_BYTES['kf_apply_ell']) / 3.0
            achieved = ab * NL / (avg_ms * 1e-3) / 1e9
            traffic = NCU_DRAM_BYTES[dom] * NL if (args.traffic is None and dom in NCU_DRAM_BYTES) else args.traffic
            roof = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                    'traffic_source': 'ncu --set full dram bytes per step at N=1e7 (profiles/r2l_ncu_full_summary_c2.csv) x steps per launch',
                    'algorithmic_bytes_per_launch': ab * NL, 'avg_launch_ms': avg_ms,
                    'share_of_step': tot / (ms_per_step * args.steps),
                    'timing': 'CUDA events around every launch of this kernel, on its stream, in a second pass of the same K steps right after the timed region'}
            if dom in FP64_OPS and dfma_peak > 0:
                f64 = FP64_OPS[dom] * NL / (avg_ms * 1e-3)
                roof['fp64'] = {'achieved_fp64_inst_per_s': f64, 'peak_dfma_per_s': dfma_peak, 'frac': f64 / dfma_peak,
                                'note': 'the kernel is bound by the fp64 pipe, not HBM (DESIGN.md section 4)'}
        line = {
            'metric': METRIC, 'value': total_steps / (ms_per_step * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, world),
            'clocks': clocks,
            'e2e': {'value': total_steps / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': int(2 * NL * 8), 'd2h_bytes_per_step': 8,
                    'h2d_GBps_per_gpu': 2 * NL * 8 / (e2e_ms * 1e-3) / 1e9,
                    'note': 'inputs are the reference-facing fp64 host arrays (dt, Y): 16 B per time step over PCIe every step, '
                            'double-buffered against compute; when h2d_GBps_per_gpu is near the link rate the end-to-end figure is copy-bound'},
            'gpu_launches': int(sum(c for c, _ in kt.values())),
            'roofline': roof,
            'iteration_bytes': {'algorithmic_bytes_per_time_step': 636,
                                'achieved_GBps': 636 * NL / (ms_per_step * 1e-3) / 1e9,
                                'frac_of_hbm_peak': 636 * NL / (ms_per_step * 1e-3) / 1e9 / peak},
            'kernels_ms_per_step': {k: v[1] / args.steps for k, v in kt.items()},
            'energy': energy,
            'with_hyper_gradient': {'ms_per_step': grad_ms, 'value': total_steps / (grad_ms * 1e-3), 'unit': UNIT,
                                    'algorithmic_bytes_per_time_step': 636,
                                    'note': 'iteration + d energy / d (variance, lengthscale); the adjoint is formed inside '
                                            'the smoother sweep, no extra HBM pass',
                                    'd_energy': [float(v) for v in dEg.reshape(-1).tolist()]},
            'fp64_peak_dfma_per_s': dfma_peak,
            'fp64_frac_by_kernel': {k: FP64_OPS[k] * NL * kt[k][0] / (kt[k][1] * 1e-3) / dfma_peak
                                    for k in FP64_OPS if k in kt and dfma_peak > 0},
        }
        if world == 1 and not args.no_cpu:
            r = run_cpu(args.cpu_sample, 1, 1)
            line['cpu_baseline'] = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--n-local', type=int, default=10_000_000, help='time steps per GPU (C2: 1e7)')
    ap.add_argument('--cpu-sample', type=int, default=2_000_000, help='time steps of the bounded CPU sample')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--traffic', type=float, default=None, help='dram bytes per launch of the dominant kernel (ncu)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == '__main__':
    main()

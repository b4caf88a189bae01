#!/usr/bin/env python
"""bench.py -- Markov-GP inference-iteration throughput (time-steps/s) on B200.

Default workload (BASELINE.json configs[4] / the north_star target; SURVEY 8d "C5"): MarkovVariationalGP,
Matern-5/2 (d = 3), Bernoulli-probit likelihood with 20-point Gauss-Hermite sites, N = 10^8 time steps, the temporally
parallel (scan) filter / smoother, fp64.  One "step" = one train_op-equivalent iteration:
model.inference(lr=1) [filter, smoother, site update, filter, smoother] followed by model.energy() [expected
log-lik, filter log-lik, expected pseudo log-lik].  value = N / time.

  python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path (C5; N GPUs share the SAME 10^8 steps)
  python bench.py --workload C2 ...                          # configs[1]: the same model at N = 10^7
  python bench.py --impl reference ...                       # CPU arm: plain-C port of the reference algorithm

With N > 1 (torchrun) the time axis of the same series is sharded over the ranks (strong scaling); the ranks exchange
the O(d^2) scan carries over NVLink peer memory / NCCL (bayesnewton_b200/distributed.py).
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

WORKLOADS = {
    'C5': dict(n_total=100_000_000, name='C5: MarkovVariationalGP Matern52 (d=3) Bernoulli-probit GH-20, scan form, N=1e8'),
    'C2': dict(n_total=10_000_000, name='C2: MarkovVariationalGP Matern52 (d=3) Bernoulli-probit GH-20, scan form, N=1e7'),
    'C1': dict(n_total=1_000, name='C1: demos/regression.py shape -- MarkovVariationalGP Matern52 (d=3) Gaussian, N=1e3, lr=1'),
    'C3': dict(n_total=1_000_000, name='C3: MarkovVariationalGP Independent[Matern32 x2] HeteroscedasticNoise (GH 20x20), N=1e6, lr=0.3'),
    'C4': dict(n_total=10_000, name='C4: MarkovVariationalGP SpatioTemporalKernel(Matern32 time x separable Matern32 space), '
                                    'N_t=1e4 x 16x16 grid (M=256, d=512), Gaussian, 5% missing, lr=1'),
}

# algorithmic bytes per time step at d = 3, D = 1, fp64 on the reference-interface layouts (full matrices,
# no As/Qs arrays, no gains): F = 121, S = 120, U = 72, V = 24, X = 33, L = 25; iteration = 2 (F + S) + U + V + X + L = 636.
# Per kernel of the fused path: the reduce pass re-reads the filter's inputs (24), the filter pass is F, the two
# smoother sweeps carry U resp. V + X in their epilogue.
ALGO_BYTES = {
    'it_reduce': 24, 'it_filter': 121, 'it_smooth_sites': 120 + 72, 'it_smooth_energy': 120 + 24 + 33, 'it_smooth': 120,
    'up_reduce': 24, 'up_filter': 121, 'up_smooth': 120, 'up_smooth_grad': 120, 'site_update': 72, 'energy_terms': 57,
    'it_from_tiled': 16, 'it_to_tiled': 16,
}
# dram__bytes_read.sum + dram__bytes_write.sum per TIME STEP of the two dominant kernels, from the ncu --set full capture
# of this path at N = 1e8 (profiles/r4b_ncu_full_summary_c5_n1e8.csv): 12.00 GB and 13.55 GB per launch
NCU_DRAM_BYTES_PER_STEP = {'it_smooth_sites': 120.0, 'it_smooth_energy': 135.5}
ITER_BYTES = 636           # per time step, SURVEY 8d
ITER_BYTES_EXECUTED = 611  # without L: compute_log_lik() is served from the filter pass of the closing update_posterior()


def block_seeded_inputs(n_total, lo, hi, seed=0, block=1 << 20):
    """C2 / C5 inputs of SURVEY 8d for the steps [lo, hi) of the global series: dt_0 = 0, dt_k = 0.1 + 0.2 u_k,
    y_k = 1[2 sin(.3 t_k) + sin(.05 t_k) + .5 eps_k > 0].  Random numbers are drawn per block of 2^20 steps from a
    generator seeded by the block index, so every rank can form ITS shard of the same global series."""
    b0, b1 = lo // block, (hi + block - 1) // block
    t_off = 0.0
    for b in range(b0):  # time offset of the first block: sum of the step lengths before it
        n = min(block, n_total - b * block)
        dtb = 0.1 + 0.2 * np.random.default_rng([seed, b]).random(n)
        if b == 0:
            dtb[0] = 0.0
        t_off += float(dtb.sum())
    dts, ys = [], []
    for b in range(b0, b1):
        n = min(block, n_total - b * block)
        rng = np.random.default_rng([seed, b])
        dtb = 0.1 + 0.2 * rng.random(n)
        if b == 0:
            dtb[0] = 0.0
        eps = rng.standard_normal(n)
        t = t_off + np.cumsum(dtb)
        t_off = float(t[-1])
        yb = (2 * np.sin(0.3 * t) + np.sin(0.05 * t) + 0.5 * eps > 0).astype(np.uint8)
        s, e = max(lo, b * block) - b * block, min(hi, b * block + n) - b * block
        dts.append(dtb[s:e])
        ys.append(yb[s:e])
    return np.concatenate(dts), np.concatenate(ys)


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


def workload_config(args, world, n_run=None):
    w = WORKLOADS[args.workload]
    cfg = {'workload': w['name'], 'n_time_steps_total': int(args.n_total),
           'parallelism': 'time-shard x%d (the same series split over the GPUs)' % world,
           'iteration': 'inference(lr=1) [F,S,U,F,S] + energy() [V,L,X]',
           'l2': 'inputs larger than L2 (>= 2 GB working set vs 126 MB L2); no flush needed'}
    if n_run is not None and n_run != args.n_total:
        cfg['workload'] += ' -- bounded sample: N=%d steps of it' % n_run
        cfg['n_time_steps_run'] = int(n_run)
    return cfg


# ------------------------------------------------------------------------------------------ CPU arm
def probe_reference_runtime():
    """the reference itself (pure Python on jax 0.4.14 + objax) can only be timed if those packages exist on the box"""
    import importlib.util
    missing = [m for m in ('jax', 'objax') if importlib.util.find_spec(m) is None]
    ref = os.path.join(ROOT, 'baseline', '_ref')
    return {'reference_runnable': not missing and os.path.isdir(ref), 'missing_modules': missing,
            'baseline_ref_present': os.path.isdir(ref)}


def run_cpu(sample_n, steps, warmup, threads=None):
    """the reference algorithm on the host (oracle/c/markov_c.c), on all the host threads it can use.  Two forms are
    tried for one iteration each and the faster one is timed: the reference's CPU default (parallel=False: sequential
    lax.scan recursions on one thread, the vmapped loops on all threads) and the temporally parallel form
    (parallel=True, ops.py:183-253, 314-354) as a time-blocked three-phase scan over all threads."""
    from oracle import cport
    if threads:
        os.environ['OMP_NUM_THREADS'] = str(threads)
    cores = int(os.environ.get('OMP_NUM_THREADS', os.cpu_count() or 1))
    dt, y = block_seeded_inputs(sample_n, 0, sample_n)
    m = cport.ViModel(3, 1.0, 1.0, 2, 0.0, dt, y.astype(np.float64))
    forms = {'sequential filter/smoother (reference CPU default parallel=False), vmapped loops OpenMP x%d' % cores: lambda: m.iteration(1.0),
             'temporally parallel form (parallel=True) as a time-blocked three-phase scan on %d threads' % cores: lambda: m.iteration_blocked(1.0)}
    m.iteration(1.0)  # first touch of the scratch arrays
    trial = {}
    for name, fn in forms.items():
        t0 = time.perf_counter()
        fn()
        trial[name] = time.perf_counter() - t0
    form = min(trial, key=trial.get)
    it = forms[form]
    for _ in range(max(0, warmup - 3)):
        it()
    t0 = time.perf_counter()
    for _ in range(steps):
        E = it()
    el = (time.perf_counter() - t0) / steps
    return {'value': sample_n / el, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'N=%d steps of the same workload, %d timed iteration(s); %s (the faster of the two forms: %s); As/Qs '
                      'materialised as in the reference' % (sample_n, steps, form, ', '.join('%.0f ms' % (v * 1e3) for v in trial.values())),
            'ms_per_step': el * 1e3, 'energy': E, 'steps': steps, 'warmup': warmup}


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.cpu_sample
    # every step is a bounded sample of the workload; all the host threads the box has (torchrun exports OMP_NUM_THREADS=1)
    r = run_cpu(n, args.steps, args.warmup, threads=os.cpu_count() or 1)
    line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, 1, n_run=n),
            'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'reference_runtime_probe': probe_reference_runtime(),
            'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def numa_local_affinity(local_rank):
    """bind this process to the cores of the NUMA node its GPU hangs off, so pinned host buffers are allocated (first
    touch) in the memory next to the PCIe root of that GPU"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        devs = [d for d in os.listdir('/sys/bus/pci/devices') if d.lower().startswith('%04x:%02x:' % (dom, bus))]
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % devs[0]).read())
        if node < 0:
            return None
        cpus = open('/sys/devices/system/node/node%d/cpulist' % node).read().strip()
        ids = []
        for part in cpus.split(','):
            a, _, b = part.partition('-')
            ids += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return node
    except Exception:
        return None


def parity_self_check(bn, distributed, _lib, rank, world, dev):
    """N > 1: a small fixed global series sharded over all ranks must reproduce rank 0's single-GPU result"""
    import torch
    import torch.distributed as dist
    n = 40_000 * world + 17
    dt, y = block_seeded_inputs(n, 0, n, seed=11)
    b = [n * r // world for r in range(world + 1)]
    kern = bn.kernels.Matern52(1.0, 1.0)
    lik = bn.likelihoods.Bernoulli(link='probit')
    nxt = dt[b[rank + 1]] if rank + 1 < world else 0.0
    m = distributed.TimeShardedMarkovGP(kern, lik, dt[b[rank]:b[rank + 1]].copy(), y[b[rank]:b[rank + 1]].astype(np.float64),
                                        nxt, _lib.BN_METHOD_VI, rank, world)
    for _ in range(2):
        m.inference(lr=0.7)
    E = m.energy()
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    if rank == 0:
        one = bn.models.MarkovVariationalGP(kernel=kern, likelihood=lik, X=np.cumsum(dt), Y=y.astype(np.float64), parallel=True)
        for _ in range(2):
            one.inference(lr=0.7)
        E1 = one.energy()
        pm1 = one.posterior_mean.reshape(-1)[b[0]:b[1]]
        err = float((m.posterior_mean.reshape(-1) - pm1).abs().max() / pm1.abs().max())
        eerr = abs(float(E) - float(E1)) / abs(float(E1))
        ok[0] = int(err < 1e-9 and eerr < 1e-9)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(int(ok))


def main_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    numa = numa_local_affinity(local_rank) if world > 1 else None
    import bayesnewton_b200 as bn
    from bayesnewton_b200 import _lib, distributed
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    NT = args.n_total
    lo, hi = NT * rank // world, NT * (rank + 1) // world
    NL = hi - lo
    L = _lib.lib()

    parity_ok = None
    if world > 1:
        parity_ok = parity_self_check(bn, distributed, _lib, rank, world, dev)

    # synthetic inputs: this rank's shard of the global series (pinned host buffers: dt fp64, labels uint8)
    dt, y8 = block_seeded_inputs(NT, lo, hi)
    dt_pin = torch.from_numpy(dt).pin_memory()
    y_pin = torch.from_numpy(y8).pin_memory()
    kern = bn.kernels.Matern52(variance=1.0, lengthscale=1.0)
    lik = bn.likelihoods.Bernoulli(link='probit')
    if world == 1:
        # the class the reference API names; X = cumulative time, Y = the labels
        model = bn.models.MarkovVariationalGP(kernel=kern, likelihood=lik, X=np.cumsum(dt), Y=y8.astype(np.float64), parallel=True)
        model_name = 'bayesnewton_b200.models.MarkovVariationalGP'
    else:
        first = torch.tensor([dt[0]], dtype=torch.float64, device=dev)
        allfirst = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allfirst, first)
        dt_next = float(allfirst[rank + 1]) if rank + 1 < world else 0.0
        model = distributed.TimeShardedMarkovGP(kern, lik, dt_pin, torch.from_numpy(y8.astype(np.float64)), dt_next,
                                                _lib.BN_METHOD_VI, rank, world)
        model_name = 'bayesnewton_b200.distributed.TimeShardedMarkovGP'
    del dt, y8

    def step():
        model.inference(lr=1.0)
        return model.energy()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    # durations behind the roofline); kept out of the region above because the event pairs cost a few per cent
    import ctypes
    L.bn_timing_enable(1)
    for _ in range(args.steps):
        E = step()
    sync_all()
    cbuf = ctypes.create_string_buffer(16384)
    L.bn_timing_report(cbuf, 16384)
    L.bn_timing_enable(0)
    kt = {}
    for ln in cbuf.value.decode().splitlines():
        name, cnt, tot = ln.split()
        kt[name] = (int(cnt), float(tot))
    energy = float(E)

    # ---- end to end through the public API with HOST buffers: every step, that step's inputs (dt fp64, labels uint8)
    # come from pinned host memory (streamed in by model.load_inputs) and the result (energy, a double) goes back to the
    # host.  The copy of step i+1's inputs runs on a second stream into the other staging buffer while step i computes.
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty(NL, dtype=torch.float64, device=dev), torch.empty(NL, dtype=torch.uint8, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def start_copy(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[b])      # the step that last used this buffer has finished
            bufs[b][0].copy_(dt_pin, non_blocking=True)
            bufs[b][1].copy_(y_pin, non_blocking=True)
            ready[b].record(copy_stream)

    e_pin = torch.zeros(2, dtype=torch.float64).pin_memory()
    e_done = [torch.cuda.Event(), torch.cuda.Event()]
    energies = []

    def e2e_step(i):
        b = i % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[b])
        start_copy(i + 1)                        # H2D of the next step's inputs overlaps this step's kernels
        model.load_inputs(bufs[b][0], bufs[b][1])
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
    del bufs

    # ---- the same iteration with the hyper-gradient pass G ("reported with and without"): the closing
    # posterior update of inference() accumulates d log-lik / d (variance, lengthscale) inside its smoother sweep
    grad = None
    if not args.no_grad:
        def step_grad():
            model.inference(lr=1.0, want_grad=True)
            return model.energy_and_grad()

        for _ in range(2):
            step_grad()
        sync_all()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gsteps = max(1, min(args.steps, 5))
        g0.record()
        for _ in range(gsteps):
            Eg, dEg = step_grad()
        g1.record()
        sync_all()
        ms3 = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        grad = (float(ms3) / gsteps, [float(v) for v in dEg.reshape(-1).tolist()])
    # ---- the fp32 build of the same fused iteration (bn_iter_*_f32), reported beside the fp64 headline, never instead of it
    fp32 = None
    if world == 1 and not args.no_fp32:
        from bayesnewton_b200 import fused
        sh = fused.FusedShard(kern, dt_pin.to(dev), y_pin.to(dev).to(torch.float64), dtype=torch.float32)
        sh.load_sites(torch.zeros(NL, device=dev), torch.full((NL,), 100.0, device=dev))

        def step32():
            sh.run(fused.SITES, lik, _lib.BN_METHOD_VI, None, 1.0, 1.0, True, want_ell=False)
            return sh.run(fused.ENERGY, lik, _lib.BN_METHOD_VI, None, 1.0, 1.0, True)
        for _ in range(args.warmup):
            step32()
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(args.steps):
            ell32, s32 = step32()
        h1.record()
        torch.cuda.synchronize()
        ms32 = h0.elapsed_time(h1) / args.steps
        # the same number of iterations from the same initial sites through the fp64 build of the same entry points
        pm32, pv32 = sh.posterior()
        pm32, pv32 = pm32.reshape(-1).double(), pv32.reshape(-1).double()
        del sh
        sh = fused.FusedShard(kern, dt_pin.to(dev), y_pin.to(dev).to(torch.float64))
        sh.load_sites(torch.zeros(NL, device=dev), torch.full((NL,), 100.0, device=dev))
        for _ in range(args.warmup + args.steps):
            ell64, s64 = step32()
        pm64, pv64 = sh.posterior()
        pm64, pv64 = pm64.reshape(-1), pv64.reshape(-1)
        fp32 = {'ms_per_step': ms32, 'value': NT / (ms32 * 1e-3), 'unit': UNIT, 'dtype': 'f32',
                'algorithmic_bytes_per_time_step': ITER_BYTES // 2,
                'note': 'storage and arithmetic in fp32 (sums in fp64, 8 KB cubic probit table); parity bar 1e-4 against fp64 '
                        '(tests/test_fp32_mode.py); two fused passes through FusedShard, no model-level host code in the loop',
                'vs_fp64_after_%d_iterations' % (args.warmup + args.steps): {
                    'post_mean_rel': float((pm32 - pm64).abs().max() / pm64.abs().max()),
                    'post_var_rel': float((pv32 - pv64).abs().max() / pv64.abs().max()),
                    'log_lik_rel': abs(float(ell32) - float(ell64)) / abs(float(ell64))}}
        del sh, pm32, pv32, pm64, pv64
        torch.cuda.empty_cache()
    clocks = sampler.stop()

    # fp64 FMA peak of this device, measured now (the second roofline: at d = 3 the update kernels are fp64-pipe bound)
    scratch = torch.empty(8 * 148 * 256 * 2, dtype=torch.float64, device=dev)
    dfma = ctypes.c_double(0.0)
    L.bn_measure_dfma_peak(scratch.data_ptr(), scratch.numel(), ctypes.byref(dfma))
    dfma_peak = dfma.value

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel by device time inside the timed region
        dom = max(kt, key=lambda k: kt[k][1]) if kt else None
        roof = None
        if dom:
            cnt, tot = kt[dom]
            avg_ms = tot / cnt
            ab = ALGO_BYTES.get(dom, 0)
            achieved = ab * NL / (avg_ms * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak,
                    'traffic': args.traffic if args.traffic else (NCU_DRAM_BYTES_PER_STEP[dom] * NL if dom in NCU_DRAM_BYTES_PER_STEP else None),
                    'peak_source': peak_src,
                    'traffic_source': ('--traffic' if args.traffic else 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of this kernel at '
                                       'N = 1e8 (profiles/r4b_ncu_full_summary_c5_n1e8.csv), scaled per time step to this launch; below '
                                       'the algorithmic bytes because the filtered states travel packed (72 B, not the 96 B of the '
                                       'reference layout) and the marginals of the site pass are never stored') if (args.traffic or dom in NCU_DRAM_BYTES_PER_STEP) else None,
                    'algorithmic_bytes_per_step': ab, 'algorithmic_bytes_per_launch': ab * NL, 'avg_launch_ms': avg_ms,
                    'share_of_step': tot / (ms_per_step * args.steps),
                    'timing': 'CUDA events around every launch of this kernel, on its stream, in a second pass of the same K steps right after the timed region'}
        line = {
            'metric': METRIC, 'value': NT / (ms_per_step * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(workload_config(args, world), model=model_name, n_time_steps_per_gpu=int(NL)),
            'clocks': clocks,
            'e2e': {'value': NT / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': int(NL * 9), 'd2h_bytes_per_step': 8,
                    'h2d_GBps_per_gpu': NL * 9 / (e2e_ms * 1e-3) / 1e9, 'numa_node_of_pinned_buffers': numa,
                    'note': 'inputs stream from pinned host memory every step through model.load_inputs(): dt as fp64 and the '
                            'Bernoulli labels as uint8 (9 B per time step over PCIe), double-buffered against compute; when '
                            'h2d_GBps_per_gpu is near the link rate the end-to-end figure is copy-bound'},
            'gpu_launches': int(sum(c for c, _ in kt.values())) // max(1, args.steps) * args.steps,
            'gpu_launches_per_step': int(sum(c for c, _ in kt.values())) // max(1, args.steps),
            'roofline': roof,
            'iteration_bytes': {'algorithmic_bytes_per_time_step': ITER_BYTES,
                                'achieved_GBps': ITER_BYTES * NT / world / (ms_per_step * 1e-3) / 1e9,
                                'frac_of_hbm_peak': ITER_BYTES * NT / world / (ms_per_step * 1e-3) / 1e9 / peak,
                                'executed_bytes_per_time_step': ITER_BYTES_EXECUTED,
                                'frac_of_hbm_peak_executed': ITER_BYTES_EXECUTED * NT / world / (ms_per_step * 1e-3) / 1e9 / peak,
                                'note': 'per GPU.  636 B = 2(F+S)+U+V+X+L on the reference-interface layouts (SURVEY 8d); L (25 B, '
                                        'compute_log_lik) is NOT executed as a pass of its own: its value comes from the filter of the '
                                        'closing update_posterior() of the same iteration (same inputs), so 611 B is the executed count'},
            'kernels_ms_per_step': {k: v[1] / args.steps for k, v in kt.items()},
            'energy': energy,
            'fp64_peak_dfma_per_s': dfma_peak,
        }
        if parity_ok is not None:
            line['parity_nranks_ok'] = parity_ok
        if grad is not None:
            line['with_hyper_gradient'] = {'ms_per_step': grad[0], 'value': NT / (grad[0] * 1e-3), 'unit': UNIT,
                                           'note': 'iteration + d energy / d (variance, lengthscale); the adjoint is formed inside '
                                                   'the smoother sweep, no extra HBM pass', 'd_energy': grad[1]}
        if fp32 is not None:
            line['fp32_mode'] = fp32
        if world == 1 and not args.no_cpu:
            r = run_cpu(args.cpu_sample, 1, 1)
            line['cpu_baseline'] = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------------------ C1 / C3 / C4
def build_generic(args, bn, world, rank):
    """(model, step(), host inputs to stream in the end-to-end leg, description of the roofline)"""
    from bayesnewton_b200 import _lib
    K = bn.kernels
    w = args.workload
    if w == 'C1':
        N = args.n_total
        x = np.linspace(-17, 147, N)
        y = np.cos(0.04 * x + 0.33 * np.pi) * np.sin(0.2 * x) + np.sqrt(0.2) * np.random.default_rng(12345).standard_normal(N)
        m = bn.models.MarkovVariationalGP(kernel=K.Matern52(1.0, 5.0), likelihood=bn.likelihoods.Gaussian(0.2), X=x, Y=y, parallel=True)
        step = lambda: (m.inference(lr=1.0), m.energy())[1]
        return m, step, (np.concatenate([[0.0], np.diff(x)]), y), dict(bound='launch latency', bytes_per_step=ITER_BYTES)
    if w == 'C3':
        N = args.n_total
        dt = 0.05 + 0.1 * np.random.default_rng(0).random(N)
        dt[0] = 0
        t = np.cumsum(dt)
        y = np.sin(0.5 * t) + np.log1p(np.exp(np.cos(0.2 * t))) * np.random.default_rng(1).standard_normal(N)
        y = (y - y.mean()) / y.std()
        kern = K.Independent([K.Matern32(1.0, 1.0), K.Matern32(1.0, 1.0)])
        lik = bn.likelihoods.HeteroscedasticNoise()
        if world == 1:
            m = bn.models.MarkovVariationalGP(kernel=kern, likelihood=lik, X=t, Y=y, parallel=True)
        else:
            from bayesnewton_b200 import latent_sharding as ls
            m = ls.LatentShardedMarkovGP(kern, lik, t, y, _lib.BN_METHOD_VI, rank, world, power=1.0)
        step = lambda: (m.inference(lr=0.3), m.energy())[1]
        # F 216, S 216, U 200, V 56, X 96, L 56 => 1272 B per step; the 400-point cubature makes U, V fp64-bound
        return m, step, (dt, y), dict(bound='fp64 (400-point cubature) / hbm', bytes_per_step=1272,
                                      kernel_bytes={'up_reduce': 56, 'up_filter': 216, 'up_smooth': 216, 'site_update': 200,
                                                    'expected_density': 56, 'gaussian_ell': 96})
    if w == 'C4':
        Nt, G = args.n_total, 16
        t = np.arange(Nt, dtype=np.float64)
        a = np.linspace(-3, 3, G)
        r = np.array([[u, v] for u in a for v in a])
        R = np.tile(r[None], (Nt, 1, 1))
        Y = (np.sin(t / 10)[:, None] + np.sin(r[:, 0])[None] + np.cos(r[:, 1])[None]
             + 0.1 * np.random.default_rng(1).standard_normal((Nt, G * G)))
        Y[np.random.default_rng(2).uniform(size=Y.shape) < 0.05] = np.nan
        kern = bn.spacetime.SpatioTemporalKernel(K.Matern32(1.0, 5.0), bn.spacetime.Separable([K.Matern32(1.0, 1.0), K.Matern32(1.0, 1.0)]), z=r)
        m = bn.models.MarkovVariationalGP(kernel=kern, likelihood=bn.likelihoods.Gaussian(1.0), X=t, Y=Y, R=R)
        step = lambda: (m.inference(lr=1.0), m.energy())[1]
        M = G * G
        d = 2 * M
        fF = M ** 3 / 3 + 2 * M * M * d + 2 * d * d * M + 8 * d * d       # 0.209 GFLOP per filter step
        fS = d ** 3 / 3 + 2 * d ** 3 + 4 * d ** 3 + 8 * d * d             #            0.852 GFLOP per smoother step
        return m, step, (None, Y), dict(bound='fp64_tensor', flops_per_step=3 * fF + 2 * fS, flops_filter=fF, flops_smoother=fS)
    raise ValueError(w)


def main_generic(args):
    import ctypes
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    import bayesnewton_b200 as bn
    from bayesnewton_b200 import _lib
    if world > 1:
        if args.workload != 'C3' or world != 2:
            raise SystemExit('%s runs on one GPU (C3 also latent-sharded on 2)' % args.workload)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    L = _lib.lib()
    model, step, (dt_h, y_h), roofdesc = build_generic(args, bn, world, rank)
    NT = args.n_total

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        E = step()
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
    # launch-bound workloads: the same step captured once in a CUDA graph and replayed (C1)
    graph_ms = None
    if args.workload == 'C1':
        try:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                step()
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                Eg = step()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(args.steps):
                g.replay()
            g1.record()
            torch.cuda.synchronize()
            graph_ms = g0.elapsed_time(g1) / args.steps
        except Exception as ex:  # noqa: BLE001 -- capture is an optimisation: report why it was not available
            graph_ms = 'unavailable: ' + repr(ex)[:200]
            torch.cuda.synchronize()
    # per-kernel device times (second pass, as in the C5 arm)
    L.bn_timing_enable(1)
    for _ in range(args.steps):
        E = step()
    sync_all()
    cbuf = ctypes.create_string_buffer(1 << 16)
    L.bn_timing_report(cbuf, 1 << 16)
    L.bn_timing_enable(0)
    kt = {}
    for ln in cbuf.value.decode().splitlines():
        name, cnt, tot = ln.split()
        kt[name] = (int(cnt), float(tot))
    energy = float(E)
    # end to end: the observations (and step lengths) of every step come from pinned host memory, the energy goes back
    y_pin = torch.from_numpy(np.ascontiguousarray(y_h)).pin_memory()
    dt_pin = torch.from_numpy(np.ascontiguousarray(dt_h)).pin_memory() if dt_h is not None else None
    y_dev = model.Y
    h2d = y_pin.numel() * 8 + (dt_pin.numel() * 8 if dt_pin is not None else 0)
    e_pin = torch.zeros(1, dtype=torch.float64).pin_memory()

    def e2e_step():
        y_dev.reshape(-1).copy_(y_pin.reshape(-1), non_blocking=True)
        if dt_pin is not None and hasattr(model, 'dt'):
            model.dt.reshape(-1).copy_(dt_pin.reshape(-1), non_blocking=True)
        st = getattr(model, '_fused', None)
        if st is not None:  # the fused iteration keeps tiled copies of its inputs
            st.set_dt(model.dt)
            st.set_data(model.Y, getattr(model, 'mask_pseudo_y', None))
        e = step()
        e_pin.copy_(e.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(e_pin[0])

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync_all()
    ms2 = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2) / args.steps
    clocks = sampler.stop()
    scratch = torch.empty(8 * 148 * 256 * 2, dtype=torch.float64, device=dev)
    dfma, dmma = ctypes.c_double(0.0), ctypes.c_double(0.0)
    L.bn_measure_dfma_peak(scratch.data_ptr(), scratch.numel(), ctypes.byref(dfma))
    L.bn_measure_dmma_peak(scratch.data_ptr(), scratch.numel(), ctypes.byref(dmma))
    if rank == 0:
        peak, peak_src = measured_peaks()
        kb = dict(ALGO_BYTES, **roofdesc.get('kernel_bytes', {}))
        cand = [k for k in kt if roofdesc['bound'] == 'fp64_tensor' or kb.get(k, 0) > 0]
        dom = max(cand, key=lambda k: kt[k][1]) if cand else None  # the dominant kernel among those that move per-step data
        roof = None
        if dom:
            cnt, tot = kt[dom]
            avg_ms = tot / cnt
            if roofdesc['bound'] == 'fp64_tensor':
                per_launch = {'st_filter': roofdesc['flops_filter'], 'st_smoother': roofdesc['flops_smoother'],
                              'st_gain': roofdesc['flops_smoother']}.get(dom, roofdesc['flops_filter']) * NT
                achieved = per_launch / (avg_ms * 1e-3) / 1e12
                pk = 2 * dmma.value / 1e12
                roof = {'bound': 'fp64_tensor', 'kernel': dom, 'achieved': achieved, 'peak': pk, 'unit': 'TFLOP/s',
                        'frac': achieved / pk if pk else None, 'traffic': None,
                        'peak_source': 'measured now: mma.sync.m8n8k4.f64 (DMMA) microbenchmark, bn_measure_dmma_peak; '
                                       'fp64 FMA pipe %.1f TFLOP/s (bn_measure_dfma_peak)' % (2 * dfma.value / 1e12),
                        'algorithmic_flops_per_launch': per_launch, 'avg_launch_ms': avg_ms,
                        'share_of_step': tot / (ms_per_step * args.steps),
                        'whole_iteration': {'algorithmic_tflop': roofdesc['flops_per_step'] * NT / 1e12,
                                            'achieved_tflops': roofdesc['flops_per_step'] * NT / (ms_per_step * 1e-3) / 1e12,
                                            'frac_of_dmma_peak': roofdesc['flops_per_step'] * NT / (ms_per_step * 1e-3) / (2 * dmma.value) if dmma.value else None}}
            else:
                ab = kb.get(dom, 0)
                achieved = ab * NT / (avg_ms * 1e-3) / 1e9
                roof = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                        'traffic': args.traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_step': ab, 'avg_launch_ms': avg_ms,
                        'share_of_step': tot / (ms_per_step * args.steps), 'note': 'workload bound: ' + roofdesc['bound']}
        line = {'metric': METRIC, 'value': NT / (ms_per_step * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': dict(workload_config(args, world), model=type(model).__module__ + '.' + type(model).__name__,
                               parallelism='single GPU' if world == 1 else 'latent-shard x%d' % world,
                               l2='working set %s L2 (126 MB)' % ('below: launch-latency bound' if args.workload == 'C1' else 'above')),
                'clocks': clocks,
                'e2e': {'value': NT / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': int(h2d),
                        'd2h_bytes_per_step': 8},
                'gpu_launches': int(sum(c for c, _ in kt.values())),
                'gpu_launches_per_step': int(sum(c for c, _ in kt.values())) // max(1, args.steps),
                'roofline': roof, 'kernels_ms_per_step': {k: v[1] / args.steps for k, v in kt.items()}, 'energy': energy,
                'fp64_peak_dfma_per_s': dfma.value, 'fp64_peak_dmma_fma_per_s': dmma.value}
        if 'bytes_per_step' in roofdesc:
            b = roofdesc['bytes_per_step']
            line['iteration_bytes'] = {'algorithmic_bytes_per_time_step': b, 'achieved_GBps': b * NT / (ms_per_step * 1e-3) / 1e9,
                                       'frac_of_hbm_peak': b * NT / (ms_per_step * 1e-3) / 1e9 / peak}
        if graph_ms is not None:
            line['cuda_graph_replay'] = ({'ms_per_step': graph_ms, 'value': NT / (graph_ms * 1e-3), 'unit': UNIT,
                                          'note': 'the same iteration captured once with torch.cuda.graph and replayed'}
                                         if not isinstance(graph_ms, str) else {'note': graph_ms})
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='C5', choices=sorted(WORKLOADS))
    ap.add_argument('--n-total', type=int, default=None, help='time steps of the whole series (default: the workload\'s)')
    ap.add_argument('--cpu-sample', type=int, default=2_000_000, help='time steps of the bounded CPU sample')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-grad', action='store_true', help='skip the with-hyper-gradient leg')
    ap.add_argument('--no-fp32', action='store_true', help='skip the fp32-mode leg')
    ap.add_argument('--traffic', type=float, default=None, help='dram bytes per launch of the dominant kernel (ncu)')
    args = ap.parse_args()
    if args.n_total is None:
        args.n_total = WORKLOADS[args.workload]['n_total']
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        main_reference(args)
    elif args.workload in ('C2', 'C5'):
        main_gpu(args)
    else:
        if args.workload == 'C4' and args.steps > 3:
            args.steps = 2   # one C4 iteration takes seconds
        main_generic(args)


if __name__ == '__main__':
    main()

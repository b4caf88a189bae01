"""Latent-sharded Markov GP (config C3: Independent[Matern32 x2] + HeteroscedasticNoise, demos/heteroscedastic.py:49-56).

CPU: two real processes (gloo, 127.0.0.1) run the PRODUCT's host logic (bayesnewton_b200.latent_sharding: per-latent
posterior updates, the two all-to-all layout exchanges, the energy all-reduce) with the per-rank array work done by
the host emulation of the same kernel bodies; checked against the oracle's JOINT model (d = 4, D = 2).
GPU: the same class on libbn_b200 -- two rank objects in threads on one device (barrier-based exchange), and two
NCCL processes when two GPUs are visible -- against the joint single-GPU model and the oracle.
Tolerance 1e-8 as for the joint heteroscedastic model (400-point cubature sums), observed ~1e-13."""
import ctypes as C
import os
import socket
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METHOD_NAMES = {1: 'vi', 2: 'ep', 3: 'newton'}


def c3_data(N, seed=3):
    rng = np.random.default_rng(seed)
    x = np.sort(0.1 * N * rng.random(N))
    y = np.sin(x) + 0.3 * (1 + np.cos(x)) * rng.standard_normal(N)
    return x, (y - y.mean()) / y.std()


class EmuBackend:
    """LibBackend's interface on CPU tensors, executed by tests/hostemu (test infrastructure)"""

    def __init__(self):
        import _emu
        self.emu, self.lib = _emu, _emu.load()

    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.float64)

    def to_dev(self, x):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).contiguous()

    def update_posterior(self, spec, dt, y, R, want_grad=False):
        sp = self.emu.spec(spec.family, list(spec.variance)[:spec.n_components], list(spec.lengthscale)[:spec.n_components])
        out = self.emu.update_posterior(self.lib, sp, dt.numpy(), y.numpy(), R.numpy(), L=8, want_grad=want_grad)
        g = torch.from_numpy(np.stack([out[3], out[4]])) if want_grad else None
        return torch.tensor(out[0]), torch.from_numpy(out[1]), torch.from_numpy(out[2]), g

    def _cub(self, lik, method, cubature):
        from bayesnewton_b200.cubature import host_table
        if method == 3:
            return None
        cx, cw, _ = host_table(cubature, lik.num_latents)
        return cx, cw

    def site_update(self, lik, method, power, Y, pm, pc, nat1, nat2, site_mean, site_cov, lr, cubature):
        o = self.emu.site_update(self.lib, METHOD_NAMES[method], 'het_softplus', 0.0, Y.numpy(), pm.numpy(), pc.numpy(),
                                 nat1.numpy(), nat2.numpy(), lr=lr, power=power, ensure_psd=True,
                                 cub=self._cub(lik, method, cubature))
        for dst, key in ((nat1, 'nat1'), (nat2, 'nat2'), (site_mean, 'site_mean'), (site_cov, 'site_cov')):
            dst.copy_(torch.from_numpy(o[key]))
        return torch.from_numpy(o['diffs'])

    def expected_density(self, lik, method, power, Y, pm, pc, nat1, nat2, cubature):
        _, s = self.emu.expected_density(self.lib, METHOD_NAMES[method], 'het_softplus', 0.0, Y.numpy(), pm.numpy(),
                                         pc.numpy(), nat1.numpy(), nat2.numpy(), power=power,
                                         cub=self._cub(lik, method, cubature))
        return torch.tensor(s)

    def gaussian_expected_log_lik(self, py, pm, pc, pv):
        s = np.zeros(1)
        a = [np.ascontiguousarray(t.numpy()) for t in (py, pm, pc, pv)]
        self.lib.emu_gaussian_expected_log_lik(C.c_longlong(py.shape[0]), py.shape[1], *[x.ctypes.data_as(C.c_void_p) for x in a],
                                               None, s.ctypes.data_as(C.c_void_p))
        return torch.tensor(s[0])

    def ep_pseudo_density(self, power, py, pv, pm, pc, nat1, nat2):
        s = np.zeros(1)
        a = [np.ascontiguousarray(t.numpy()) for t in (py, pv, pm, pc, nat1, nat2)]
        self.lib.emu_ep_pseudo_density(C.c_longlong(py.shape[0]), py.shape[1], C.c_double(power), 1,
                                       *[x.ctypes.data_as(C.c_void_p) for x in a], None, s.ctypes.data_as(C.c_void_p))
        return torch.tensor(s[0])


class HostLik:
    """the two attributes of likelihoods.HeteroscedasticNoise the host logic reads (importing the real class is fine
    on CPU too; this keeps the CPU worker free of any device query)"""
    num_latents = 2


def _oracle(method, x, y, iters, lr):
    from oracle import model, sites, ssm
    o = model.MarkovGP(ssm.Independent([ssm.Matern32(1.0, 1.0), ssm.Matern32(0.7, 2.0)]), sites.HeteroscedasticNoise(), x, y,
                       method=method, power=0.5)
    out = []
    for _ in range(iters):
        o.inference(lr=lr)
        out.append((o.post_mean.copy(), o.post_cov.copy(), o.energy()))
    return o, out


def _gloo_worker(rank, world, port, method_id, N, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from bayesnewton_b200 import kernels as K, latent_sharding as ls
    x, y = c3_data(N)
    m = ls.LatentShardedMarkovGP(K.Independent([K.Matern32(1.0, 1.0), K.Matern32(0.7, 2.0)]), HostLik(), x, y, method_id,
                                 rank, world, power=0.5, backend=EmuBackend(), comm=ls.TorchDistComm(world))
    res = []
    for it in range(2):
        m.inference(lr=0.3, want_grad=(it == 1))
        E = float(m.energy())
        pm, pc = m.posterior_time_slice()
        res.append((pm.numpy().copy(), pc.numpy().copy(), E))
    E, g = m.energy_and_grad()
    out[rank] = (res, g.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('method_id', [1, 2, 3])
def test_two_rank_gloo_latent_sharded(method_id):
    import torch.multiprocessing as mp
    from _data import rel_err
    from oracle import grad
    N, world = 37, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(world, port, method_id, N, out), nprocs=world, join=True)
    x, y = c3_data(N)
    o, ref = _oracle(METHOD_NAMES[method_id], x, y, 2, 0.3)
    for it in range(2):
        pm = np.concatenate([out[r][0][it][0] for r in range(world)])
        pc = np.concatenate([out[r][0][it][1] for r in range(world)])
        assert rel_err(pm, ref[it][0]) < 1e-8 and rel_err(pc, ref[it][1]) < 1e-8
        for r in range(world):
            assert abs(out[r][0][it][2] - ref[it][2]) <= 1e-8 * abs(ref[it][2])
    # hyper-gradient of the energy w.r.t. [variance_c; lengthscale_c]: -d ell / d theta of the JOINT filter
    _, g0 = grad.ell_grad_adjoint(o.kernel, o.dt, o.site_mean, o.site_cov)
    for r in range(world):
        g = out[r][1]
        assert rel_err(np.stack([g[0], g[1]], 1).reshape(-1), -g0) < 1e-8


# ------------------------------------------------------------------------------------------ GPU
class ThreadComm:
    """collectives between rank objects living in threads of ONE process (single-GPU validation of the exchanges)"""

    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.bar = threading.Barrier(world)
        self.local = threading.local()

    def bind(self, rank):
        self.local.rank = rank

    def _exchange(self, value):
        r = self.local.rank
        torch.cuda.current_stream().synchronize()
        self.slots[r] = value
        self.bar.wait()
        got = list(self.slots)
        self.bar.wait()
        return got

    def all_to_all(self, send, out_splits, in_splits):
        r = self.local.rank
        pieces = self._exchange(torch.split(send, list(in_splits)))
        return torch.cat([pieces[q][r] for q in range(self.world)])

    def all_reduce(self, t):
        vals = self._exchange(t.clone())
        t.copy_(sum(vals[1:], vals[0].clone()))
        return t

    def all_gather(self, t):
        return torch.stack(self._exchange(t.clone()))


@pytest.fixture(scope='module')
def bn():
    assert torch.cuda.is_available(), 'the -m gpu tests need a CUDA device'
    import bayesnewton_b200 as bn
    return bn


def _run_threads(bn, method_id, x, y, world, iters, lr):
    from bayesnewton_b200 import latent_sharding as ls
    K = bn.kernels
    comm = ThreadComm(world) if world > 1 else None
    results, errors = [None] * world, []

    def work(rank):
        try:
            torch.cuda.set_device(0)
            if comm is not None:
                comm.bind(rank)
            m = ls.LatentShardedMarkovGP(K.Independent([K.Matern32(1.0, 1.0), K.Matern32(0.7, 2.0)]),
                                         bn.likelihoods.HeteroscedasticNoise(), x, y, method_id, rank, world, power=0.5,
                                         comm=comm)
            res = []
            for it in range(iters):
                m.inference(lr=lr, want_grad=(it == iters - 1))
                E = float(m.energy())
                pm, pc = m.posterior_time_slice()
                res.append((pm.cpu().numpy(), pc.cpu().numpy(), E))
            _, g = m.energy_and_grad()
            results[rank] = (res, g.cpu().numpy())
        except Exception as e:  # surface worker failures in the main thread
            errors.append(e)
            if comm is not None:
                comm.bar.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.gpu
@pytest.mark.parametrize('method_id', [1, 2, 3])
@pytest.mark.parametrize('world', [1, 2])
def test_gpu_latent_sharded_vs_oracle(bn, method_id, world):
    from _data import rel_err
    from oracle import grad
    N = 150
    x, y = c3_data(N)
    out = _run_threads(bn, method_id, x, y, world, 2, 0.3)
    o, ref = _oracle(METHOD_NAMES[method_id], x, y, 2, 0.3)
    for it in range(2):
        pm = np.concatenate([out[r][0][it][0] for r in range(world)])
        pc = np.concatenate([out[r][0][it][1] for r in range(world)])
        assert rel_err(pm, ref[it][0]) < 1e-8 and rel_err(pc, ref[it][1]) < 1e-8
        assert abs(out[0][0][it][2] - ref[it][2]) <= 1e-8 * abs(ref[it][2])
    _, g0 = grad.ell_grad_adjoint(o.kernel, o.dt, o.site_mean, o.site_cov)
    g = out[0][1]
    assert rel_err(np.stack([g[0], g[1]], 1).reshape(-1), -g0) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize('method_id', [1, 3])
def test_gpu_latent_sharded_matches_joint_model_c3_size(bn, method_id):
    """N = 10^5 (a tenth of C3; the 400-point VI cubature dominates): two latent shards == the joint d = 4 model"""
    from _data import rel_err
    N = 100_000
    x, y = c3_data(N)
    out = _run_threads(bn, method_id, x, y, 2, 1, 0.3)
    K = bn.kernels
    cls = {1: bn.models.MarkovVariationalGP, 3: bn.models.MarkovNewtonGP}[method_id]
    g = cls(kernel=K.Independent([K.Matern32(1.0, 1.0), K.Matern32(0.7, 2.0)]),
            likelihood=bn.likelihoods.HeteroscedasticNoise(), X=x, Y=y, parallel=True)
    g.inference(lr=0.3)
    pm = np.concatenate([out[r][0][0][0] for r in range(2)])
    pc = np.concatenate([out[r][0][0][1] for r in range(2)])
    assert rel_err(pm, g.posterior_mean.cpu().numpy()) < 1e-8
    assert rel_err(pc, g.posterior_variance.cpu().numpy()) < 1e-8
    E = float(g.energy())
    assert abs(out[0][0][0][2] - E) <= 1e-8 * abs(E)


def _nccl_worker(rank, world, port, method_id, N, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import bayesnewton_b200 as bn
    from bayesnewton_b200 import latent_sharding as ls
    K = bn.kernels
    x, y = c3_data(N)
    m = ls.LatentShardedMarkovGP(K.Independent([K.Matern32(1.0, 1.0), K.Matern32(0.7, 2.0)]),
                                 bn.likelihoods.HeteroscedasticNoise(), x, y, method_id, rank, world, power=0.5)
    m.inference(lr=0.3)
    E = float(m.energy())
    pm, pc = m.posterior_time_slice()
    out[rank] = (pm.cpu().numpy(), pc.cpu().numpy(), E)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_gpu_two_process_nccl_latent_sharded(bn):
    """the real thing: one process per GPU, NCCL all-to-all over NVLink (needs two visible GPUs)"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    from _data import rel_err
    N, world = 20_000, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(world, port, 1, N, out), nprocs=world, join=True)
    x, y = c3_data(N)
    ref = _run_threads(bn, 1, x, y, 1, 1, 0.3)
    pm = np.concatenate([out[r][0] for r in range(world)])
    pc = np.concatenate([out[r][1] for r in range(world)])
    assert rel_err(pm, ref[0][0][0][0]) < 1e-9 and rel_err(pc, ref[0][0][0][1]) < 1e-9
    assert abs(out[0][2] - ref[0][0][0][2]) <= 1e-9 * abs(ref[0][0][0][2])

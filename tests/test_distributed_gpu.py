"""N > 1 path on GPUs: one process per GPU (torch.distributed, NCCL), the time-sharded model with the carries exchanged
(a) over NVLink peer memory by the library's own kernel (bn_carry_exchange) and (b) by NCCL all-gather; both against
the single-GPU model.  Needs two visible GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, mode, fused, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    os.environ['BN_B200_CARRY_EXCHANGE'] = mode
    os.environ['BN_B200_FUSED'] = '1' if fused else '0'
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import bayesnewton_b200 as bn
    from bayesnewton_b200 import _lib, distributed
    from _data import bench_inputs
    t, dt, y = bench_inputs(N)
    b = distributed.shard_bounds(N, world)
    lo, hi = b[rank], b[rank + 1]
    dt_next = float(dt[hi]) if hi < N else 0.0
    m = distributed.TimeShardedMarkovGP(bn.kernels.Matern52(1.0, 1.0), bn.likelihoods.Bernoulli(), torch.from_numpy(dt[lo:hi].copy()),
                                        torch.from_numpy(y[lo:hi].copy()), dt_next, _lib.BN_METHOD_VI, rank, world)
    used_p2p = distributed.peer_exchange(world) is not None
    Es = []
    for _ in range(3):
        m.inference(lr=0.7)
        Es.append(float(m.energy()))
    E, g = m.energy_and_grad()
    out[rank] = (m.posterior_mean.cpu().numpy(), m.posterior_variance.cpu().numpy(), Es, float(E), g.cpu().numpy(), used_p2p)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4, 8])
@pytest.mark.parametrize('mode', ['p2p', 'nccl'])
@pytest.mark.parametrize('fused', [True, False])
def test_gpu_multi_process_time_sharded(world, mode, fused):
    """`world` ranks (one process per GPU), the fused tiled iteration and the unfused entry points, against one GPU"""
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs (run with gpurun --gpus %d)' % (world, world))
    import torch.multiprocessing as mp
    import bayesnewton_b200 as bn
    from _data import bench_inputs, rel_err
    N = 200_003
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, N, mode, fused, out), nprocs=world, join=True)
    t, dt, y = bench_inputs(N)
    ref = bn.models.MarkovVariationalGP(kernel=bn.kernels.Matern52(1.0, 1.0), likelihood=bn.likelihoods.Bernoulli(), X=t, Y=y, parallel=True)
    Es = []
    for _ in range(3):
        ref.inference(lr=0.7)
        Es.append(float(ref.energy()))
    pm = np.concatenate([out[r][0] for r in range(world)])
    pc = np.concatenate([out[r][1] for r in range(world)])
    assert rel_err(pm, ref.posterior_mean.cpu().numpy()) < 1e-9 and rel_err(pc, ref.posterior_variance.cpu().numpy()) < 1e-9
    for r in range(world):
        assert np.allclose(out[r][2], Es, rtol=1e-9, atol=0)
        assert out[r][5] == (mode == 'p2p'), 'the peer-memory exchange was expected to be %s' % ('on' if mode == 'p2p' else 'off')
    assert all(out[r][2] == out[0][2] for r in range(world))  # every rank holds the same energy, bit for bit

"""N > 1 path on CPU: two real processes (torch.distributed, gloo, 127.0.0.1) run the PRODUCT's host logic
(bayesnewton_b200.distributed.sharded_update_posterior: reduce -> all-gather -> filter -> all-gather -> smooth,
and the energy all-reduce pattern) with each rank's kernels replaced by their host emulation (tests/hostemu,
the same __host__ __device__ chunk bodies).  Checked against the oracle run on the whole series."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class EmuShard:
    """duck-types distributed.TimeShard (up_reduce / up_filter / up_smooth / world) on CPU tensors"""

    def __init__(self, emu, spec, dt, rank, world, L=8):
        self.emu, self.rank, self.world = emu, rank, world
        self.dt = np.ascontiguousarray(dt, dtype=np.float64)
        self.N, self.D = self.dt.shape[0], spec.n_components
        emu.emu_rank_new.restype = C.c_void_p
        emu.emu_rank_new.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int]
        for f in (emu.emu_rank_free, emu.emu_rank_reduce, emu.emu_rank_filter, emu.emu_rank_smooth):
            f.restype = None
        emu.emu_rank_free.argtypes = [C.c_void_p]
        emu.emu_rank_kf_len.argtypes = emu.emu_rank_rts_len.argtypes = [C.c_void_p]
        emu.emu_rank_reduce.argtypes = [C.c_void_p] * 5
        emu.emu_rank_filter.argtypes = [C.c_void_p] * 5
        emu.emu_rank_smooth.argtypes = [C.c_void_p] * 6
        self.h = emu.emu_rank_new(C.addressof(spec), self.N, L, rank, world)
        assert self.h
        self.kf_len, self.rts_len = emu.emu_rank_kf_len(self.h), emu.emu_rank_rts_len(self.h)
        self.keep = []

    @staticmethod
    def _np(t):
        return np.ascontiguousarray(t.numpy() if torch.is_tensor(t) else t, dtype=np.float64)

    def up_reduce(self, y, R, want_grad=False):
        y, R = self._np(y), self._np(R)
        self.keep = [y, R]
        carry = np.zeros(self.kf_len)
        self.emu.emu_rank_reduce(self.h, self.dt.ctypes.data, y.ctypes.data, R.ctypes.data, carry.ctypes.data)
        return torch.from_numpy(carry)

    def up_filter(self, kf_carries, y, R, mask=None, want_ell=True, want_grad=False):
        kc = self._np(kf_carries)
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        ell, carry = np.zeros(1), np.zeros(self.rts_len)
        self.emu.emu_rank_filter(self.h, kc.ctypes.data, None if mk is None else mk.ctypes.data, ell.ctypes.data,
                                 carry.ctypes.data)
        return torch.from_numpy(ell)[0], torch.from_numpy(carry)

    def up_smooth(self, rts_carries, want_grad=False):
        rc = self._np(rts_carries)
        pm, pc = np.zeros((self.N, self.D, 1)), np.zeros((self.N, self.D, self.D))
        g = np.zeros((2, self.D))
        self.emu.emu_rank_smooth(self.h, rc.ctypes.data, pm.ctypes.data, pc.ctypes.data,
                                 g[0].ctypes.data if want_grad else None, g[1].ctypes.data if want_grad else None)
        if want_grad:
            return torch.from_numpy(pm), torch.from_numpy(pc), torch.from_numpy(g)
        return torch.from_numpy(pm), torch.from_numpy(pc)


def _worker(rank, world, port, fam, vs, ls, N, seed, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import _emu
    from _data import filter_problem
    from bayesnewton_b200 import distributed
    emu = _emu.load()
    D = len(vs)
    dt, y, R, mask = filter_problem(N, D=D, seed=seed)
    b = distributed.shard_bounds(N, world)
    lo, hi = b[rank], b[rank + 1]
    sp = _emu.spec(fam, vs, ls)
    shard = EmuShard(emu, sp, dt[lo:hi], rank, world)
    ell, sm, sP = distributed.sharded_update_posterior(shard, y[lo:hi], R[lo:hi], mask[lo:hi], want_ell=True)
    tot = ell.clone().reshape(1)
    dist.all_reduce(tot)  # the scalar all-reduce of TimeShardedMarkovGP.energy
    # the hyper-gradient pass of the same update (no mask): local shares, one all-reduce
    _, sm2, sP2, g = distributed.sharded_update_posterior(shard, y[lo:hi], R[lo:hi], None, want_ell=True, want_grad=True)
    g = g.clone()
    dist.all_reduce(g)
    out[rank] = (float(tot[0]), sm.numpy().copy(), sP.numpy().copy(), g.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('fam,vs,ls', [(3, [1.3], [0.9]), (2, [1.0, 0.5], [1.0, 2.0])])
def test_two_rank_gloo_sharded_update(fam, vs, ls):
    from _data import filter_problem, rel_err
    from oracle import kalman, ssm
    N, seed, world = 61, 5, 2
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, fam, vs, ls, N, seed, out), nprocs=world, join=True)
    dt, y, R, mask = filter_problem(N, D=len(vs), seed=seed)
    mk = {3: ssm.Matern52, 2: ssm.Matern32}[fam]
    k = mk(vs[0], ls[0]) if len(vs) == 1 else ssm.Independent([mk(v, l) for v, l in zip(vs, ls)])
    e0, (fm, fP) = kalman.kalman_filter(dt, k, y, R, mask)
    sm, sP, _ = kalman.rauch_tung_striebel_smoother(np.concatenate([dt[1:], [0.0]]), k, fm, fP)
    assert abs(out[0][0] - e0) < 1e-9 * abs(e0) and abs(out[1][0] - e0) < 1e-9 * abs(e0)
    pm = np.concatenate([out[0][1], out[1][1]])
    pc = np.concatenate([out[0][2], out[1][2]])
    assert rel_err(pm, sm) < 1e-9 and rel_err(pc, sP) < 1e-9
    from oracle import grad
    _, g0 = grad.ell_grad_adjoint(k, dt, y, R)
    for r in range(world):
        g = out[r][3]  # [2, NC] -> the oracle's (variance_c, lengthscale_c) pairs
        assert rel_err(np.stack([g[0], g[1]], 1).reshape(-1), g0) < 1e-9

"""Device-memory plumbing shared by the host-side modules (torch is used for buffers and streams only)."""
import numpy as np
import torch

from . import _lib

_workspaces = {}


def device():
    if not torch.cuda.is_available():
        raise _lib.BnError('bayesnewton_b200 needs a CUDA device: there is no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def as_dev(x, dtype=torch.float64):
    """contiguous device tensor of the given dtype from a numpy array / tensor / scalar"""
    if torch.is_tensor(x):
        t = x.to(device=device(), dtype=dtype)
    else:
        t = torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device=device())
    return t.contiguous()


def as_mask(mask):
    """[N,D,1] bool -> contiguous uint8 device tensor (None stays None)"""
    if mask is None:
        return None
    if torch.is_tensor(mask):
        return mask.to(device=device(), dtype=torch.uint8).contiguous()
    return torch.as_tensor(np.ascontiguousarray(mask).astype(np.uint8), device=device()).contiguous()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def workspace(N, d, D):
    """cached per-device scratch buffer, grown on demand; (tensor, nbytes)"""
    need = _lib.lib().bn_workspace_bytes(int(N), int(d), int(D))
    key = torch.cuda.current_device()
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(int(need), dtype=torch.uint8, device=device())
        _workspaces[key] = ws
    return ws, ws.numel()


_up_workspaces = {}


def up_workspace(spec, N):
    """cached per-device scratch of the fused posterior update (holds the filtered states: ~72 B/step at d = 3)"""
    need = _lib.lib().bn_update_posterior_workspace_bytes(spec, int(N))
    key = torch.cuda.current_device()
    ws = _up_workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = None
        _up_workspaces.pop(key, None)
        ws = torch.empty(int(need), dtype=torch.uint8, device=device())
        _up_workspaces[key] = ws
    return ws, ws.numel()

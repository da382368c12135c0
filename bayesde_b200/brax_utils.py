"""Config 1 (row a20): the fixed-step solver of the JAX toy, brax/utils/sdeint.py:8-60, on the CUDA step kernel.

    states = sdeint(f, g, s_init, bm_dim, ts, rng)        # all states after every step  (:30-60)
    final  = _sdeint(f, g, s_init, bm_dim, ts, rng)       # last state only               (:8-28)

Same call shape as the reference: `f(t, state, bm_w)` (quirk Q5: the drift receives the unit normals of the
first `bm_dim` state entries, brax/_impl/layers.py:76-77,111-113), `g(t, state)` the diagonal diffusion; per step
    bm ~ N(0, I) over the whole state;  state <- state + h f(t0, state, bm[:bm_dim]) + sqrt(h) bm g(t0, state).
`f` / `g` are torch callables on CUDA tensors (the toy's SDELayer stack is user model code); the state update is
the fused elementwise kernel bsde_step_update and the normals come from the in-kernel Philox stream
(bsde_philox_increments with unit variance) or, for exact comparison, from a tensor `rng` of shape
(len(ts)-1, D).  There is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from .sdeint import _StepUpdate, philox_increments


def _normals(rng, nsteps, D, device):
    if isinstance(rng, torch.Tensor):
        if tuple(rng.shape) != (nsteps, D):
            raise ValueError(f"unit normals must have shape {(nsteps, D)}, got {tuple(rng.shape)}")
        return rng.to(device, torch.float32).contiguous()
    seed, sid = (rng if isinstance(rng, (tuple, list)) else (rng, 0))
    # unit-variance draws: N(0, dt_k) with dt_k = 1
    return philox_increments(int(seed), int(sid), 1, np.ones(nsteps, dtype=np.float32), D, device)[0]


def _scan(f, g, s_init, bm_dim, ts, rng, keep):
    if not s_init.is_cuda:
        raise _lib.BsdeError("sdeint needs CUDA tensors (there is no CPU fallback)")
    ts = torch.as_tensor(ts, dtype=torch.float32)
    D = s_init.numel()
    bm = _normals(rng, len(ts) - 1, D, s_init.device)
    state = s_init.to(torch.float32).contiguous()
    states = []
    for k in range(len(ts) - 1):
        t0 = ts[k].to(state.device)
        h = float(ts[k + 1] - ts[k])
        f_out = f(t0, state, bm[k, :bm_dim]).contiguous()
        g_out = g(t0, state).contiguous()
        dW = (bm[k] * float(np.sqrt(np.float32(h)))).contiguous()
        state = _StepUpdate.apply(_lib.METHOD["euler_maruyama"], h, state, f_out, g_out, dW, None, None)
        if keep:
            states.append(state)
    return torch.stack(states) if keep else state


def sdeint(f, g, s_init, bm_dim, ts, rng):
    """brax/utils/sdeint.py:30-60: returns the (len(ts)-1, D) states after every step."""
    return _scan(f, g, s_init, bm_dim, ts, rng, True)


def _sdeint(f, g, s_init, bm_dim, ts, rng):
    """brax/utils/sdeint.py:8-28: returns the final state."""
    return _scan(f, g, s_init, bm_dim, ts, rng, False)

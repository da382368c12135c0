"""Generate the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

jax_block_*.npz : seeded inputs + fp64 oracle outputs (x_T, per-sample KL, w trajectory, gradients) of one
                  SDEBNN block (oracle/jax_path.py restating brax/_impl/brax.py + jaxsde/sdeint.py).
The oracle itself is pinned by tests/test_oracle.py; fixtures from reference code that runs in this
container (torchbnn conv stack, torch toy) are produced by make_golden_torch.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import jax_path as jp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "jax_block_a": dict(S=2, B=2, H=8, C_in=5, fx_dim=8, nsteps=6, sigma=0.2, stl=False, fw_dims=(2, 16, 2),
                        time_grid="reference", seed=11),
    "jax_block_b": dict(S=1, B=3, H=4, C_in=20, fx_dim=16, nsteps=5, sigma=0.05, stl=False, fw_dims=(1, 8, 1),
                        time_grid="uniform", seed=12),
}


def make(name, S, B, H, C_in, fx_dim, nsteps, sigma, stl, fw_dims, time_grid, seed):
    g = torch.Generator().manual_seed(seed)
    blk = jp.SDEBNNBlock((B, H, H, C_in), 0, fx_dim, "softplus", "softplus", sigma, stl, nsteps,
                         time_grid_mode=time_grid)
    w0 = jp.init_fx(blk.layout, blk.P, g, torch.float64).requires_grad_(True)
    fw = [tuple(t.requires_grad_(True) for t in lay) for lay in jp.init_fw(blk.P, fw_dims, g, torch.float64, 0.05)]
    x = torch.randn(S, B, H, H, C_in, generator=g, dtype=torch.float64).requires_grad_(True)
    dW = torch.stack([jp.sample_increments(blk.P, nsteps, g, torch.float64, time_grid) for _ in range(S)])
    cot = torch.randn(S, B, H, H, C_in, generator=g, dtype=torch.float64)
    xs, kls, ws = [], [], []
    for s in range(S):
        xo, kl, wt = blk.apply((w0, fw), x[s], dW[s])
        xs.append(xo), kls.append(kl), ws.append(wt)
    xo, kl, wt = torch.stack(xs), torch.stack(kls), torch.stack(ws)
    leaves = [x, w0] + [t for lay in fw for t in lay]
    grads = torch.autograd.grad((xo * cot).sum() + 1e-3 * kl.sum(), leaves)
    out = dict(meta=np.array([S, B, H, C_in, fx_dim, nsteps, len(fw_dims)] + list(fw_dims)), sigma=sigma,
               time_grid=time_grid, x=x.detach().numpy(), w0=w0.detach().numpy(), dW=dW.numpy(), cot=cot.numpy(),
               xT=xo.detach().numpy(), kl=kl.detach().numpy(), wtraj=wt.detach().numpy())
    for i, lay in enumerate(fw):
        for n, t in zip(("W", "b", "wt", "wtb", "bt"), lay):
            out[f"fw{i}_{n}"] = t.detach().numpy()
    for i, gr in enumerate(grads):
        out[f"grad{i}"] = gr.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", v) for k, v in out.items() if k in ("x", "w0", "xT", "kl")})


if __name__ == "__main__":
    for n, kw in CASES.items():
        make(n, **kw)

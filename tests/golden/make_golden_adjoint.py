"""Golden fixture for SURVEY 8.6 row N1 (run from the repo root: python tests/golden/make_golden_adjoint.py).

jax_block_adjoint.npz : seeded inputs + fp64 oracle outputs of one SDEBNN block solved with `sdeint_ito` (variable-step loop on the
                        Philox Brownian tree, oracle/jax_adjoint.py) and differentiated with the reference's stochastic adjoint
                        (`_sdeint_ito_rev`, jaxsde/jaxsde/sde_vjp.py:180-192): x_T, per-sample KL, w at the output times, the
                        reconstructed y0 and every gradient.  The oracle itself is pinned by tests/test_adjoint.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_adjoint import _adjoint_case, _oracle_block_adjoint  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED, SID, DT, KLW = 17, 2, 0.06, 1e-3


def main():
    blk, w0, fw, x, cot = _adjoint_case(False, S=2, B=2, H=8, C_in=5, fx_dim=8, nsteps=4, fw_dims=(2, 8, 2), seed=4)
    blk.w_stale = torch.zeros(blk.P, dtype=torch.float64)          # stl = False: g_aug never reads it
    res, unravel = _oracle_block_adjoint(blk, w0, fw, x, SEED, SID, DT, cot, KLW)
    P, x_dim = blk.P, blk.x_dim
    out = dict(meta=np.array([x.shape[0], x.shape[1], x.shape[2], x.shape[4], 8, blk.nsteps, SEED, SID]), dt=DT, kl_w=KLW, sigma=blk.diff_coef,
               fw_dims=np.array([2, 8, 2]), x=x.numpy(), w0=w0.numpy(), cot=cot.numpy(),
               xT=np.stack([r[0][-1][:x_dim].numpy() for r in res]), kl=np.array([float(r[0][-1][x_dim + P:].sum()) for r in res]),
               wrows=np.stack([r[0][:, x_dim:x_dim + P].numpy() for r in res]),
               x0_rec=np.stack([r[1][:x_dim].numpy() for r in res]), w0_rec=np.stack([r[1][x_dim:x_dim + P].numpy() for r in res]),
               grad_x=np.stack([r[2][:x_dim].numpy() for r in res]), grad_w0=sum(r[2][x_dim:x_dim + P] for r in res).numpy())
    for i, lay in enumerate(fw):
        for n, t in zip(("W", "b", "wt", "wtb", "bt"), lay):
            out[f"fw{i}_{n}"] = t.numpy()
    for i, t in enumerate(t for lay in unravel(sum(r[3] for r in res)) for t in lay):
        out[f"grad_fw{i}"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "jax_block_adjoint.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items() if k in ("x", "xT", "kl", "wrows", "grad_w0")})


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of config 2, the torch toy of
examples/torch/sdebnn_toy1d.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Pinned: tests/test_oracle.py checks this restatement against tests/golden/torch_toy.npz, which was produced by
running the reference's own em_solver / construct_odenet / loglik_fn in this container
(tests/golden/make_golden_torch_toy.py).

Restated lines: em_solver sdebnn_toy1d.py:19-32 (prior drift at the NEW state, kl += dt*|u*u|/2);
construct_odenet :35-41; ConcatLinear torchbnn/_impl/diffeq_layers.py:178-188 (time channel first);
TimeDependentSwish torchbnn/_impl/basic.py:298-312; objective :112-126.
"""
import math

import torch
import torch.nn.functional as F


def beta(p, idx, t):
    """TimeDependentSwish gate: Linear(1,m) -> Softplus -> Linear(m,dim) -> Softplus at times t [K]."""
    h = F.softplus(t[:, None] @ p[f"layers.{idx}.beta.0.weight"].T + p[f"layers.{idx}.beta.0.bias"])
    return F.softplus(h @ p[f"layers.{idx}.beta.2.weight"].T + p[f"layers.{idx}.beta.2.bias"])


def drift(p, t, z):
    """f(t, z) for z [n,1], scalar t."""
    tt = torch.ones_like(z) * t
    a1 = torch.cat([tt, z], 1) @ p["layers.0.weight"].T + p["layers.0.bias"]
    h1 = a1 * torch.sigmoid(a1 * beta(p, 1, t.reshape(1)))
    a2 = torch.cat([tt, h1], 1) @ p["layers.2.weight"].T + p["layers.2.bias"]
    h2 = a2 * torch.sigmoid(a2 * beta(p, 3, t.reshape(1)))
    return torch.cat([tt, h2], 1) @ p["layers.4.weight"].T + p["layers.4.bias"]


def em_solver(p, z0, t, dW, sigma=math.sqrt(2.0), theta=0.1):
    z, zs = z0, [z0]
    kl = torch.zeros(1, dtype=z0.dtype)
    for k in range(len(t) - 1):
        dt = t[k + 1] - t[k]
        f = drift(p, t[k], z)
        z = z + f * dt + sigma * dW[k]
        zs.append(z)
        u = (f - (-theta * z)) / sigma
        kl = kl + dt * torch.abs(u * u) * 0.5
    return torch.stack(zs), kl


def objective(zT, kl, T=3.0, data=(5.0, -5.0), kl_scale=0.2):
    ll = sum(torch.distributions.Cauchy(zT, torch.ones((), dtype=zT.dtype)).log_prob(torch.tensor(x, dtype=zT.dtype)) for x in data)
    e = ll - kl_scale / T * kl
    return (torch.logsumexp(e, 0) - math.log(zT.shape[0])).mean()

"""Model assembly and training step of examples/jax/sdebnn_classification.py (--model sdenet) on the
fused B200 path: Augment -> [SDEBNN x nb -> SqueezeDownsample]... -> MeanField(Flatten, Dense, LogSoftmax),
loss = nll + kl * kl_coef / train_size, mean of per-sample gradients, Adam.

The S Monte-Carlo samples of `vmap(grad(loss))` (script :218) are integrated together: every image is
expanded to S copies, sample s uses its own weight path w_s(t) (Philox sample counter s).
"""
import torch
import torch.distributed as dist
from torch import nn

from . import arch, brax


def _leaves_with_paths(tree, path=()):
    if isinstance(tree, torch.Tensor):
        return [(path, tree)]
    out = []
    for i, t in enumerate(tree):
        out.extend(_leaves_with_paths(t, path + (i,)))
    return out


def _rebuild(tree, leaves_iter):
    if isinstance(tree, torch.Tensor):
        return next(leaves_iter)
    return type(tree)(_rebuild(t, leaves_iter) for t in tree)


class SDEBNNClassifier(nn.Module):
    def __init__(self, input_size=(32, 32, 3), aug=2, nblocks=(2, 2, 2), fx_dim=64, fw_dims=(2, 64, 2),
                 fx_actfn="softplus", fw_actfn="softplus", diff_coef=1e-4, stl=False, nsteps=20, block_type=0,
                 num_classes=10, nsamples=1, ou_dw=True, w_init=-1.0, b_init=-1.0, time_grid="reference",
                 math="fp32", method="euler_maruyama", remat=False, seed=0, device="cuda"):
        super().__init__()
        self.nsamples, self.input_size = nsamples, tuple(input_size)
        mf = brax.MeanField
        layers, kinds = [mf(arch.Augment(aug))], ["free"]
        nb_i = 0
        for i, nb in enumerate(nblocks):
            fw = arch.MLP(fw_dims, actfn=fw_actfn, xt=True, ou_dw=ou_dw, nonzero_w=w_init, nonzero_b=b_init)
            for _ in range(nb):
                layers.append(brax.SDEBNN(block_type, fx_dim, fx_actfn, fw, diff_coef=diff_coef, stl=stl, xt=True,
                                          nsteps=nsteps, remat=remat, method=method, time_grid=time_grid, math=math,
                                          stream_id=nb_i))
                kinds.append("sde")
                nb_i += 1
            if i < len(nblocks) - 1:
                layers.append(mf(arch.SqueezeDownsample(2)))
                kinds.append("free")
        layers.append(mf(arch.serial(arch.Flatten, arch.Dense(num_classes), arch.LogSoftmax)))
        kinds.append("head")
        self.kinds = kinds
        init_random_params, self._predict = brax.bnn_serial(*layers)  # script :213
        _, params = init_random_params(seed, (-1,) + self.input_size)
        self._tree = params
        leaves = [t for _, t in _leaves_with_paths(params)]
        self.leaves = nn.ParameterList([nn.Parameter(t.to(device=device, dtype=torch.float32).contiguous())
                                        for t in leaves])

    # ---- parameter pytree <-> nn.Parameters
    def params(self):
        return _rebuild(self._tree, iter(self.leaves))

    def block_params(self):
        p = self.params()
        return [p[i] for i, k in enumerate(self.kinds) if k == "sde"]

    def head_params(self):
        return self.params()[-1]

    def oracle_ordered_parameters(self):
        out = []
        for (w0, _, fwp) in self.block_params():
            out.append(w0)
            out += [t for lay in arch.fw_layers(fwp) for t in lay]
        (mean, logstd) = self.head_params()
        (Wm, bm), (Wl, bl) = mean[1], logstd[1]
        return out + [Wm, bm, Wl, bl]

    @torch.no_grad()
    def load_oracle_params(self, params_r, head_r):
        """Copy an oracle parameter set (oracle.jax_path.SDEBNNClassifier.init) into this module."""
        flat_r = []
        for (w0, fw) in params_r:
            flat_r.append(w0)
            flat_r += [t for lay in fw for t in lay]
        (Wm, bm), (Wl, bl) = head_r
        flat_r += [Wm, bm, Wl, bl]
        for p, r in zip(self.oracle_ordered_parameters(), flat_r):
            p.copy_(r.to(p.dtype).reshape(p.shape))

    def explicit_rng(self, dWs, eps_head):
        """Per-layer rng list for `bnn_serial`: explicit increments per block, explicit head noise."""
        it = iter(dWs)
        return [next(it) if k == "sde" else ((0, eps_head) if k == "head" else 0) for k in self.kinds]

    # ---- forward / loss
    def forward(self, x, rng):
        """x: (B,H,W,C).  Returns (logp (S,B,classes), kl (S,))."""
        S = self.nsamples
        xs = x.unsqueeze(0).expand(S, *x.shape).contiguous()
        logp, kl, _ = self._predict(self.params(), xs, rng=rng, mc_samples=S)
        return logp, kl

    def loss(self, x, labels, kl_scale, rng):
        """mean over samples of  nll_s + kl_s * kl_scale  (script :27-51, :231)."""
        logp, kl = self.forward(x, rng)
        idx = labels.view(1, -1, 1).expand(logp.shape[0], -1, 1)
        nll = -logp.gather(-1, idx).squeeze(-1).mean(dim=1)  # (S,)
        loss = (nll + kl * kl_scale).mean() if kl_scale > 0 else nll.mean()
        return loss, nll.mean(), kl.mean()


class Trainer:
    """One optimiser step of the script's `update` (:220-245): per-sample grads averaged, Adam.
    With torch.distributed initialised, gradients are all-reduced (mean) once per step over NCCL --
    the only collective on this path."""

    def __init__(self, model, lr=1e-3, kl_coef=1e-3, train_size=45000):
        self.model = model
        self.kl_scale = kl_coef / train_size
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, fused=True)
        self.step_idx = 0
        self._flat = None

    def _allreduce(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)

    def step(self, x, labels, seed=None):
        self.opt.zero_grad(set_to_none=True)
        seed = self.step_idx if seed is None else seed
        loss, nll, kl = self.model.loss(x, labels, self.kl_scale, rng=seed)
        loss.backward()
        self._allreduce()
        self.opt.step()
        self.step_idx += 1
        return loss.detach(), nll.detach(), kl.detach()

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
                 math="fp32", method="euler_maruyama", remat=False, seed=0, device="cuda", sample_offset=0, total_samples=None):
        """`nsamples` weight paths are integrated by THIS module (rank); with the MC samples sharded over ranks
        (samples x batch grid) they are global samples [sample_offset, sample_offset + nsamples) of
        `total_samples`: Philox sample counters and the head's noise rows are the global ones, so G ranks with S/G samples each
        reproduce one rank with S samples."""
        super().__init__()
        self.nsamples, self.input_size = nsamples, tuple(input_size)
        self.sample_offset = int(sample_offset)
        self.total_samples = int(total_samples) if total_samples is not None else nsamples
        if self.sample_offset < 0 or self.sample_offset + nsamples > self.total_samples:
            raise ValueError(f"samples [{sample_offset}, {sample_offset + nsamples}) outside 0..{self.total_samples}")
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
            out += [t for lay in arch.fw_layers(fwp) for t in lay] + arch.fw_betas(fwp)
        (mean, logstd) = self.head_params()
        (Wm, bm), (Wl, bl) = mean[1], logstd[1]
        return out + [Wm, bm, Wl, bl]

    @torch.no_grad()
    def load_oracle_params(self, params_r, head_r):
        """Copy an oracle parameter set (oracle.jax_path.SDEBNNClassifier.init) into this module."""
        flat_r = []
        for (w0, fw) in params_r:
            flat_r.append(w0)
            flat_r += [t for lay in fw for t in lay[:5]] + [lay[5] for lay in fw if len(lay) > 5]   # swish betas last
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
        logp, kl, _ = self._predict(self.params(), xs, rng=rng, mc_samples=S, mc_offset=self.sample_offset,
                                    mc_total=self.total_samples)
        return logp, kl

    def loss(self, x, labels, kl_scale, rng):
        """mean over samples of  nll_s + kl_s * kl_scale  (script :27-51, :231)."""
        logp, kl = self.forward(x, rng)
        idx = labels.view(1, -1, 1).expand(logp.shape[0], -1, 1)
        nll = -logp.gather(-1, idx).squeeze(-1).mean(dim=1)  # (S,)
        loss = (nll + kl * kl_scale).mean() if kl_scale > 0 else nll.mean()
        return loss, nll.mean(), kl.mean()


def sample_shard(total_samples, world, rank):
    """Samples x batch grid of SURVEY 8.5 for `world` ranks: (first global sample, samples per rank, batch group, ranks per
    sample group).  world <= S: S/world samples per rank, every rank sees the whole batch.  world > S: one sample per rank,
    world/S ranks share a sample and split the batch (they recompute the same weight path from the same Philox key)."""
    if total_samples % world == 0:
        per = total_samples // world
        return rank * per, per, 0, 1
    if world % total_samples == 0:
        group = world // total_samples
        return rank // group, 1, rank % group, group
    raise ValueError(f"cannot lay {total_samples} samples on {world} ranks: one must divide the other")


class Trainer:
    """One optimiser step of the script's `update` (:220-245): per-sample grads averaged, Adam.
    With torch.distributed initialised, gradients are all-reduced (mean) once per step over NCCL --
    the only collective on this path.  `overlap=True` launches the all-reduce of each SDEBNN block's
    gradients (one bucket per block, in reverse-pass order) from a post-accumulate hook on a side stream as soon as the
    block's reverse pass has produced them, so the collective runs under the remaining backward work.

    `acc_grad` (script :220-240, `--acc_grad`): the batch is cut into ceil(B / acc_grad)-image chunks whose gradients are
    averaged with equal weights under the SAME sample rngs.  As shipped the loop re-uses chunk 0 for i = 1 and never
    reaches the last chunk (`batch[(i - 1) * bsz : i * bsz]`, :236); `acc_grad_mode="reference"` keeps that, "intended"
    walks chunks 0..acc_grad-1."""

    def __init__(self, model, lr=1e-3, kl_coef=1e-3, train_size=45000, acc_grad=1, acc_grad_mode="reference", overlap=False):
        self.model = model
        self.kl_scale = kl_coef / train_size
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, fused=next(model.parameters()).is_cuda)
        self.step_idx = 0
        self.acc_grad, self.acc_grad_mode = int(acc_grad), acc_grad_mode
        if acc_grad_mode not in ("reference", "intended"):
            raise ValueError(f"acc_grad_mode must be 'reference' or 'intended', got {acc_grad_mode!r}")
        self.overlap = bool(overlap)
        self._buckets, self._pending, self._side = None, [], None

    # ---- gradient all-reduce
    def _dist(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _allreduce(self):
        if not self._dist():
            return
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)

    def _make_buckets(self):
        """One bucket per top-level layer of the network (an SDEBNN block's w0 + f_w tensors; the head), keyed by parameter."""
        leaves = list(self.model.parameters())
        paths = [pth for pth, _ in _leaves_with_paths(self.model._tree)]
        buckets = {}
        for p, pth in zip(leaves, paths):
            buckets.setdefault(pth[0], []).append(p)
        self._buckets = list(buckets.values())
        self._bucket_of = {id(p): i for i, b in enumerate(self._buckets) for p in b}
        self._left = [0] * len(self._buckets)
        for p in leaves:
            p.register_post_accumulate_grad_hook(self._hook)

    def _hook(self, p):
        if not self._armed:
            return
        i = self._bucket_of[id(p)]
        self._left[i] -= 1
        if self._left[i] == 0:
            self._launch_bucket(i)

    def _launch_bucket(self, i):
        grads = [q.grad for q in self._buckets[i]]
        cur = torch.cuda.current_stream() if grads[0].is_cuda else None
        if cur is not None:
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(cur)
            ctx = torch.cuda.stream(self._side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            flat = torch._utils._flatten_dense_tensors(grads)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        self._pending.append((work, flat, grads))

    def _finish_buckets(self):
        w = dist.get_world_size()
        for i, left in enumerate(self._left):   # a bucket with a leaf that received no gradient was never launched
            if left > 0:
                live = [q for q in self._buckets[i] if q.grad is not None]
                if live:
                    flat = torch._utils._flatten_dense_tensors([q.grad for q in live])
                    self._pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, [q.grad for q in live]))
        for work, flat, grads in self._pending:
            work.wait()   # NCCL: makes the current stream wait for the collective; gloo: blocks
            if flat.is_cuda:
                flat.record_stream(torch.cuda.current_stream())
            flat.div_(w)
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)
        self._pending = []

    # ---- one step
    def _chunks(self, n):
        if self.acc_grad <= 1:
            return [slice(0, n)]
        bsz = -(-n // self.acc_grad)  # math.ceil(batch / acc_grad), script :222
        if self.acc_grad_mode == "reference":
            idx = [0] + [i - 1 for i in range(1, self.acc_grad)]       # first_batch, then (i - 1) * bsz : i * bsz  (:223, :236)
        else:
            idx = list(range(self.acc_grad))
        return [slice(i * bsz, min((i + 1) * bsz, n)) for i in idx]

    def step(self, x, labels, seed=None):
        self.opt.zero_grad(set_to_none=True)
        seed = self.step_idx if seed is None else seed
        chunks = self._chunks(x.shape[0])
        overlap = self.overlap and self._dist() and len(chunks) == 1
        if overlap and self._buckets is None:
            self._make_buckets()
        self._armed = overlap
        if overlap:
            self._left = [len(b) for b in self._buckets]
        tot = None
        for sl in chunks:
            loss, nll, kl = self.model.loss(x[sl], labels[sl], self.kl_scale, rng=seed)
            (loss / len(chunks)).backward()     # running mean of :237 == equal-weight mean of the chunk gradients
            vals = torch.stack([loss.detach(), nll.detach(), kl.detach()]) / len(chunks)
            tot = vals if tot is None else tot + vals
        self._armed = False
        if overlap:
            self._finish_buckets()
        else:
            self._allreduce()
        self.opt.step()
        self.step_idx += 1
        return tot[0], tot[1], tot[2]

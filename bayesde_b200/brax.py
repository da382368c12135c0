"""Host-side mirror of brax/_impl/brax.py: `SDEBNN`, `MeanField`, `normal_logprob`, `bnn_serial`,
with the reference's constructor keywords and `(init_fun, apply_fun)` pairs, on torch CUDA tensors.

Differences a user of the reference needs to know (all additive):
  * MC samples are integrated together: `inputs` may be (B,H,W,C) (one sample) or (S,B,H,W,C); `kl`
    comes back per sample, shape (S,) (the reference vmaps a one-sample function over rngs,
    examples/jax/sdebnn_classification.py:218).
  * `rng` is an int seed / (seed, stream) pair for the in-kernel Philox stream, or a tensor of explicit
    increments (S, nsteps-1, P) for exact comparison with the reference.
  * extra keywords `method`, `time_grid`, `math` select the scheme ("euler_maruyama" as brax.py:113),
    the scan's time bookkeeping ("reference" = sdeint.py:105-106 as shipped, "uniform" = intended) and
    the conv arithmetic ("fp32" CUDA cores, "tc" tcgen05 tensor cores).
"""
import math

import numpy as np
import torch

from . import _fused, _lib
from .arch import FxSpec, Layer, _gen, fw_betas, fw_layers, split
from .sdeint import sdeint_ito, sdeint_ito_fixed_grid


class _AugDynamics:
    """f_aug / g_aug of one SDEBNN block (brax.py:49-78) as a structured object: passing its bound
    methods to `sdeint_ito_fixed_grid` selects the fused kernels."""
    _bsde_structured = True

    def __init__(self, fx_spec, fw_spec, x_shape, diff_coef, stl, w_drift, math_mode, w_stale, stream_id, remat=False):
        self.remat = remat
        self.fx, self.fw = fx_spec, fw_spec
        self.x_shape = tuple(x_shape)  # (S,B,H,W,C)
        self.diff_coef, self.stl, self.w_drift, self.math = diff_coef, stl, w_drift, math_mode
        self.w_stale, self.stream_id = w_stale, stream_id
        self.x_dim = int(np.prod(self.x_shape[1:]))  # per sample, as brax.py:45
        self.P = fx_spec.P
        self.w0_per_sample = False
        self.sample_offset = 0

    # f_aug / g_aug (brax.py:49-78) as plain callables: VALUES only (no autograd), evaluated by the same kernels as the fused
    # solve -- one Euler iteration of K1 / K2 with dt = 1 and zero noise gives y + f(y), so f = (y + f) - y (exact to the fp32
    # rounding of that sum).  Passing the bound methods to `sdeint_ito_fixed_grid` selects the fused solve instead.
    def _one_step(self, y, t, args, sigma_noise):
        S, P, x_dim = self.x_shape[0], self.P, self.x_dim
        yy = y.reshape(S, x_dim + 2 * P).to(torch.float32)
        x = yy[:, :x_dim].reshape(self.x_shape).contiguous()
        w = yy[:, x_dim:x_dim + P].contiguous()
        cfg, layers = self._config(2, np.asarray([0.0, 1.0], dtype=np.float32), P if self.stl else 0, "uniform", args, y.device)
        cfg.set_step(float(t), 1.0)
        cfg.want_kltraj = True
        dW = sigma_noise.reshape(S, 1, P).to(torch.float32).contiguous() if sigma_noise is not None else torch.zeros(S, 1, P, device=y.device)
        with torch.no_grad():
            wtraj, _, _ = _fused.wsde_forward(cfg, w, layers, dW, 0, 0, self.w_stale, self._betas)
            xs, _ = _fused.xpath_forward(cfg, x, wtraj, keep_acts=False)
        return yy, xs[1].reshape(S, x_dim), wtraj[:, 1], cfg.kltraj[:, 1]

    def f_aug(self, y, t, args):
        """brax.py:49-62: [dx, dw, dkl] at (y, t).  Values only; use `sdeint_ito_fixed_grid(f_aug, g_aug, ...)` for the solve."""
        yy, x1, w1, kl1 = self._one_step(y, t, args, None)
        x_dim, P = self.x_dim, self.P
        out = torch.cat([x1 - yy[:, :x_dim], w1 - yy[:, x_dim:x_dim + P], kl1], dim=1)
        return out.reshape(y.shape)

    def g_aug(self, y, t, args):
        """brax.py:64-78: [0_x, sigma 1_P, stl ? u(w_stale, t; stop_gradient(params)) : 0_P]  (quirk Q2: the stale vector)."""
        S, P, x_dim = self.x_shape[0], self.P, self.x_dim
        out = torch.zeros(S, x_dim + 2 * P, device=y.device, dtype=torch.float32)
        out[:, x_dim:x_dim + P] = self.diff_coef
        if self.stl:
            # the kl row's noise coefficient: a unit increment on every w row adds exactly u_stale to the kl rows
            _, _, _, kl_a = self._one_step(y, t, args, torch.ones(S, P, device=y.device))
            _, _, _, kl_b = self._one_step(y, t, args, None)
            out[:, x_dim + P:] = kl_a - kl_b
        return out.reshape(y.shape)

    def _unpack(self, y0, rep):
        S, B, H, W, Cc = self.x_shape
        P, x_dim = self.P, self.x_dim
        y0 = y0.reshape(S, x_dim + 2 * P)
        x = y0[:, :x_dim].reshape(self.x_shape)
        w0 = y0[:, x_dim:x_dim + P]
        if not self.w0_per_sample:
            w0 = w0[0]
        if rep not in (0, P):
            raise ValueError(f"rep must be 0 or w_dim={P} for SDEBNN dynamics, got {rep}")
        return x, w0

    def _config(self, nsteps, ts_np, rep, time_grid, args, device):
        S, B = self.x_shape[:2]
        cfg = _fused.BlockConfig(self.fx, self.fw, S, B, nsteps, self.diff_coef, bool(rep), time_grid, self.math,
                                 ts=ts_np, remat=self.remat)
        cfg.sample_offset = self.sample_offset
        self._betas = []
        if self.w_drift:
            layers = fw_layers(args)
            self._betas = fw_betas(args)
        else:
            layers = _fused.zero_fw_layers(self.P, device)
            cfg.fw = dict(hidden_dims=(1,), actfn="softplus", ou_dw=False)
        return cfg, layers

    def _jax_trees(self, keys, tsl, rep, device):
        """One reference-layout tree per sample for JAX keys (`vmap` over rngs, examples/jax/sdebnn_classification.py:218): the state
        of a sample is [x (x_dim), w (P), kl (P)] with rep tying (brax.py:110-111); only the w rows [x_dim, x_dim + P) are drawn."""
        from .sdeint import BrownianTree
        keys = np.asarray(keys, dtype=np.uint32).reshape(-1, 2)
        S, P = self.x_shape[0], self.P
        if keys.shape[0] != S:
            raise ValueError(f"need one PRNGKey per sample: {S} samples, {keys.shape[0]} keys")
        return [BrownianTree(tsl[0], self.x_dim + 2 * P, tsl[-1], k, depth=10, rep=rep, device=device) for k in keys]

    def _solve(self, y0, ts, rng, args, method, rep, time_grid, want_ys=False):
        """y0: (S, x_dim + 2P) flat states (brax.py:110) or (x_dim + 2P,).  Returns the layer's view of the solution,
        (x_T, kl [S], wtraj [S, T, P]) -- and, with `want_ys`, the reference's `ys` (T, D) (sdeint.py:138) assembled from the
        scan's x checkpoints, the w trajectory and the kl-row trajectory."""
        S, P = self.x_shape[0], self.P
        x, w0 = self._unpack(y0, rep)
        # method: g_aug is state-independent as shipped (quirk Q2), so the Milstein correction
        # (sdeint.py:36-39) and the Heun average are identically zero / equal: all three coincide.
        nsteps = len(ts)
        ts_np = ts.detach().cpu().numpy() if isinstance(ts, torch.Tensor) else np.asarray(ts, dtype=np.float32)
        cfg, layers = self._config(nsteps, ts_np, rep, time_grid, args, y0.device)
        if _is_jax_keys(rng):  # replay of a JAX run: increments of the reference's own threefry tree
            from .sdeint import tree_increments
            tsl = [float(v) for v in ts_np]
            rng = torch.stack([tree_increments(tr, tsl, time_grid, self.x_dim, P)
                               for tr in self._jax_trees(rng, tsl, rep, y0.device)])
        if isinstance(rng, torch.Tensor):
            dW, seed, sid = rng.to(torch.float32).reshape(S, cfg.nit, P).contiguous(), 0, 0
        elif isinstance(rng, (tuple, list)):
            dW, seed, sid = None, int(rng[0]), int(rng[1])
        else:
            dW, seed, sid = None, int(rng), self.stream_id
        flat = [t for lay in layers for t in lay] + list(self._betas)
        cfg.track = torch.is_grad_enabled()
        cfg.want_kltraj = bool(want_ys)
        xT, kl, wtraj = _fused.SDEBNNSolve.apply(cfg, dW, seed, sid, self.w_stale, x, w0, *flat)
        if not want_ys:
            return xT, kl, wtraj
        # ys[k] = [x_k, w_k, kl_k] per sample.  Differentiable where the reference's own consumer differentiates it
        # (brax.py:117-119): the x rows and the SUM of the kl rows of ys[-1]; the other rows are values.
        T = cfg.nit + 1
        ys = torch.cat([cfg.xs.reshape(T, S, -1), wtraj.transpose(0, 1), cfg.kltraj.transpose(0, 1)], dim=2)
        last = torch.cat([xT.reshape(S, -1), wtraj[:, -1],
                          cfg.kltraj[:, -1] + ((kl - kl.detach()) / P).unsqueeze(1)], dim=1)
        ys = torch.cat([ys[:-1], last.unsqueeze(0)], dim=0)
        cfg.xs = cfg.kltraj = None
        return ys.reshape(T, -1) if y0.dim() == 1 else ys

    def _solve_adjoint(self, y0, ts, rng, args, dt, method, rep):
        """`sdeint_ito` (jaxsde/jaxsde/sdeint.py:11-13) for these dynamics: steps of `dt` on a Philox Brownian tree over the
        S*P weight rows (the x rows have g = 0; with rep = P the kl rows reuse the w rows' path, brownian.py:69-74), reverse
        pass by the stochastic adjoint.  Returns (x_T, kl [S], w at the output times [S, len(ts), P])."""
        from .sdeint import BrownianTree
        S, P = self.x_shape[0], self.P
        x, w0 = self._unpack(y0, rep)
        if isinstance(rng, torch.Tensor):
            raise ValueError("sdeint_ito needs a seed, a (seed, stream_id) pair or PRNGKeys: its Brownian path is queried at arbitrary times")
        tsl = [float(v) for v in (ts.detach().cpu().tolist() if isinstance(ts, torch.Tensor) else ts)]
        cfg, layers = self._config(2, np.asarray([0.0, 1.0], dtype=np.float32), rep, "uniform", args, y0.device)
        if _is_jax_keys(rng):
            trees, x_dim = self._jax_trees(rng, tsl, rep, y0.device), self.x_dim
            bm = lambda t: torch.cat([tr.query(t, x_dim, P) for tr in trees])
        else:
            seed, sid = (int(rng[0]), int(rng[1])) if isinstance(rng, (tuple, list)) else (int(rng), self.stream_id)
            bm = BrownianTree(tsl[0], S * P, tsl[-1], (seed, sid), depth=10, rep=0, device=y0.device)
        flat = [t for lay in layers for t in lay] + list(self._betas)
        return _fused.SDEBNNAdjointSolve.apply(cfg, tsl, float(dt), bm, self.w_stale, x, w0, *flat)


def _is_jax_keys(rng):
    return isinstance(rng, np.ndarray) and rng.dtype == np.uint32 and rng.shape[-1] == 2


def SDEBNN(fx_block_type, fx_dim, fx_actfn, fw, diff_coef=1e-4, name="sdebnn", stl=False, xt=False, nsteps=20,
           remat=False, w_drift=True, stax_api=False, infer_initial_state=False, initial_state_prior_std=0.1,
           method="euler_maruyama", time_grid="reference", math="fp32", stream_id=0):
    """brax/_impl/brax.py:15-135.  Returns Layer(init_fun, apply_fun).

    `remat=True`: the reverse pass recomputes the conv intermediates of every scan iteration from the
    stored x_k (the jax.checkpoint analogue, brax.py:131-132); `remat=False` (default, as in the reference)
    keeps them from the forward pass (more memory, ~15 % less arithmetic)."""
    if not xt:
        raise NotImplementedError("SDEBNN(xt=False): build_fx layers are time-dependent (arch.py:176-205); "
                                  "the reference script always passes xt=True")
    if not hasattr(fw, "spec") or fw.spec is None:
        raise NotImplementedError("fw must be built with bayesde_b200.arch.MLP for the fused path")
    ts = torch.linspace(0.0, 1.0, nsteps)  # brax.py:31
    cache = {}

    def make_layer(input_shape, device=None):
        key = tuple(input_shape)
        if key not in cache:
            fx = FxSpec(fx_block_type, (1,) + tuple(input_shape[-3:]), fx_dim, fx_actfn)
            # brax.py:38-41: g_aug's `flat_w` is a throw-away fx.init from PRNGKey(0) (quirk Q2)
            cache[key] = (fx, fx.init(0))
        if device is None:
            return cache[key]
        # the device copy is made once: a pageable host -> device copy per call stalls the host until the stream has drained
        # (one pipeline bubble per block and step: tools/step_gaps.py)
        dkey = (key, str(device))
        if dkey not in cache:
            cache[dkey] = (cache[key][0], cache[key][1].to(device))
        return cache[dkey]

    def init_fun(rng, input_shape, device=None):
        fx, _ = make_layer(input_shape)
        init_w0 = fx.init(rng).to(device)
        logstd_w0 = (torch.zeros_like(init_w0) - 4.0) if infer_initial_state else ()
        if w_drift:
            out_shape, fw_params = fw.init(rng, (fx.P,))  # same rng as fx.init, brax.py:90
            assert tuple(out_shape) == (fx.P,), "fw needs to have the same input and output shapes"
            fw_params = _to_device(fw_params, device)
        else:
            fw_params = ()
        return tuple(input_shape), (init_w0, logstd_w0, fw_params)

    def apply_fun(params, inputs, rng, full_output=False, fixed_grid=True, **kwargs):
        init_w0, logstd_w0, fw_params = params
        x = inputs
        squeeze_s = x.dim() == 4
        if squeeze_s:
            x = x.unsqueeze(0)
        S = x.shape[0]
        fx, w_stale = make_layer(x.shape[1:], x.device)
        P = fx.P
        dyn = _AugDynamics(fx, fw.spec, x.shape, diff_coef, stl, w_drift, math, w_stale, stream_id,
                           remat=remat)
        # ranks that shard the MC samples: local sample s is global sample mc_offset + s of mc_total (Philox sample counter)
        dyn.sample_offset = int(kwargs.get("mc_offset", 0))
        S_total = int(kwargs.get("mc_total", 0)) or S
        kl = x.new_zeros(S)
        if infer_initial_state:  # brax.py:100-106
            if isinstance(rng, (tuple, list)) and len(rng) == 2 and isinstance(rng[1], torch.Tensor):
                eps, rng = rng[1], rng[0]
            else:
                w0_rng, rng = split(rng, 2)
                eps = torch.randn(S_total, P, generator=_gen(w0_rng, x.device), device=x.device)
                eps = eps[dyn.sample_offset:dyn.sample_offset + S]
            mean_w0 = init_w0
            w0 = eps * torch.exp(logstd_w0) + mean_w0
            kl = (normal_logprob(w0, mean_w0, logstd_w0)
                  - normal_logprob(w0, 0.0, math_log(initial_state_prior_std))).sum(-1)
            dyn.w0_per_sample = True
        else:
            w0 = init_w0.expand(S, P)
        y0 = torch.cat([x.reshape(S, -1), w0, x.new_zeros(S, P)], dim=1)  # brax.py:110
        rep = P if stl else 0  # brax.py:111
        if fixed_grid:
            xo, kl_sde, wtraj = sdeint_ito_fixed_grid(dyn.f_aug, dyn.g_aug, y0, ts, rng, fw_params, method=method,
                                                      rep=rep, time_grid=time_grid, _layer_view=True)
        else:  # brax.py:114-116 ("using stochastic adjoint"); `dt` (keyword) is sdeint_ito's step, default 1e-6 as there
            xo, kl_sde, wtraj = sdeint_ito(dyn.f_aug, dyn.g_aug, y0, ts, rng, fw_params, dt=kwargs.get("dt", 1e-6),
                                           method="euler_maruyama", rep=rep)
        kl = kl + kl_sde  # brax.py:119
        if squeeze_s:
            xo, kl = xo[0], kl[0]
        if stax_api:
            return xo
        if full_output:
            w = wtraj[0] if squeeze_s else wtraj
            return xo, kl, {name + "_w": w}
        return xo, kl

    return Layer(init_fun, apply_fun)


def math_log(v):
    return math.log(v)


def _to_device(tree, device):
    if isinstance(tree, torch.Tensor):
        return tree.to(device)
    return type(tree)(_to_device(t, device) for t in tree)


def tree_leaves(tree):
    if isinstance(tree, torch.Tensor):
        return [tree]
    out = []
    for t in tree:
        out.extend(tree_leaves(t))
    return out


def tree_map(fn, tree):
    if isinstance(tree, torch.Tensor):
        return fn(tree)
    return type(tree)(tree_map(fn, t) for t in tree)


def _unflatten_like(flat, tree, lead=0):
    """Inverse of ravel_pytree for `tree`; `flat` may carry `lead` leading (sample) dims."""
    pos = [0]

    def rec(t):
        if isinstance(t, torch.Tensor):
            n = t.numel()
            v = flat[..., pos[0]:pos[0] + n].reshape(*flat.shape[:lead], *t.shape)
            pos[0] += n
            return v
        return type(t)(rec(q) for q in t)

    return rec(tree)


def normal_logprob(z, mean, log_std):
    """brax.py:171-178."""
    c = math.log(2 * math.pi)
    if isinstance(log_std, (int, float)):
        # a Python scalar stays on the host: a device tensor made from it is a pageable host -> device copy, which stalls the
        # host until the stream has drained (the head's KL sits between the forward and the reverse pass of every step)
        tmp = (z - mean) * math.exp(-log_std)
        return -0.5 * (tmp * tmp + (2 * log_std + c))
    log_std = torch.as_tensor(log_std, dtype=z.dtype, device=z.device)
    tmp = (z - mean) * torch.exp(-log_std)
    return -0.5 * (tmp * tmp + 2 * log_std + c)


def MeanField(layer, prior_std=0.1, disable=False):
    """brax.py:138-168: mean-field Gaussian over the wrapped layer's parameters.

    With inputs carrying a leading sample axis (S,B,...) one parameter draw is made per sample and
    the returned kl has shape (S,).  `rng` may be `(seed, eps)` with eps (S, nparams) explicit noise."""
    init_fun, apply_fun = layer

    def wrapped_init_fun(rng, input_shape):
        output_shape, params_mean = init_fun(rng, input_shape)
        params_logstd = tree_map(lambda x: torch.zeros_like(x) - 4.0, params_mean)
        return output_shape, (params_mean, params_logstd)

    def wrapped_apply_fun(params, input, rng, **kwargs):
        params_mean, params_logstd = params
        leaves = tree_leaves(params_mean)
        sampled = kwargs.pop("mc_samples", None)
        S = sampled if sampled is not None else None
        # sample-sharded ranks draw the GLOBAL (mc_total, n) noise and keep their rows [mc_offset, mc_offset + S)
        mc_off, mc_tot = int(kwargs.get("mc_offset", 0)), int(kwargs.get("mc_total", 0)) or S
        # brax.py:150,165: the key is split and the wrapped layer receives `rng=next_rng` (a stochastic wrapped layer, e.g.
        # MeanField(SDEBNN(..., stax_api=True)) of --meanfield_sdebnn, draws its Brownian path from it)
        explicit = isinstance(rng, (tuple, list)) and len(rng) == 2 and isinstance(rng[1], torch.Tensor)
        base = rng[0] if explicit else rng
        rng_eps, rng_next = split(base, 2) if base is not None else (None, None)
        if not leaves:  # parameter-free layer: kl = sum over an empty vector
            out = apply_fun(params_mean, input, rng=rng_next, **kwargs)
            return out, input.new_zeros(S) if S else input.new_zeros(())
        flat_mean = torch.cat([t.reshape(-1) for t in leaves])
        flat_logstd = torch.cat([t.reshape(-1) for t in tree_leaves(params_logstd)])
        if explicit:
            eps = rng[1]
        else:
            shape = (mc_tot, flat_mean.numel()) if S else (flat_mean.numel(),)
            eps = torch.randn(shape, generator=_gen(rng_eps, flat_mean.device), device=flat_mean.device)
            if S:
                eps = eps[mc_off:mc_off + S]
        if disable:
            flat_params = flat_mean.expand_as(eps) if S else flat_mean
            kl = flat_params.new_zeros(flat_params.shape)
        else:
            flat_params = eps * torch.exp(flat_logstd) + flat_mean  # brax.py:155
            kl = normal_logprob(flat_params, flat_mean, flat_logstd) - normal_logprob(
                flat_params, 0.0, math.log(prior_std))
        p = _unflatten_like(flat_params, params_mean, lead=1 if S else 0)
        output = apply_fun(p, input, rng=rng_next, **kwargs)
        return output, kl.sum(-1)

    return Layer(wrapped_init_fun, wrapped_apply_fun)


def bnn_serial(*layers):
    """brax.py:181-219.  `rng` may be a list with one entry per layer (explicit per-layer seeds /
    increments / (seed, eps) pairs) instead of a seed that is split."""
    nlayers = len(layers)
    init_funs, apply_funs = zip(*layers)

    def init_fun(rng, input_shape):
        params = []
        for init_f in init_funs:
            rng, layer_rng = split(rng, 2)
            input_shape, param = init_f(layer_rng, input_shape)
            params.append(param)
        return input_shape, params

    def apply_fun(params, inputs, **kwargs):
        rng = kwargs.pop("rng", None)
        if isinstance(rng, list) and len(rng) == nlayers:
            rngs = rng
        else:
            rngs = split(rng, nlayers) if rng is not None else (None,) * nlayers
        total_kl = 0
        infodict = {}
        for fun, param, r in zip(apply_funs, params, rngs):
            output = fun(param, inputs, rng=r, **kwargs)
            if len(output) == 2:
                inputs, layer_kl = output
            elif len(output) == 3:
                inputs, layer_kl, info = output
                infodict.update(info)
            else:
                raise RuntimeError(f"Expected 2 or 3 outputs but got {len(output)}.")
            total_kl = total_kl + layer_kl
        return inputs, total_kl, infodict

    return Layer(init_fun, apply_fun)

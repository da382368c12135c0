"""Config 1 (row a20): host-side mirror of brax/_impl/layers.py for the JAX toy -- `SDELayer` and the layer constructors
examples/jax/sdebnn_toy1d.py:32-95 (`mlp`) builds it from -- on the fused kernels bsde_sdelayer_fwd / bsde_sdelayer_bwd
(csrc/sdelayer.cu: the whole 19-step posterior or prior solve in one launch, one CTA per MC sample).

Same names, argument order and return values as the reference:

    ConcatSquashLineartx(out_dim)                     layers.py:19-42
    DiffEqWrappertx(layer)                            layers.py:45-52
    Const(c) / Affine(mult, const)                    layers.py:251-269
    Additive(prior_layer, post_layer)                 layers.py:272-292
    AugmentedLayer(pad)                               (zero-pads the last axis, as arch.Augment)
    Swish_(out_dim) / ACTIVATIONS                     layers.py:296-326
    SDELayer(driftx, prior_driftw, post_driftw, diffusion, x_dim, w_dim, setting) -> (init_fun, apply_fun)     layers.py:164-248
    mlp(rng, x_dim, hidden_dim, activn, setting)      examples/jax/sdebnn_toy1d.py:32-95

`apply_fun(params, inputs, rng, entire=False)` returns `(x, w, kl, prior_x, prior_w, prior_kl)` at t = 1 (ts = linspace(0, 1, 20),
layers.py:242-245) or, with entire=True, the trajectories on linspace(0, 1, 90) (layers.py:233-240).  The layers are ordinary
(init, apply) pairs evaluated by torch (used for `init` and for callers that evaluate a layer by itself); they carry a `.spec`
that the fused solve understands.  A structure the kernel does not implement raises NotImplementedError -- there is no eager
fallback for the solve.

Additions: `rng` is an int seed / (seed, stream_id) pair for the in-kernel Philox stream, or a tensor of unit normals of the w rows
((nt-1, P) or (S, nt-1, P)) for exact comparison; `num_samples=S` integrates S samples at once (the reference vmaps over rngs,
sdebnn_toy1d.py:127) and every output gains a leading sample axis.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .arch import Layer, _gen, split


class _Spec(Layer):
    spec = None


def _with_spec(init_fun, apply_fun, **spec):
    lay = _Spec(init_fun, apply_fun)
    lay.spec = spec
    return lay


def ConcatSquashLineartx(out_dim, W_init="he_normal", b_init="normal"):
    """y = sigmoid(w_t t + w_tb) (x W + b) + b_t t on inputs (t, x).  `W_init` / `b_init`: "he_normal" / "normal" (std 1e-2) or
    "zeros" (sdebnn_toy1d.py:69)."""
    def init_fun(rng, input_shape):
        g = _gen(rng)
        din = int(input_shape[-1])
        sw = 0.0 if W_init == "zeros" else math.sqrt(2.0 / din)
        sb = 0.0 if b_init == "zeros" else 1e-2
        mk = lambda *s, sd: torch.randn(*s, generator=g) * sd
        return tuple(input_shape[:-1]) + (out_dim,), (mk(din, out_dim, sd=sw), mk(out_dim, sd=sb), mk(out_dim, sd=sb), mk(out_dim, sd=sb),
                                                      mk(out_dim, sd=sb))

    def apply_fun(params, inputs, **kwargs):
        t, x = inputs
        W, b, w_t, w_tb, b_t = params
        return t, (x @ W + b) * torch.sigmoid(w_t * t + w_tb) + b_t * t

    return _with_spec(init_fun, apply_fun, kind="csl", out_dim=out_dim)


def DiffEqWrappertx(layer):
    init_fun, layer_apply = layer[0], layer[1]

    def apply_fun(params, inputs, **kwargs):
        t, x = inputs
        return t, layer_apply(params, x, **kwargs)

    return _with_spec(init_fun, apply_fun, kind="wrap", inner=getattr(layer, "spec", None))


def Const(const=1e-3):
    return _with_spec(lambda rng, s: (s, ()), lambda params, inputs, **kw: torch.ones_like(inputs) * const, kind="const", const=float(const))


def Affine(mult=0.0, const=1e-3):
    return _with_spec(lambda rng, s: (s, ()), lambda params, inputs, **kw: inputs * mult + const, kind="affine", mult=float(mult),
                      const=float(const))


def Additive(layer1, layer2):
    def init_fun(rng, input_shape):
        o1, p1 = layer1[0](rng, input_shape)
        _, p2 = layer2[0](rng, input_shape)
        return o1, (p1, p2)

    def apply_fun(params, inputs, **kwargs):
        p1, p2 = params
        _, out1 = layer1[1](p1, inputs, **kwargs)
        t2, out2 = layer2[1](p2, inputs, **kwargs)
        return t2, out2 + out1

    return _with_spec(init_fun, apply_fun, kind="additive", first=getattr(layer1, "spec", None), second=getattr(layer2, "spec", None),
                      layers=(layer1, layer2))


def AugmentedLayer(pad_zeros):
    def init_fun(rng, input_shape):
        return tuple(input_shape[:-1]) + (input_shape[-1] + pad_zeros,), ()

    return _with_spec(init_fun, lambda params, inputs, **kw: F.pad(inputs, (0, pad_zeros)), kind="augment", pad=pad_zeros)


def Swish_(out_dim, beta_init=0.5):
    return _with_spec(lambda rng, s: (s, (torch.ones(out_dim) * beta_init,)),
                      lambda params, x, **kw: x * torch.sigmoid(x * F.softplus(params[0])), kind="act", act="swish", out_dim=out_dim)


def _elementwise(name, fn):
    return _with_spec(lambda rng, s: (s, ()), lambda params, x, **kw: fn(x), kind="act", act=name)


ACTIVATIONS = {"softplus": _elementwise("softplus", F.softplus), "tanh": _elementwise("tanh", torch.tanh), "elu": _elementwise("elu", F.elu),
               "rbf": _elementwise("rbf", lambda x: torch.exp(-x ** 2)), "swish": Swish_}


def get_activn(name):
    """brax.utils.registry.get_activn for the activations the fused kernels implement ('relu' / 'swish_nobeta' of the reference's
    registry are not: relu's kink and the parameter-free swish have no kernel path)."""
    if name not in ACTIVATIONS:
        raise NotImplementedError(f"activation {name!r} is not available in the fused B200 path")
    return ACTIVATIONS[name]


def serial_tx(*layers):
    """stax.serial over (t, x) layers, keeping the member specs."""
    def init_fun(rng, input_shape):
        params, shape = [], input_shape
        for lay, k in zip(layers, split(rng, len(layers))):
            shape, p = lay[0](k, shape)
            params.append(p)
        return shape, params

    def apply_fun(params, inputs, **kwargs):
        for lay, p in zip(layers, params):
            inputs = lay[1](p, inputs, **kwargs)
        return inputs

    return _with_spec(init_fun, apply_fun, kind="serial", members=[getattr(l, "spec", None) for l in layers])


def shape_dependent(make_layer):
    """layers.shape_dependent: the layer is built from the input shape at init / apply time."""
    cache = {}

    def get(shape):
        key = int(shape[-1])
        if key not in cache:
            cache[key] = make_layer(shape)
        return cache[key]

    def init_fun(rng, input_shape):
        return get(input_shape)[0](rng, input_shape)

    def apply_fun(params, inputs, **kwargs):
        return get(inputs[1].shape)[1](params, inputs, **kwargs)

    lay = _with_spec(init_fun, apply_fun, kind="shape_dependent")
    lay.get = get
    return lay


# ----------------------------------------------------------------------------------------------------------------------------
# structure recognition: what the kernel implements
# ----------------------------------------------------------------------------------------------------------------------------
def _mlp_spec(layer, in_dim):
    """(H, act) of a CSL(H) act CSL(H) act CSL(in_dim) stack, else NotImplementedError."""
    if (getattr(layer, "spec", None) or {}).get("kind") == "shape_dependent":
        layer = layer.get((-1, in_dim))
    sp = getattr(layer, "spec", None) or {}
    m = sp.get("members") or []
    ok = (sp.get("kind") == "serial" and len(m) == 5 and all(m[i] and m[i].get("kind") == "csl" for i in (0, 2, 4))
          and all(m[i] and m[i].get("kind") == "wrap" and (m[i].get("inner") or {}).get("kind") == "act" for i in (1, 3)))
    if not ok or m[0]["out_dim"] != m[2]["out_dim"] or m[4]["out_dim"] != in_dim or m[1]["inner"]["act"] != m[3]["inner"]["act"]:
        raise NotImplementedError("the fused SDELayer implements ConcatSquashLineartx(H) -> act -> ConcatSquashLineartx(H) -> act -> "
                                  "ConcatSquashLineartx(in_dim) stacks (examples/jax/sdebnn_toy1d.py:48-83)")
    return m[0]["out_dim"], m[1]["inner"]["act"]


def _const_of(layer, what):
    sp = getattr(layer, "spec", None) or {}
    inner = sp.get("inner") or {}
    if sp.get("kind") == "wrap" and inner.get("kind") == "const":
        return ("const", inner["const"])
    if sp.get("kind") == "wrap" and inner.get("kind") == "affine":
        return ("affine", inner["mult"], inner["const"])
    raise NotImplementedError(f"{what}: the fused SDELayer implements DiffEqWrappertx(Const(c)) / DiffEqWrappertx(Affine(-1, 0))")


def _csl_leaves(tree):
    return [q for q in tree if len(q) == 5], [q[0] for q in tree if len(q) == 1]


class _SDELayerSolve(torch.autograd.Function):
    """All states of `sdeint(f, g, [w0, x0, 0], P, ts, rng)` for S samples: (S, nt, P + B D + 1), row 0 = initial state."""

    @staticmethod
    def forward(ctx, cfg, eps, ts, x0, w0, *fw_flat):
        L = _lib.lib()
        dev = x0.device
        d = cfg["desc"]()
        S, nt, P, BD = d.S, d.nt, d.P, d.B * d.D
        lays = [tuple(t.contiguous() for t in fw_flat[5 * i:5 * i + 5]) for i in range(3)] if fw_flat else []
        betas = [t.contiguous() for t in fw_flat[15:]]
        fp = _driftw_struct(lays, betas, P, d.H, cfg["act"]) if lays else None
        traj = torch.empty(S, nt, P + BD + 1, device=dev, dtype=torch.float32)
        x0c, w0c, tsc = x0.contiguous().float(), w0.contiguous().float(), ts.contiguous()
        st = L.bsde_sdelayer_fwd(C.byref(d), C.byref(fp) if fp is not None else None, _lib.ptr(tsc), _lib.ptr(w0c), _lib.ptr(x0c), _lib.ptr(eps),
                                 _lib.ptr(traj), _lib.stream_ptr())
        _lib.check(st, "bsde_sdelayer_fwd")
        ctx.cfg, ctx.eps, ctx.lays, ctx.betas = cfg, eps, lays, betas
        ctx.save_for_backward(tsc, traj)
        ctx.x_shape, ctx.w_shape = x0.shape, w0.shape
        return traj

    @staticmethod
    def backward(ctx, gtraj):
        L = _lib.lib()
        cfg = ctx.cfg
        tsc, traj = ctx.saved_tensors
        d = cfg["desc"]()
        S, P, BD = d.S, d.P, d.B * d.D
        dev = traj.device
        glays = [tuple(torch.zeros((S,) + tuple(t.shape), device=dev) for t in lay) for lay in ctx.lays]
        gbetas = [torch.zeros((S,) + tuple(t.shape), device=dev) for t in ctx.betas]
        fp = _driftw_struct(ctx.lays, ctx.betas, P, d.H, cfg["act"]) if ctx.lays else None
        gfp = _driftw_struct(glays, gbetas, P, d.H, cfg["act"]) if ctx.lays else None
        g0 = torch.empty(S, P + BD + 1, device=dev, dtype=torch.float32)
        gt = gtraj.contiguous().float()
        st = L.bsde_sdelayer_bwd(C.byref(d), C.byref(fp) if fp is not None else None, _lib.ptr(tsc), _lib.ptr(ctx.eps), _lib.ptr(traj),
                                 _lib.ptr(gt), None, _lib.ptr(g0), C.byref(gfp) if gfp is not None else None, _lib.stream_ptr())
        _lib.check(st, "bsde_sdelayer_bwd")
        gw0 = g0[:, :P] if len(ctx.w_shape) == 2 else g0[:, :P].sum(0)
        gx0 = g0[:, P:P + BD].sum(0).reshape(ctx.x_shape)
        flat = [t.sum(0) for lay in glays for t in lay] + [t.sum(0) for t in gbetas]
        return (None, None, None, gx0, gw0, *flat)


def _driftw_struct(lays, betas, P, H, act):
    fp = _lib.FwParams()
    fp.P, fp.nhid, fp.act = P, 2, _lib.ACT[act]
    fp.dims[0] = fp.dims[1] = H
    for name, t in zip(("W1", "b1", "wt1", "wtb1", "bt1"), lays[0]):
        setattr(fp, name, _lib.ptr(t).value)
    for name, t in zip(("Wm", "bm", "wtm", "wtbm", "btm"), lays[1]):
        getattr(fp, name)[0] = _lib.ptr(t).value
    for name, t in zip(("W4", "b4", "wt4", "wtb4", "bt4"), lays[2]):
        setattr(fp, name, _lib.ptr(t).value)
    if act == "swish":
        fp.beta1 = _lib.ptr(betas[0]).value
        fp.betam[0] = _lib.ptr(betas[1]).value
    return fp


def SDELayer(driftx_layer, prior_driftw_layer, post_driftw_layer, diffusion_layer, x_dim, w_dim, setting):
    """brax/_impl/layers.py:164-248 (point estimate of W0: `priorw0_sigma` < 0, the toy's default)."""
    if setting.get("priorw0_sigma", -1.0) >= 0:
        raise NotImplementedError("Bayesian W0 (priorw0_sigma >= 0) needs layers.BayesianLayer, which the reference does not define")
    H, act = _mlp_spec(driftx_layer, x_dim)
    prior = _const_of(prior_driftw_layer, "prior_driftw_layer")
    if prior not in (("const", 0.0), ("affine", -1.0, 0.0)):
        raise NotImplementedError("prior drift must be Const(0.) or Affine(-1, 0.) (examples/jax/sdebnn_toy1d.py:45-47)")
    prior_ou = prior[0] == "affine"
    diff = _const_of(diffusion_layer, "diffusion_layer")
    if diff[0] != "const":
        raise NotImplementedError("diffusion must be DiffEqWrappertx(Const(c))")
    sigma = diff[1]
    psp = getattr(post_driftw_layer, "spec", None) or {}
    diff_drift, driftw_net = False, post_driftw_layer
    if psp.get("kind") == "additive":
        diff_drift = True
        if (psp["first"] or {}) != (getattr(prior_driftw_layer, "spec", None) or {}):
            raise NotImplementedError("Additive(prior, driftw): the first layer must be the prior drift layer (sdebnn_toy1d.py:82-83)")
    driftw = True
    if psp.get("kind") == "wrap" and (psp.get("inner") or {}).get("kind") == "const" and psp["inner"]["const"] == 0.0:
        driftw = False  # DiffEqWrappertx(Const(0.)), sdebnn_toy1d.py:64-65
    else:
        Hw, actw = _mlp_spec(psp["layers"][1] if diff_drift else post_driftw_layer, w_dim)
        if (Hw, actw) != (H, act):
            raise NotImplementedError("driftx and driftw must share hidden width and activation (sdebnn_toy1d.py:40-83)")
    driftw_scale = float(setting.get("driftw_scale", 1.0))

    def init_fun(rng, input_shape):
        k1, k2, k3, k4 = split(rng, 4)
        _, W0 = driftx_layer[0](k1, input_shape)
        _, W_driftwt = post_driftw_layer[0](k2, (-1, w_dim))
        _, W_difft = diffusion_layer[0](k3, (-1, w_dim))
        _, W_priorwt = prior_driftw_layer[0](k4, (-1, w_dim))
        W_driftwt = _tree_map(lambda w: w * driftw_scale, W_driftwt)  # layers.py:189-190
        return tuple(input_shape), (W0, W_driftwt, W_difft, W_priorwt, None)

    def apply_fun(params, inputs, rng, entire=False, num_samples=None, **kwargs):
        W0, W_driftwt, _W_diffusion, _W_priorwt, W0_post = params
        if W0_post is not None:
            raise NotImplementedError("Bayesian W0")
        if not inputs.is_cuda:
            raise _lib.BsdeError("SDELayer needs CUDA tensors (there is no CPU fallback)")
        dev = inputs.device
        x_lays, x_betas = _csl_leaves(W0)
        # flat_W0 = ravel_pytree(W0): per layer W, b, w_t, w_tb, b_t, then the (beta,) leaf of the activation after it
        pieces = []
        for i, lay in enumerate(x_lays):
            pieces += [t.reshape(-1) for t in lay]
            if i < len(x_betas):
                pieces.append(x_betas[i])
        flat_W0 = torch.cat(pieces).to(dev, torch.float32)
        P = flat_W0.numel()
        if P != w_dim:
            raise ValueError(f"w_dim = {w_dim} but the driftx parameters flatten to {P} entries")
        fw_flat = []
        if driftw:
            tree = W_driftwt[1] if diff_drift else W_driftwt
            lays, betas = _csl_leaves(tree)
            fw_flat = [t for lay in lays for t in lay] + betas
        B, D = int(inputs.shape[0]), int(inputs.shape[-1])
        nt = 90 if entire else 20
        ts = torch.linspace(0.0, 1.0, nt, device=dev)   # layers.py:235,242
        eps, seed, sid, S = None, 0, 0, num_samples or 1
        if isinstance(rng, torch.Tensor):
            eps = rng.to(dev, torch.float32)
            eps = eps.unsqueeze(0) if eps.dim() == 2 else eps
            S = eps.shape[0]
            if tuple(eps.shape) != (S, nt - 1, P):
                raise ValueError(f"unit normals must have shape (S, {nt - 1}, {P}), got {tuple(eps.shape)}")
            eps = eps.contiguous()
        else:
            seed, sid = (int(rng[0]), int(rng[1])) if isinstance(rng, (tuple, list)) else (int(rng), 0)

        def make_desc(prior_solve):
            def desc():
                d = _lib.SDELayerDesc()
                d.S, d.B, d.D, d.H, d.nt, d.act = S, B, D, H, nt, _lib.ACT[act]
                d.driftw, d.diffw = int(driftw), int(bool(setting.get("diffw", True)) and sigma != 0.0)
                d.prior_ou, d.diff_drift, d.stop_grad = int(prior_ou), int(diff_drift), int(bool(setting.get("stop_grad", False)))
                d.prior_solve, d.w0_per_sample, d.sigma, d.P = int(prior_solve), 0, sigma, P
                d.seed, d.stream_id, d.sample_offset = seed & 0xFFFFFFFFFFFFFFFF, sid & 0xFFFFFFFF, 0
                return d
            return desc

        x0 = inputs.reshape(B, D)
        post = _SDELayerSolve.apply(dict(desc=make_desc(False), act=act), eps, ts, x0, flat_W0, *fw_flat)
        # the prior solve reuses the rng (layers.py:238,244): same normals; it has no driftw parameters
        prior_tr = _SDELayerSolve.apply(dict(desc=make_desc(True), act=act), eps, ts, x0, flat_W0)
        BD = B * D

        def unravel(y):  # state_unravel_fn: (w, x, kl); returned in the order x, w, kl (layers.py:239-246)
            return y[..., P:P + BD].reshape(y.shape[:-1] + tuple(inputs.shape)), y[..., :P], y[..., P + BD]

        sel = (lambda tr: tr[:, 1:]) if entire else (lambda tr: tr[:, -1])
        out = unravel(sel(post)) + unravel(sel(prior_tr))
        if num_samples is None and not (isinstance(rng, torch.Tensor) and rng.dim() == 3):
            out = tuple(o[0] for o in out)
        return out

    return _with_spec(init_fun, apply_fun, kind="sdelayer", H=H, act=act)


def _tree_map(fn, tree):
    if isinstance(tree, torch.Tensor):
        return fn(tree)
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(fn, t) for t in tree)
    return tree


def mlp(rng, x_dim, hidden_dim, activn, setting):
    """examples/jax/sdebnn_toy1d.py:32-95.  Returns ((init_fun, apply_fun), flat_w_dim) of AugmentedLayer -> SDELayer."""
    prior_driftw_layer = DiffEqWrappertx(Const(0.0))
    if setting["prior_driftw"] == "ou":
        prior_driftw_layer = DiffEqWrappertx(Affine(-1, 0.0))
    augmented_x_dim = x_dim + setting["aug_dim"]
    activ_fn = get_activn(activn)
    if activn == "swish":
        activ_fn = activ_fn(hidden_dim)

    def make_driftx_layer(input_shape):
        return serial_tx(ConcatSquashLineartx(hidden_dim), DiffEqWrappertx(activ_fn), ConcatSquashLineartx(hidden_dim),
                         DiffEqWrappertx(activ_fn), ConcatSquashLineartx(int(input_shape[-1])))

    driftx_layer = shape_dependent(make_driftx_layer)
    _, w_init = driftx_layer[0](rng, (-1, augmented_x_dim))
    flat_w_dim = sum(t.numel() for lay in w_init for t in lay)
    if not setting["driftw"]:
        driftw_layer = DiffEqWrappertx(Const(0.0))
    else:
        def make_driftw_layer(input_shape):
            final = ConcatSquashLineartx(int(input_shape[-1]), W_init="zeros", b_init="zeros")
            if setting["stop_grad"] and setting["diff_drift"]:
                final = ConcatSquashLineartx(int(input_shape[-1]))
            return serial_tx(ConcatSquashLineartx(hidden_dim), DiffEqWrappertx(activ_fn), ConcatSquashLineartx(hidden_dim),
                             DiffEqWrappertx(activ_fn), final)

        driftw_layer = shape_dependent(make_driftw_layer)
        if setting["diff_drift"]:
            driftw_layer = Additive(prior_driftw_layer, driftw_layer)
    if not setting["diffw"]:
        diffusion_layer = DiffEqWrappertx(Const(0.0))
    else:
        diffusion_layer = DiffEqWrappertx(Const(setting["diff_const"]))
    sde = SDELayer(driftx_layer, prior_driftw_layer, driftw_layer, diffusion_layer, augmented_x_dim, flat_w_dim, setting)
    aug = AugmentedLayer(setting["aug_dim"])

    def init_fun(rng_, input_shape):
        k1, k2 = split(rng_, 2)
        shape, _ = aug[0](k1, input_shape)
        out_shape, p = sde[0](k2, shape)
        return out_shape, [(), p]

    def apply_fun(params, inputs, rng=None, **kwargs):
        return sde[1](params[1], aug[1](params[0], inputs), rng, **kwargs)

    return Layer(init_fun, apply_fun), flat_w_dim


def mse_loss(preds, targets, noise_std=3):   # sdebnn_toy1d.py:97-101
    return ((targets - preds) ** 2).sum(-1).mean() / (2 * noise_std ** 2)


def laplace_loss(preds, targets, noise_std=3):   # sdebnn_toy1d.py:103-107
    return (targets - preds).abs().sum(-1).mean()


def elbo(params, batch, apply_fn, rng, kl_scale, loss_fn=laplace_loss, num_samples=10, noise_std=0.1):
    """sdebnn_toy1d.py:118-131: the vmap over rngs is the kernel's sample axis."""
    inputs, targets = batch
    preds, _, kl, _, _, _ = apply_fn(params, inputs, rng=rng, entire=False, num_samples=num_samples)
    preds = preds[..., -1][..., None]
    return loss_fn(preds, targets[None], noise_std) + kl.mean() * kl_scale

"""GPU parity tests: the CUDA path (through the C ABI / the reference-shaped Python API) against
the CPU oracle on the same seeded inputs and the same Brownian increments.

Tolerances (fp32 path, `math="fp32"`; the reference states none -- SURVEY.md 8.3):
  * single conv / MLP evaluation vs fp64 oracle:      rtol 1e-5, atol 1e-5
  * w trajectory after all scan iterations:          rtol 1e-5, atol 1e-6
  * sum-KL, loss:                                    rtol 1e-4
  * gradients vs fp64 oracle autograd:               rtol 1e-3 (atol scaled to the tensor's max)
"""
import os
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import jax_path as jp

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _close(a, b, rtol, atol, what=""):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{a.numel()} mismatches, max err {err.max():.3e} " \
                          f"(ref max {b.abs().max():.3e})"


def _gclose(a, b, rtol=1e-3, what=""):
    b64 = b.detach().double().cpu()
    _close(a, b, rtol, rtol * 1e-2 * float(b64.abs().max()) + 1e-12, what)


def _mark(msg):
    """Progress marker (BSDE_TEST_MARKS=<file>): written and flushed BEFORE a launch, so that a hung kernel can be named."""
    path = os.environ.get("BSDE_TEST_MARKS")
    if path:
        with open(path, "a") as f:
            f.write(msg + "\n")
            f.flush()
            os.fsync(f.fileno())


def _conv_struct(L):
    from bayesde_b200 import _lib
    c = _lib.ConvLayer()
    c.cin, c.cout, c.ksize = L["cin"], L["cout"], L["k"]
    c.hin, c.win, c.hout, c.wout = L["hin"], L["win"], L["hout"], L["wout"]
    c.sn, c.kn, c.off, c.sd = L["sn"], L["kn"], L["off"], L["sd"]
    c.has_time, c.time_first = 1, 0
    c.w_off, c.b_off = L["w_off"], L["b_off"]
    c.w_tap_stride, c.w_k_stride, c.w_n_stride = (L["cin"] + 1) * L["cout"], L["cout"], 1
    return c


@pytest.mark.parametrize("C_in,fx_dim,H", [(5, 64, 8), (3, 8, 6), (20, 64, 16), (80, 16, 4)])
def test_tconv_layers_forward_and_wgrad(built_lib, C_in, fx_dim, H):
    """Every conv geometry of build_fx block_type 0 (stride 1 SAME, stride 2 SAME, transposed, Cout=C)
    against the oracle's ConcatConv2D, plus the weight gradient against fp64 autograd."""
    from bayesde_b200 import _lib
    from bayesde_b200.arch import FxSpec
    L = _lib.lib()
    S, B, t = 2, 3, 0.37
    fx = FxSpec(0, (1, H, H, C_in), fx_dim, "softplus")
    layout, P = jp.fx_layout(0, C_in, fx_dim)
    assert P == fx.P
    g = torch.Generator().manual_seed(0)
    w = torch.randn(S, P, generator=g, dtype=torch.float64) * 0.2
    conv_entries = [e for e in layout if e["kind"] == "conv"]
    for li, (Ld, e) in enumerate(zip(fx.layers, conv_entries)):
        x = torch.randn(S, B, Ld["hin"], Ld["win"], Ld["cin"], generator=g, dtype=torch.float64)
        n = e["k"] ** 2 * (e["cin"] + 1) * e["cout"]
        refs = []
        for s in range(S):
            ws = w[s].clone().requires_grad_(True)
            refs.append((jp.concat_conv2d(x[s], t, ws[e["w_off"]:e["w_off"] + n].reshape(e["w_shape"]),
                                          ws[e["b_off"]:e["b_off"] + e["cout"]], e["stride"], e["transpose"]), ws))
        ref = torch.stack([r[0] for r in refs])
        assert ref.shape[2] == Ld["hout"] and ref.shape[4] == Ld["cout"]
        cs = _conv_struct(Ld)
        xd = x.float().to(DEV).contiguous()
        wd = w.float().to(DEV).contiguous()
        out = torch.empty(ref.shape, device=DEV)
        st = L.bsde_tconv_fwd(C.byref(cs), S, B, _lib.ptr(xd), _lib.ptr(wd), P, t, 0, _lib.EPI_BIAS, 1.0, None,
                              _lib.ptr(out), 0, None, 0, _lib.stream_ptr())
        _lib.check(st, "tconv_fwd")
        _close(out, ref, 1e-5, 1e-5, f"conv layer {li} forward")
        # weight gradient
        dy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
        gref = []
        for s in range(S):
            (gw,) = torch.autograd.grad((refs[s][0] * dy[s]).sum(), refs[s][1])
            gref.append(gw)
        gref = torch.stack(gref)
        gw = torch.zeros(S, P, device=DEV)
        nb = L.bsde_tconv_wgrad_workspace_bytes(C.byref(cs), S, B)
        ws_buf = torch.empty(nb, dtype=torch.uint8, device=DEV)
        st = L.bsde_tconv_wgrad(C.byref(cs), S, B, _lib.ptr(xd), _lib.ptr(dy.float().to(DEV).contiguous()), t, 0.5,
                                _lib.ptr(gw), P, 0, C.c_void_p(ws_buf.data_ptr()), nb, _lib.stream_ptr())
        _lib.check(st, "tconv_wgrad")
        _gclose(gw, 0.5 * gref, 1e-4, f"conv layer {li} wgrad")


def _product_fw_params(fw_layers, device):
    """oracle list of 5-tuples -> reference-shaped pytree ((), [lay, (), ..., lay]) on device."""
    tree = []
    for i, lay in enumerate(fw_layers):
        tree.append(tuple(t.float().to(device).contiguous().requires_grad_(True) for t in lay[:5]))
        if i < len(fw_layers) - 1:   # the activation's parameters: () or the trainable swish's (beta,)
            tree.append(tuple(t.float().to(device).contiguous().requires_grad_(True) for t in lay[5:]))
    return ((), tree)


def _block_case(S=2, B=2, H=8, C_in=5, fx_dim=8, nsteps=6, sigma=0.2, stl=False, fw_dims=(2, 16, 2),
                time_grid="reference", seed=0, fx_actfn="softplus", fw_actfn="softplus"):
    g = torch.Generator().manual_seed(seed)
    blk = jp.SDEBNNBlock((B, H, H, C_in), 0, fx_dim, fx_actfn, fw_actfn, sigma, stl, nsteps,
                         time_grid_mode=time_grid)
    w0 = jp.init_fx(blk.layout, blk.P, g, torch.float64)
    fw = jp.init_fw(blk.P, fw_dims, g, torch.float64, last_std=0.05, actfn=fw_actfn, beta_jitter=0.2)
    if fx_actfn == "swish":   # distinct beta per channel (the init value is the constant 0.5)
        for e in blk.layout:
            if e["kind"] == "beta":
                w0[e["off"]:e["off"] + e["n"]] += 0.3 * torch.randn(e["n"], generator=g, dtype=torch.float64)
    x = torch.randn(S, B, H, H, C_in, generator=g, dtype=torch.float64)
    dW = torch.stack([jp.sample_increments(blk.P, nsteps, g, torch.float64, time_grid) for _ in range(S)])
    return blk, w0, fw, x, dW


def _oracle_block(blk, w0, fw, x, dW, gx_seed=1, kl_w=1e-3):
    """Per-sample oracle solves (the reference's vmap) + a scalar objective for gradients."""
    w0 = w0.clone().requires_grad_(True)
    fw = [tuple(t.clone().requires_grad_(True) for t in lay) for lay in fw]
    x = x.clone().requires_grad_(True)
    xs, kls, ws = [], [], []
    for s in range(x.shape[0]):
        xo, kl, wtraj = blk.apply((w0, fw), x[s], dW[s])
        xs.append(xo), kls.append(kl), ws.append(wtraj)
    xo, kl, wt = torch.stack(xs), torch.stack(kls), torch.stack(ws)
    g = torch.Generator().manual_seed(gx_seed)
    cot = torch.randn(xo.shape, generator=g, dtype=torch.float64)
    obj = (xo * cot).sum() + kl_w * kl.sum()
    grads = torch.autograd.grad(obj, [x, w0] + [t for lay in fw for t in lay[:5]] + [lay[5] for lay in fw if len(lay) > 5])
    return xo, kl, wt, cot, grads


@pytest.mark.parametrize("kw", [dict(), dict(stl=True), dict(time_grid="uniform", nsteps=5),
                                dict(S=1, B=3, H=4, C_in=20, fx_dim=16, fw_dims=(1, 8, 1)),
                                dict(fx_actfn="tanh", fw_actfn="tanh", fw_dims=(2,)),
                                dict(fx_actfn="elu", fw_actfn="elu", fw_dims=(2, 8, 8, 2), sigma=1e-2),
                                dict(fx_actfn="rbf", fw_actfn="rbf", fw_dims=(2, 8, 2)),
                                dict(fx_actfn="rbf", fw_actfn="softplus", stl=True, S=1, B=3, H=4, C_in=20, fx_dim=16),
                                # trainable swish (arch.py:14-29): beta of f_x lives in w(t), beta of f_w is a parameter vector
                                dict(fx_actfn="swish", fw_actfn="swish", fw_dims=(2, 8, 2)),
                                dict(fx_actfn="swish", fw_actfn="softplus", stl=True, S=1, B=3, H=4, C_in=20, fx_dim=16),
                                dict(fx_actfn="softplus", fw_actfn="swish", fw_dims=(2,)),
                                dict(fx_actfn="swish", fw_actfn="swish", fw_dims=(1, 8, 8, 1), remat=True)])
def test_sdebnn_block_forward_backward(built_lib, kw):
    """One SDEBNN block through the reference-shaped API vs the oracle: x_T, per-sample KL, the whole
    w trajectory (full_output), and gradients w.r.t. x, w0 and every f_w tensor."""
    from bayesde_b200 import arch, brax
    kw = dict(kw)
    remat = kw.pop("remat", False)
    blk, w0, fw, x, dW = _block_case(**kw)
    stl = kw.get("stl", False)
    fw_layer = arch.MLP(kw.get("fw_dims", (2, 16, 2)), actfn=kw.get("fw_actfn", "softplus"), xt=True, ou_dw=True)
    layer = brax.SDEBNN(0, blk.layout[0]["cout"], kw.get("fx_actfn", "softplus"), fw_layer, diff_coef=blk.diff_coef,
                        stl=stl, xt=True, nsteps=blk.nsteps, time_grid=blk.time_grid_mode, remat=remat)
    # the oracle's stale vector must be the product's (fx.init(PRNGKey(0)) stand-in), brax.py:38-41
    _, (w_stale_like, _, _) = layer.init(0, tuple(x.shape[1:]))
    blk.w_stale = w_stale_like.double()
    xo_r, kl_r, wt_r, cot, grads_r = _oracle_block(blk, w0, fw, x, dW)

    w0d = w0.float().to(DEV).requires_grad_(True)
    fwd = _product_fw_params(fw, DEV)
    xd = x.float().to(DEV).requires_grad_(True)
    xo, kl, info = layer.apply((w0d, (), fwd), xd, dW.float().to(DEV), full_output=True)
    _close(xo, xo_r, 1e-4, 1e-5, "x_T")
    _close(info["sdebnn_w"], wt_r, 1e-5, 1e-6, "w trajectory")
    _close(kl, kl_r, 1e-4, 1e-6, "sum KL")
    obj = (xo * cot.float().to(DEV)).sum() + 1e-3 * kl.sum()
    leaves = [xd, w0d] + [t for lay in arch.fw_layers(fwd) for t in lay] + arch.fw_betas(fwd)
    grads = torch.autograd.grad(obj, leaves)
    names = (["x", "w0"] + [f"fw[{i}].{n}" for i in range(len(fw)) for n in ("W", "b", "w_t", "w_tb", "b_t")]
             + [f"fw[{i}].beta" for i in range(len(arch.fw_betas(fwd)))])
    assert len(grads) == len(grads_r)
    # rbf: the data-gradient epilogues rebuild the pre-activation from the stored activation (sign in the last mantissa bit,
    # |x| = sqrt(-ln a), which loses relative precision near a = 1): 5e-3 instead of 1e-3 (measured worst entry 1.5e-5 of max)
    gt = 5e-3 if kw.get("fx_actfn") == "rbf" else 1e-3
    for n, a, b in zip(names, grads, grads_r):
        _gclose(a, b, gt, f"grad {n}")


@pytest.mark.parametrize("stl", [False, True])
def test_structured_route_returns_ys_and_callable_dynamics(built_lib, stl):
    """Drop-in boundary (jaxsde/jaxsde/sdeint.py:138): `sdeint_ito_fixed_grid(dyn.f_aug, dyn.g_aug, y0, ts, rng, fw_params)`
    returns ys (T, D) also on the fused route -- x rows, w rows and kl rows of EVERY scan time against the oracle's generic
    solve of the flat state, `ys[-1][:x_dim]` / `ys[-1][x_dim+P:].sum()` differentiable as brax.py:117-119 consumes them --
    and `f_aug` / `g_aug` are callable (values) like the reference's closures (brax.py:49-78)."""
    from bayesde_b200 import arch, brax, sdeint_ito_fixed_grid
    blk, w0, fw, x, dW = _block_case(S=1, B=2, H=6, C_in=4, fx_dim=8, nsteps=5, stl=stl)
    fw_layer = arch.MLP((2, 16, 2), xt=True, ou_dw=True)
    layer = brax.SDEBNN(0, 8, "softplus", fw_layer, diff_coef=blk.diff_coef, stl=stl, xt=True, nsteps=blk.nsteps)
    _, (w_stale_like, _, _) = layer.init(0, tuple(x.shape[1:]))
    blk.w_stale = w_stale_like.double()
    P, x_dim = blk.P, blk.x_dim
    # oracle: the generic solver on the flat state [x, w, kl] (brax.py:110-113)
    w0r = w0.clone().requires_grad_(True)
    y0_r = torch.cat([x[0].reshape(-1), w0r, torch.zeros(P, dtype=torch.float64)])
    inc = torch.cat([dW[0].new_zeros(dW.shape[1], x_dim), dW[0], dW[0]], dim=1)
    ts = torch.linspace(0.0, 1.0, blk.nsteps)
    ys_r = jp.sdeint_ito_fixed_grid(blk.f_aug, blk.g_aug, y0_r, ts.double(), inc, fw, method="euler_maruyama", rep=P if stl else 0)
    obj_r = ys_r[-1][:x_dim].pow(2).sum() + 1e-3 * ys_r[-1][x_dim + P:].sum()
    (g_r,) = torch.autograd.grad(obj_r, w0r)

    dyn = brax._AugDynamics(arch.FxSpec(0, (1,) + tuple(x.shape[2:]), 8, "softplus"), fw_layer.spec, x.shape, blk.diff_coef, stl,
                            True, "fp32", w_stale_like.to(DEV), 0)
    w0d = w0.float().to(DEV).requires_grad_(True)
    y0 = torch.cat([x[0].reshape(-1).float().to(DEV), w0d, torch.zeros(P, device=DEV)])
    fwd = _product_fw_params(fw, DEV)
    ys = sdeint_ito_fixed_grid(dyn.f_aug, dyn.g_aug, y0, ts, dW.float().to(DEV), fwd, method="euler_maruyama", rep=P if stl else 0)
    assert ys.shape == ys_r.shape == (blk.nsteps, x_dim + 2 * P)
    _close(ys[:, :x_dim], ys_r[:, :x_dim], 1e-4, 1e-5, "ys x rows")
    _close(ys[:, x_dim:x_dim + P], ys_r[:, x_dim:x_dim + P], 1e-5, 1e-6, "ys w rows")
    _close(ys[:, x_dim + P:], ys_r[:, x_dim + P:], 1e-4, 1e-4 * float(ys_r[:, x_dim + P:].abs().max()), "ys kl rows")
    obj = ys[-1][:x_dim].pow(2).sum() + 1e-3 * ys[-1][x_dim + P:].sum()
    (g_d,) = torch.autograd.grad(obj, w0d)
    _gclose(g_d, g_r, 1e-3, "d/dw0 through ys[-1]")
    # the closures themselves, at an interior state and time
    yk, t = ys_r[2].detach(), 0.37
    f_r, g_ref = blk.f_aug(yk, torch.tensor(t, dtype=torch.float64), fw), blk.g_aug(yk, torch.tensor(t, dtype=torch.float64), fw)
    f_d = dyn.f_aug(yk.float().to(DEV), t, fwd)
    g_d2 = dyn.g_aug(yk.float().to(DEV), t, fwd)
    for name, sl in (("dx", slice(0, x_dim)), ("dw", slice(x_dim, x_dim + P)), ("dkl", slice(x_dim + P, None))):
        _tc_close(f_d[sl], f_r[sl], 2e-5, f"f_aug {name}")
        _tc_close(g_d2[sl], g_ref[sl], 2e-5, f"g_aug {name}")


def test_meanfield_wraps_a_stochastic_layer(built_lib):
    """`--meanfield_sdebnn` (examples/jax/sdebnn_classification.py:180-194): MeanField(SDEBNN(..., stax_api=True)).  The wrapper
    must hand the wrapped layer `rng=next_rng` (brax.py:150,165) -- SDEBNN.apply_fun needs it for its Brownian path."""
    from bayesde_b200 import arch, brax
    inner = brax.SDEBNN(0, 8, "softplus", arch.MLP((2, 8, 2), xt=True, ou_dw=True), diff_coef=0.1, xt=True, nsteps=4,
                        stax_api=True)
    init, apply = brax.MeanField(inner)
    _, params = init(3, (2, 6, 6, 4))
    params = brax.tree_map(lambda t: t.to(DEV), params)
    x = torch.randn(2, 6, 6, 4, device=DEV)
    out1, kl1 = apply(params, x, rng=5)
    out2, kl2 = apply(params, x, rng=5)
    out3, _ = apply(params, x, rng=6)
    assert out1.shape == x.shape and torch.isfinite(out1).all() and kl1.dim() == 0 and float(kl1) > 0
    assert torch.equal(out1, out2) and torch.equal(kl1, kl2) and not torch.equal(out1, out3)


def test_block_types_1_and_2(built_lib):
    """arch.py:189-205: block_type 1 (3x3, 3x3, 3x3 as shipped: the `(1, 1)` of arch.py:193 lands in ConcatConv2D's unused W_init
    slot, diffeq_layers.py:124, so kernel stays 3) and 2 (3x3, 3x3)."""
    from bayesde_b200 import arch, brax
    # P of block type 1 with a 3x3 middle conv: 9*(C+1)*F + F + 9*(F+1)*F + F + 9*(F+1)*C + C
    assert arch.FxSpec(1, (1, 6, 6, 4), 8, "softplus").P == 9 * 5 * 8 + 8 + 9 * 9 * 8 + 8 + 9 * 9 * 4 + 4
    assert jp.fx_layout(1, 4, 8)[1] == arch.FxSpec(1, (1, 6, 6, 4), 8, "softplus").P
    for bt in (1, 2):
        g = torch.Generator().manual_seed(3)
        S, B, H, Cc, nsteps = 1, 2, 6, 4, 4
        blk = jp.SDEBNNBlock((B, H, H, Cc), bt, 8, "softplus", "softplus", 0.1, False, nsteps)
        w0 = jp.init_fx(blk.layout, blk.P, g, torch.float64)
        fw = jp.init_fw(blk.P, (2, 8, 2), g, torch.float64, last_std=0.05)
        x = torch.randn(S, B, H, H, Cc, generator=g, dtype=torch.float64)
        dW = torch.stack([jp.sample_increments(blk.P, nsteps, g, torch.float64)])
        xo_r, kl_r, wt_r, cot, grads_r = _oracle_block(blk, w0, fw, x, dW)
        layer = brax.SDEBNN(bt, 8, "softplus", arch.MLP((2, 8, 2), xt=True, ou_dw=True), diff_coef=0.1, xt=True,
                            nsteps=nsteps)
        w0d = w0.float().to(DEV).requires_grad_(True)
        fwd = _product_fw_params(fw, DEV)
        xd = x.float().to(DEV).requires_grad_(True)
        xo, kl = layer.apply((w0d, (), fwd), xd, dW.float().to(DEV))
        _close(xo, xo_r, 1e-4, 1e-5, f"block_type {bt} x_T")
        _close(kl, kl_r, 1e-4, 1e-6, f"block_type {bt} KL")
        grads = torch.autograd.grad((xo * cot.float().to(DEV)).sum() + 1e-3 * kl.sum(), [xd, w0d])
        _gclose(grads[0], grads_r[0], 1e-3, "grad x")
        _gclose(grads[1], grads_r[1], 1e-3, "grad w0")


def test_methods_coincide_for_sdebnn(built_lib):
    """g_aug is state-independent (quirk Q2) => milstein == euler_maruyama == euler_heun; the oracle's
    generic solver (autograd Jacobian) confirms it, and the fused path accepts all three names."""
    from bayesde_b200 import arch, brax
    blk, w0, fw, x, dW = _block_case(S=1, B=1, H=4, C_in=3, fx_dim=4, nsteps=4)
    outs = {}
    for m in ("euler_maruyama", "milstein", "euler_heun"):
        blk.method = m
        xo, kl, _ = blk.apply((w0, fw), x[0], dW[0])
        outs[m] = (xo, kl)
        layer = brax.SDEBNN(0, 4, "softplus", arch.MLP((2, 16, 2), xt=True, ou_dw=True), diff_coef=0.2, xt=True,
                            nsteps=4, method=m)
        xo_d, kl_d = layer.apply((w0.float().to(DEV), (), _product_fw_params(fw, DEV)), x.float().to(DEV),
                                 dW.float().to(DEV))
        _close(xo_d[0], xo, 1e-4, 1e-5, m)
        _close(kl_d[0], kl, 1e-4, 1e-6, m)
    _close(outs["milstein"][0], outs["euler_maruyama"][0], 1e-12, 1e-12)
    _close(outs["euler_heun"][1], outs["euler_maruyama"][1], 1e-12, 1e-12)
    with pytest.raises(ValueError):
        brax.SDEBNN(0, 4, "softplus", arch.MLP((2, 16, 2), xt=True), xt=True, nsteps=4, method="rk4").apply(
            (w0.float().to(DEV), (), _product_fw_params(fw, DEV)), x.float().to(DEV), dW.float().to(DEV))


def test_generic_route_vs_oracle(built_lib):
    """Arbitrary callables f, g (state-dependent diagonal g) through the generic route, all methods,
    against the oracle's restatement of sdeint.py on the same increments; gradient w.r.t. args."""
    from bayesde_b200 import sdeint_ito_fixed_grid
    D, T = 257, 9
    g_ = torch.Generator().manual_seed(5)
    y0 = torch.randn(D, generator=g_, dtype=torch.float64)
    A = torch.randn(D, generator=g_, dtype=torch.float64) * 0.5
    inc = torch.randn(T - 1, D, generator=g_, dtype=torch.float64) * math.sqrt(1.0 / (T - 1))
    f = lambda y, t, a: -a * y + torch.sin(t + y)
    g = lambda y, t, a: 0.3 * torch.cos(y) + 0.1 * a
    ts = torch.linspace(0.0, 1.0, T)
    for method in ("euler_maruyama", "milstein", "euler_heun"):
        for tg in ("reference", "uniform"):
            a_r = A.clone().requires_grad_(True)
            ys_r = jp.sdeint_ito_fixed_grid(f, g, y0, ts.double(), inc, a_r, method=method, time_grid_mode=tg)
            (gr,) = torch.autograd.grad(ys_r[-1].pow(2).sum(), a_r)
            a_d = A.float().to(DEV).requires_grad_(True)
            ys = sdeint_ito_fixed_grid(f, g, y0.float().to(DEV), ts, inc.float().to(DEV), a_d, method=method,
                                       time_grid=tg)
            assert ys.shape == (T, D)
            _close(ys, ys_r, 1e-5, 1e-5, f"{method}/{tg}")
            (gd,) = torch.autograd.grad(ys[-1].pow(2).sum(), a_d)
            _gclose(gd, gr, 1e-3, f"{method}/{tg} grad")
    with pytest.raises(ValueError, match="Unknown method"):
        sdeint_ito_fixed_grid(f, g, y0.float().to(DEV), ts, 0, A.float().to(DEV), method="heun")


def test_philox_increments_law_and_in_kernel_stream(built_lib):
    """(i) mean 0 / variance dt_k (analogue of jaxsde/tests/test_brownian.py:15-45);
    (ii) the in-kernel stream of K1 equals the materialised one: solving with rng=seed matches solving
    with the materialised increments passed explicitly; (iii) rep tying (brownian.py:69-74)."""
    from bayesde_b200 import arch, brax, sdeint_ito_fixed_grid
    from bayesde_b200.sdeint import philox_increments
    dtk = np.array([0.05, 0.0, 0.2, 0.05], dtype=np.float32)
    z = philox_increments(1234, 7, 3, dtk, 200_000, DEV)
    assert z.shape == (3, 4, 200_000)
    for k, h in enumerate(dtk):
        v = z[:, k].double()
        if h == 0:
            assert v.abs().max() == 0
            continue
        assert abs(v.mean().item()) < 4 * math.sqrt(h / v.numel())
        assert abs(v.var().item() / h - 1) < 0.01
        assert abs((v ** 4).mean().item() / (3 * h * h) - 1) < 0.03  # Gaussian kurtosis
    assert abs(torch.corrcoef(torch.stack([z[0, 0], z[1, 0]]))[0, 1].item()) < 0.01
    assert not torch.equal(philox_increments(1234, 8, 1, dtk, 1000, DEV), z[:1, :, :1000])

    blk, w0, fw, x, _ = _block_case(S=3, B=1, H=4, C_in=3, fx_dim=4, nsteps=5)
    layer = brax.SDEBNN(0, 4, "softplus", arch.MLP((2, 16, 2), xt=True, ou_dw=True), diff_coef=0.2, xt=True,
                        nsteps=5, stream_id=11)
    params = (w0.float().to(DEV), (), _product_fw_params(fw, DEV))
    from bayesde_b200._fused import reference_time_grid
    _, dtk5 = reference_time_grid(np.linspace(0, 1, 5, dtype=np.float32))
    dW = philox_increments(99, 11, 3, dtk5, blk.P, DEV)
    xa, kla = layer.apply(params, x.float().to(DEV), 99)
    xb, klb = layer.apply(params, x.float().to(DEV), dW)
    assert torch.equal(xa, xb) and torch.equal(kla, klb)

    ys = sdeint_ito_fixed_grid(lambda y, t, a: 0 * y, lambda y, t, a: torch.ones_like(y), torch.zeros(12, device=DEV),
                               torch.linspace(0, 1, 5), 3, rep=4, time_grid="uniform")
    assert torch.equal(ys[:, 8:], ys[:, 4:8]) and not torch.equal(ys[:, 0:4], ys[:, 4:8])


def test_classifier_loss_and_grads(built_lib):
    """examples/jax/sdebnn_classification.py model (Augment -> SDEBNN blocks -> squeeze -> ... -> mean-field
    Dense head), loss = nll + kl*kl_coef/train_size averaged over S samples, gradients of every parameter."""
    from bayesde_b200.classification import SDEBNNClassifier
    S, B = 2, 2
    kw = dict(input_size=(8, 8, 3), aug=1, nblocks=(1, 1), fx_dim=8, fw_dims=(2, 8, 2), diff_coef=0.1, nsteps=5)
    net_r = jp.SDEBNNClassifier(batch=B, **kw)
    params_r, head_r = net_r.init(0, torch.float64, fw_last_std=0.05)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, 8, 8, 3, generator=g, dtype=torch.float64)
    labels = torch.tensor([1, 7])
    onehot = torch.nn.functional.one_hot(labels, 10).double()
    dWs = [torch.stack([jp.sample_increments(b.P, 5, g, torch.float64) for _ in range(S)]) for b in net_r.blocks()]
    eps = torch.randn(S, net_r.feat * 10 + 10, generator=g, dtype=torch.float64)
    kl_scale = 1e-3 / 100.0

    leaves_r = []
    pr = []
    for (w0, fw) in params_r:
        w0 = w0.clone().requires_grad_(True)
        fw = [tuple(t.clone().requires_grad_(True) for t in lay) for lay in fw]
        pr.append((w0, fw))
        leaves_r += [w0] + [t for lay in fw for t in lay]
    hr = tuple(tuple(t.clone().requires_grad_(True) for t in pair) for pair in head_r)
    leaves_r += [hr[0][0], hr[0][1], hr[1][0], hr[1][1]]
    losses = []
    for s in range(S):
        loss, nll, kl, _ = net_r.loss(pr, hr, x, onehot, [d[s] for d in dWs], eps[s], kl_scale)
        losses.append(loss)
    loss_r = torch.stack(losses).mean()  # mean over samples of per-sample grads (script :231)
    grads_r = torch.autograd.grad(loss_r, leaves_r)

    net = SDEBNNClassifier(nsamples=S, device=DEV, **kw)
    net.load_oracle_params(params_r, head_r)
    loss, nll, kl = net.loss(x.float().to(DEV), labels.to(DEV), kl_scale,
                             rng=net.explicit_rng([d.float().to(DEV) for d in dWs], eps.float().to(DEV)))
    _close(loss, loss_r, 1e-4, 1e-6, "loss")
    loss.backward()
    leaves = net.oracle_ordered_parameters()
    assert len(leaves) == len(leaves_r)
    for i, (p, gr) in enumerate(zip(leaves, grads_r)):
        _gclose(p.grad, gr, 2e-3, f"param {i} grad")


# ------------------------------------------------------------------------------------------------
# tensor-core path (math="tc": tcgen05 kind::tf32 operands, fp32 accumulation).  TF32 keeps 10 mantissa
# bits, so the tolerance is stated relative to the tensor's scale: 2e-3 * max|ref| per conv evaluation.
# ------------------------------------------------------------------------------------------------
def _tc_close(a, b, rel, what):
    b64 = b.detach().double().cpu()
    _close(a, b, 0.0, rel * float(b64.abs().max()) + 1e-12, what)


@pytest.mark.parametrize("C_in,fx_dim,H,B", [(5, 64, 8, 3), (20, 64, 16, 2), (80, 64, 8, 5), (4, 16, 6, 3),
                                             # the taps-on-N conv kernel (64 / 32 -> <= 7 channels, BIAS and ACCUM epilogues) and the thin
                                             # weight-gradient kernel (64 <-> <= 7 channels): every channel count they serve, one-tile and
                                             # many-tile position grids, a full-size 32x32 image, and the counts just outside (fallbacks)
                                             (1, 64, 12, 2), (2, 64, 4, 1), (6, 64, 32, 1), (7, 64, 6, 3), (8, 64, 4, 2), (5, 32, 8, 2),
                                             (3, 64, 16, 9)])
def test_tconv_tc_layers_forward_and_dgrad(built_lib, C_in, fx_dim, H, B):
    """Every forward geometry and every data gradient on the tcgen05 path vs the fp64 oracle."""
    from bayesde_b200 import _lib
    from bayesde_b200.arch import FxSpec
    L = _lib.lib()
    S, t = 2, 0.61
    fx = FxSpec(0, (1, H, H, C_in), fx_dim, "softplus")
    layout, P = jp.fx_layout(0, C_in, fx_dim)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(S, P, generator=g, dtype=torch.float64) * 0.2
    conv_entries = [e for e in layout if e["kind"] == "conv"]
    for li, (Ld, e) in enumerate(zip(fx.layers, conv_entries)):
        x = torch.randn(S, B, Ld["hin"], Ld["win"], Ld["cin"], generator=g, dtype=torch.float64)
        n = e["k"] ** 2 * (e["cin"] + 1) * e["cout"]
        xr = x.clone().requires_grad_(True)
        ref = torch.stack([jp.concat_conv2d(xr[s], t, w[s, e["w_off"]:e["w_off"] + n].reshape(e["w_shape"]),
                                            w[s, e["b_off"]:e["b_off"] + e["cout"]], e["stride"], e["transpose"])
                           for s in range(S)])
        cs = _conv_struct(Ld)
        xd, wd = x.float().to(DEV).contiguous(), w.float().to(DEV).contiguous()
        out = torch.empty(ref.shape, device=DEV)
        nb = L.bsde_tconv_fwd_workspace_bytes(C.byref(cs), S, 1)
        assert nb > 0
        ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
        _mark(f"tc layers [{C_in}-{fx_dim}-{H}-{B}] layer {li} forward")
        st = L.bsde_tconv_fwd(C.byref(cs), S, B, _lib.ptr(xd), _lib.ptr(wd), P, t, 0, _lib.EPI_BIAS, 1.0, None,
                              _lib.ptr(out), 1, C.c_void_p(ws.data_ptr()), nb, _lib.stream_ptr())
        _lib.check(st, "tconv_fwd tc")
        torch.cuda.synchronize()
        _tc_close(out, ref, 2e-3, f"tc conv layer {li} forward")
        # data gradient through the same kernel: dgrad descriptor built like csrc/api.cu dgrad_layer()
        dy = torch.randn(ref.shape, generator=g, dtype=torch.float64)
        (gx_ref,) = torch.autograd.grad((ref * dy).sum(), xr)
        D = _lib.ConvLayer()
        for f_, _t in _lib.ConvLayer._fields_:
            setattr(D, f_, getattr(cs, f_))
        D.cin, D.cout = cs.cout, cs.cin
        D.hin, D.win, D.hout, D.wout = cs.hout, cs.wout, cs.hin, cs.win
        D.sn, D.kn, D.off, D.sd = cs.sd, -cs.kn, -cs.off, cs.sn
        D.has_time, D.time_first, D.b_off = 0, 0, -1
        D.w_k_stride, D.w_n_stride = cs.w_n_stride, cs.w_k_stride
        gx = torch.zeros(x.shape, device=DEV)
        nb = L.bsde_tconv_fwd_workspace_bytes(C.byref(D), S, 1)
        ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
        _mark(f"tc layers [{C_in}-{fx_dim}-{H}-{B}] layer {li} dgrad")
        st = L.bsde_tconv_fwd(C.byref(D), S, B, _lib.ptr(dy.float().to(DEV).contiguous()), _lib.ptr(wd), P, t, 0,
                              _lib.EPI_ACCUM, 1.0, _lib.ptr(gx), _lib.ptr(gx), 1, C.c_void_p(ws.data_ptr()), nb,
                              _lib.stream_ptr())
        _lib.check(st, "tconv dgrad tc")
        torch.cuda.synchronize()
        _tc_close(gx, gx_ref, 2e-3, f"tc conv layer {li} dgrad")
        # weight gradient on the tensor cores (direct form, or swapped form for the lhs-dilated layer)
        gref = []
        for s_ in range(S):
            ws_ = w[s_].clone().requires_grad_(True)
            o_ = jp.concat_conv2d(x[s_], t, ws_[e["w_off"]:e["w_off"] + n].reshape(e["w_shape"]),
                                  ws_[e["b_off"]:e["b_off"] + e["cout"]], e["stride"], e["transpose"])
            (g_,) = torch.autograd.grad((o_ * dy[s_]).sum(), ws_)
            gref.append(g_)
        gref = torch.stack(gref)
        gwd = torch.zeros(S, P, device=DEV)
        nb = L.bsde_tconv_wgrad_workspace_bytes(C.byref(cs), S, B)
        ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
        _mark(f"tc layers [{C_in}-{fx_dim}-{H}-{B}] layer {li} wgrad")
        st = L.bsde_tconv_wgrad(C.byref(cs), S, B, _lib.ptr(xd), _lib.ptr(dy.float().to(DEV).contiguous()), t, 0.5,
                                _lib.ptr(gwd), P, 1, C.c_void_p(ws.data_ptr()), nb, _lib.stream_ptr())
        _lib.check(st, "tconv_wgrad tc")
        torch.cuda.synchronize()
        lo, hi = e["w_off"], e["b_off"] + e["cout"]
        _tc_close(gwd[:, lo:hi], 0.5 * gref[:, lo:hi], 2e-3, f"tc conv layer {li} wgrad")


@pytest.mark.parametrize("kw", [dict(), dict(S=1, B=3, H=4, C_in=20, fx_dim=16, fw_dims=(1, 8, 1)),
                                dict(S=2, B=4, H=8, C_in=80, fx_dim=64, nsteps=4),
                                dict(fx_actfn="rbf", fw_actfn="rbf"), dict(fx_actfn="rbf", S=2, B=4, H=8, C_in=80, fx_dim=64, nsteps=4),
                                dict(fx_actfn="swish", fw_actfn="swish", fx_dim=64), dict(fx_actfn="swish", S=2, B=4, H=8, C_in=20, fx_dim=64, nsteps=4)])
def test_sdebnn_block_tc(built_lib, kw):
    """One SDEBNN block with math="tc" vs the fp64 oracle.  Stated tolerance (TF32 operands, 10-bit
    mantissa, nsteps-1 Euler steps): 5e-3 * max|ref| on x_T, 2e-2 * max|ref| on every gradient (the reverse
    pass chains 4 TF32 data-gradient convs per iteration); the w path and the KL never touch the tensor
    cores and keep the fp32 tolerances."""
    from bayesde_b200 import arch, brax
    blk, w0, fw, x, dW = _block_case(**kw)
    xo_r, kl_r, wt_r, cot, grads_r = _oracle_block(blk, w0, fw, x, dW)
    layer = brax.SDEBNN(0, blk.layout[0]["cout"], kw.get("fx_actfn", "softplus"),
                        arch.MLP(kw.get("fw_dims", (2, 16, 2)), actfn=kw.get("fw_actfn", "softplus"), xt=True, ou_dw=True),
                        diff_coef=blk.diff_coef, xt=True, nsteps=blk.nsteps, math="tc")
    w0d = w0.float().to(DEV).requires_grad_(True)
    fwd = _product_fw_params(fw, DEV)
    xd = x.float().to(DEV).requires_grad_(True)
    xo, kl, info = layer.apply((w0d, (), fwd), xd, dW.float().to(DEV), full_output=True)
    _tc_close(xo, xo_r, 5e-3, "tc x_T")
    _close(info["sdebnn_w"], wt_r, 1e-5, 1e-6, "w trajectory")
    _close(kl, kl_r, 1e-4, 1e-6, "sum KL")
    obj = (xo * cot.float().to(DEV)).sum() + 1e-3 * kl.sum()
    leaves = [xd, w0d] + [t for lay in arch.fw_layers(fwd) for t in lay] + arch.fw_betas(fwd)
    grads = torch.autograd.grad(obj, leaves)
    assert len(grads) == len(grads_r)
    for i, (a, b) in enumerate(zip(grads, grads_r)):
        # rbf: act' = -2 x a changes sign with the pre-activation, so the two-element gate gradients of the first f_w layer are sums
        # with heavy cancellation (|ref| ~ 1e-4) and TF32 rounding shows at 4e-2 of their own maximum: 8e-2
        _tc_close(a, b, 8e-2 if kw.get("fx_actfn") == "rbf" else 2e-2, f"tc grad {i}")


@pytest.mark.parametrize("math", ["fp32", "tc"])
def test_remat_modes_agree(built_lib, math):
    """remat=True (recompute conv intermediates from x_k, the jax.checkpoint analogue of brax.py:131-132)
    and remat=False (keep them) run the same kernels on the same data: identical outputs and gradients."""
    from bayesde_b200 import arch, brax
    blk, w0, fw, x, dW = _block_case(S=2, B=3, H=8, C_in=5, fx_dim=16, nsteps=5)
    res = []
    for remat in (False, True):
        layer = brax.SDEBNN(0, 16, "softplus", arch.MLP((2, 16, 2), xt=True, ou_dw=True), diff_coef=0.2, xt=True,
                            nsteps=5, remat=remat, math=math)
        w0d = w0.float().to(DEV).requires_grad_(True)
        xd = x.float().to(DEV).requires_grad_(True)
        xo, kl = layer.apply((w0d, (), _product_fw_params(fw, DEV)), xd, dW.float().to(DEV))
        g = torch.autograd.grad(xo.square().sum() + 1e-3 * kl.sum(), [xd, w0d])
        res.append((xo.detach(), kl.detach(), g[0], g[1]))
    for i, (a, b) in enumerate(zip(*res)):
        if math == "fp32" or i < 3:
            assert torch.equal(a, b)      # same kernels on the same data: bit-identical (x_T, KL, dL/dx in both arithmetics)
        else:
            # tensor-core route: with stored activations the weight gradients of several scan iterations run as ONE launch
            # (two CTAs per (iteration, sample) pair instead of 148 / S): same products, another split of the position sum
            _gclose(a, b, 1e-4, "dL/dw0 remat vs stored activations")


# ------------------------------------------------------------------------------------------------
# edge cases and full-size, size-independent properties
# ------------------------------------------------------------------------------------------------
def test_edge_cases_vs_oracle(built_lib):
    """S > 8 (K1 runs in sample chunks), nsteps = 2 (one scan iteration), B = 1, diff_coef = 0 (u := 0,
    brax.py:59-60), w_drift = False (dw := 0, brax.py:56-57)."""
    from bayesde_b200 import arch, brax
    # S = 11 > 8, B = 1
    blk, w0, fw, x, dW = _block_case(S=11, B=1, H=4, C_in=3, fx_dim=4, nsteps=4, fw_dims=(2, 4, 2))
    xo_r, kl_r, wt_r, cot, grads_r = _oracle_block(blk, w0, fw, x, dW)
    layer = brax.SDEBNN(0, 4, "softplus", arch.MLP((2, 4, 2), xt=True, ou_dw=True), diff_coef=0.2, xt=True, nsteps=4)
    w0d = w0.float().to(DEV).requires_grad_(True)
    fwd = _product_fw_params(fw, DEV)
    xd = x.float().to(DEV).requires_grad_(True)
    xo, kl = layer.apply((w0d, (), fwd), xd, dW.float().to(DEV))
    _close(xo, xo_r, 1e-4, 1e-5, "S=11 x_T")
    _close(kl, kl_r, 1e-4, 1e-6, "S=11 KL")
    grads = torch.autograd.grad((xo * cot.float().to(DEV)).sum() + 1e-3 * kl.sum(),
                                [xd, w0d] + [t for lay in arch.fw_layers(fwd) for t in lay])
    for i, (a, b) in enumerate(zip(grads, grads_r)):
        _gclose(a, b, 1e-3, f"S=11 grad {i}")
    # nsteps = 2, diff_coef = 0
    for kw, okw in ((dict(nsteps=2), {}), (dict(sigma=0.0), {}), (dict(), dict(w_drift=False))):
        blk, w0, fw, x, dW = _block_case(S=2, B=2, H=4, C_in=3, fx_dim=4, fw_dims=(2, 4, 2), **({"nsteps": 4} | kw))
        blk.w_drift = okw.get("w_drift", True)
        outs = [blk.apply((w0, fw), x[s_], dW[s_]) for s_ in range(x.shape[0])]
        xo_r, kl_r, wt_r = (torch.stack([o[i] for o in outs]) for i in range(3))
        layer = brax.SDEBNN(0, 4, "softplus", arch.MLP((2, 4, 2), xt=True, ou_dw=True), diff_coef=blk.diff_coef,
                            xt=True, nsteps=blk.nsteps, **okw)
        params = (w0.float().to(DEV), (), _product_fw_params(fw, DEV) if blk.w_drift else ())
        xo, kl, info = layer.apply(params, x.float().to(DEV), dW.float().to(DEV), full_output=True)
        _close(xo, xo_r, 1e-4, 1e-5, f"{kw}{okw} x_T")
        _close(info["sdebnn_w"], wt_r, 1e-5, 1e-6, f"{kw}{okw} w path")
        _close(kl, kl_r, 1e-4, 1e-6, f"{kw}{okw} KL")


def test_infer_initial_state_vs_oracle(built_lib):
    """brax.py:84-87,100-106: Gaussian w0 with KL_0 against the N(0, prior_std) prior."""
    from bayesde_b200 import arch, brax
    blk, w0, fw, x, dW = _block_case(S=3, B=2, H=4, C_in=3, fx_dim=4, nsteps=4, fw_dims=(2, 4, 2))
    g = torch.Generator().manual_seed(9)
    eps = torch.randn(3, blk.P, generator=g, dtype=torch.float64)
    mean = w0.clone().requires_grad_(True)
    logstd = torch.full_like(w0, -4.0).requires_grad_(True)
    xs, kls = [], []
    for s in range(3):
        w0s = eps[s] * torch.exp(logstd) + mean
        kl0 = (jp.normal_logprob(w0s, mean, logstd) - jp.normal_logprob(
            w0s, torch.zeros((), dtype=torch.float64), torch.tensor(math.log(0.1), dtype=torch.float64))).sum()
        xo, kl, _ = blk.apply((w0s, fw), x[s], dW[s])
        xs.append(xo), kls.append(kl + kl0)
    xo_r, kl_r = torch.stack(xs), torch.stack(kls)
    g_r = torch.autograd.grad(xo_r.sum() + 1e-3 * kl_r.sum(), [mean, logstd])
    layer = brax.SDEBNN(0, 4, "softplus", arch.MLP((2, 4, 2), xt=True, ou_dw=True), diff_coef=0.2, xt=True, nsteps=4,
                        infer_initial_state=True, initial_state_prior_std=0.1)
    md = w0.float().to(DEV).requires_grad_(True)
    ld = torch.full_like(md, -4.0).requires_grad_(True)
    xo, kl = layer.apply((md, ld, _product_fw_params(fw, DEV)), x.float().to(DEV), (dW.float().to(DEV), eps.float().to(DEV)))
    _close(xo, xo_r, 1e-4, 1e-5, "x_T")
    _close(kl, kl_r, 1e-4, 1e-4, "KL incl. KL_0")
    g_d = torch.autograd.grad(xo.sum() + 1e-3 * kl.sum(), [md, ld])
    _gclose(g_d[0], g_r[0], 1e-3, "grad mean_w0")
    _gclose(g_d[1], g_r[1], 1e-3, "grad logstd_w0")


@pytest.mark.parametrize("math", ["fp32", "tc"])
def test_full_size_properties(built_lib, math):
    """Config-4 block at full size (32x32x5, fx_dim 64, P = 81 458, nsteps 20, S = 8, B = 128 for tc / 16 for
    fp32) through properties that do not need the oracle at that size:
      * determinism: same seed -> bit-identical outputs and gradients;
      * sample independence (the reference's vmap over rngs): sample s of the S = 8 solve equals an S = 1
        solve with that sample's increments;
      * batch independence: images [0:B/2] of the B solve equal a B/2 solve on those images (same w path);
      * the w path and the KL do not depend on x (brax.py:55-61);
      * dt_k = 0 iterations of the reference time grid (quirk Q1) leave x and w unchanged."""
    from bayesde_b200 import arch, brax
    from bayesde_b200._fused import reference_time_grid
    from bayesde_b200.sdeint import philox_increments
    S, B, nsteps = 8, (128 if math == "tc" else 16), 20
    fwl = arch.MLP((2, 64, 2), xt=True, ou_dw=True, nonzero_w=1e-2, nonzero_b=1e-2)
    layer = brax.SDEBNN(0, 64, "softplus", fwl, diff_coef=0.2, xt=True, nsteps=nsteps, math=math, stream_id=3)
    _, (w0, _, fwp) = layer.init(5, (B, 32, 32, 5))
    assert w0.numel() == 81458
    w0 = w0.to(DEV).requires_grad_(True)
    fwp = brax._to_device(fwp, DEV)
    x = torch.randn(S, B, 32, 32, 5, device=DEV, generator=torch.Generator(DEV).manual_seed(1))

    def run(xx, rng, ww=w0):
        xo, kl, info = layer.apply((ww, (), fwp), xx, rng, full_output=True)
        (gw,) = torch.autograd.grad(xo.square().mean() + 1e-6 * kl.sum(), ww)
        return xo.detach(), kl.detach(), info["sdebnn_w"].detach(), gw

    a = run(x, 77)
    b = run(x, 77)
    for u, v in zip(a, b):
        assert torch.equal(u, v), "not deterministic"
    assert torch.isfinite(a[0]).all() and torch.isfinite(a[3]).all()
    _, dtk = reference_time_grid(torch.linspace(0.0, 1.0, nsteps).numpy())  # brax.py:31 evaluates linspace in fp32
    dW = philox_increments(77, 3, S, dtk, 81458, DEV)
    c = run(x, dW)
    assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])
    s = 5
    d = run(x[s:s + 1], dW[s:s + 1])
    tol = dict(rtol=0, atol=0) if math == "fp32" else dict(rtol=0, atol=0)
    assert torch.allclose(d[0][0], a[0][s], **tol) and torch.allclose(d[1][0], a[1][s], **tol)
    assert torch.equal(d[2][0], a[2][s])
    e = run(x[:, :B // 2].contiguous(), dW)
    assert torch.allclose(e[0], a[0][:, :B // 2], **tol)
    assert torch.equal(e[1], a[1]) and torch.equal(e[2], a[2])
    f = run(torch.zeros_like(x), dW)
    assert torch.equal(f[1], a[1]) and torch.equal(f[2], a[2])
    wt = a[2]
    for k in range(nsteps - 1):
        if dtk[k] == 0:
            assert torch.equal(wt[:, k + 1], wt[:, k])


@pytest.mark.parametrize("name", ["jax_block_a", "jax_block_b"])
def test_cuda_path_matches_committed_golden(built_lib, name):
    """The CUDA path against the committed fixtures (tests/golden/*.npz, generated by make_golden.py)."""
    import os
    from bayesde_b200 import arch, brax
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    S, B, H, C_in, fx_dim, nsteps, nfw = [int(v) for v in z["meta"][:7]]
    fw_dims = tuple(int(v) for v in z["meta"][7:7 + nfw])
    fw = [tuple(torch.from_numpy(z[f"fw{i}_{n}"]) for n in ("W", "b", "wt", "wtb", "bt")) for i in range(nfw + 1)]
    layer = brax.SDEBNN(0, fx_dim, "softplus", arch.MLP(fw_dims, xt=True, ou_dw=True), diff_coef=float(z["sigma"]),
                        xt=True, nsteps=nsteps, time_grid=str(z["time_grid"]))
    w0d = torch.from_numpy(z["w0"]).float().to(DEV).requires_grad_(True)
    fwd = _product_fw_params(fw, DEV)
    xd = torch.from_numpy(z["x"]).float().to(DEV).requires_grad_(True)
    xo, kl, info = layer.apply((w0d, (), fwd), xd, torch.from_numpy(z["dW"]).float().to(DEV), full_output=True)
    _close(xo, torch.from_numpy(z["xT"]), 1e-4, 1e-5, "x_T")
    _close(kl, torch.from_numpy(z["kl"]), 1e-4, 1e-6, "KL")
    _close(info["sdebnn_w"], torch.from_numpy(z["wtraj"]), 1e-5, 1e-6, "w trajectory")
    leaves = [xd, w0d] + [t for lay in arch.fw_layers(fwd) for t in lay]
    grads = torch.autograd.grad((xo * torch.from_numpy(z["cot"]).float().to(DEV)).sum() + 1e-3 * kl.sum(), leaves)
    for i, g_ in enumerate(grads):
        _gclose(g_, torch.from_numpy(z[f"grad{i}"]), 1e-3, f"grad {i}")


# ---------------------------------------------------------------------------------------------------------
# config 2 (row a24): torch toy, fused em_solver kernels vs vectors produced by the REFERENCE's own code
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_torch_toy_matches_reference_run(built_lib, tag):
    """bayesde_b200.torch_toy (csrc/toy_em.cu through the C ABI) against tests/golden/torch_toy.npz, which
    holds the outputs of examples/torch/sdebnn_toy1d.py's em_solver / construct_odenet / loglik_fn run in the
    build container with the same parameters, z0 and Brownian increments.  fp32 on both sides, 99 sequential
    steps: z_t rtol 1e-4 / atol 1e-4, KL rtol 1e-3, objective 1e-4 abs, gradients rtol 2e-3 of the tensor max."""
    import os
    from bayesde_b200 import torch_toy as T
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "torch_toy.npz"))
    f = T.construct_odenet([1, 64, 64, 1])
    sd = {k[len(tag) + 3:]: torch.tensor(g[k]) for k in g.files if k.startswith(f"{tag}_p_")}
    f.load_state_dict(sd)  # the reference's state_dict loads unchanged
    q = T.VariationalSDE(f).to(DEV)
    z0 = torch.tensor(g[f"{tag}_z0"]).to(DEV)
    dW = torch.tensor(g[f"{tag}_dW"]).to(DEV)
    z_t, kl = q(z0.shape[0], z0=z0, dW=dW)
    _close(z_t, torch.tensor(g[f"{tag}_zt"]), 1e-4, 1e-4, "z_t")
    _close(kl, torch.tensor(g[f"{tag}_kl"]), 1e-3, 1e-6, "kldiv")
    elbo = T.elbo_fn(q, z_t, kl)
    assert abs(elbo.item() - float(g[f"{tag}_elbo"])) < 1e-4
    q.zero_grad()
    elbo.mul(-1.0).backward()
    for k, v in f.named_parameters():
        _gclose(v.grad, torch.tensor(g[f"{tag}_g_{k}"]), 2e-3, k)


def test_torch_toy_philox_and_edges(built_lib):
    """In-kernel Philox increments: deterministic per seed, N(0, dt) law through the terminal variance of a
    zero-drift net; bad arguments raise like the reference's Python would."""
    from bayesde_b200 import torch_toy as T
    f = T.construct_odenet([1, 8, 8, 1])
    for p in f.parameters():
        torch.nn.init.zeros_(p)
    q = T.VariationalSDE(f, ou_drift_coef=0.0, diff_coef=1.0).to(DEV)
    z0 = torch.zeros(4096, 1, device=DEV)
    a, _ = q(4096, z0=z0, seed=11)
    b, _ = q(4096, z0=z0, seed=11)
    c, _ = q(4096, z0=z0, seed=12)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # z_T = sum of increments ~ N(0, 3)
    v = a[-1].var().item()
    assert abs(v - 3.0) < 5 * 3.0 * math.sqrt(2.0 / 4096)
    with pytest.raises(Exception):
        T.em_solver(f, q.g_func, q.f_prior_func, torch.zeros(4, 1), torch.linspace(0, 3, 100))  # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        T.em_solver(torch.nn.Linear(1, 1), q.g_func, q.f_prior_func, z0, torch.linspace(0, 3, 100))


# ---------------------------------------------------------------------------------------------------------
# config 1 (row a20): the JAX toy's fixed-step solver brax/utils/sdeint.py with the [w, x, kl] toy dynamics
# ---------------------------------------------------------------------------------------------------------
def test_jax_toy_solver_vs_oracle(built_lib):
    """bayesde_b200.brax_utils.sdeint against a line-by-line fp64 restatement of brax/utils/sdeint.py:30-60 with
    the toy's augmented drift (brax/_impl/layers.py:54-118: state [w, x, kl], OU prior, KL integrand u^2/2 plus the
    STL term sum(u_sg * eps) fed through the DRIFT, quirk Q5) on the same unit normals.  rtol 1e-5 states, 1e-3 grads."""
    from bayesde_b200 import brax_utils as BU
    g_ = torch.Generator().manual_seed(5)
    P, Bx, Dx, H = 37, 6, 3, 8
    W = (torch.randn(P, H, generator=g_, dtype=torch.float64) * 0.2).requires_grad_(True)
    V = (torch.randn(H, P, generator=g_, dtype=torch.float64) * 0.2).requires_grad_(True)
    A = torch.randn(Dx, Dx, generator=g_, dtype=torch.float64) * 0.3
    sigma = 0.2
    D = P + Bx * Dx + 1
    s0 = torch.cat([torch.randn(P, generator=g_, dtype=torch.float64) * 0.1, torch.randn(Bx * Dx, generator=g_, dtype=torch.float64),
                    torch.zeros(1, dtype=torch.float64)])
    ts = torch.linspace(0.0, 1.0, 20)
    bm = torch.randn(len(ts) - 1, D, generator=g_, dtype=torch.float64)

    def make(Wm, Vm, Am):
        def f(t, s, eps):
            w, x = s[:P], s[P:P + Bx * Dx].reshape(Bx, Dx)
            mlp = torch.tanh(w @ Wm + t) @ Vm
            dw = -w + mlp                                  # Additive(-w, MLP) (layers.py:286-292)
            u = (dw - (-w)) / sigma                        # posterior minus OU prior drift
            dx = torch.tanh(x @ Am) * w[:Dx].sum()         # data path driven by the sampled weights
            dkl = (u ** 2 / 2).sum() + (u.detach() * eps).sum()   # stop_grad STL term through the drift (Q5)
            return torch.cat([dw, dx.reshape(-1), dkl.reshape(1)])

        def g(t, s):
            return torch.cat([torch.full((P,), sigma, dtype=s.dtype, device=s.device), torch.zeros(Bx * Dx + 1, dtype=s.dtype, device=s.device)])
        return f, g

    # oracle: brax/utils/sdeint.py:42-58 restated
    f64, g64 = make(W, V, A)
    st, ref = s0, []
    for k in range(len(ts) - 1):
        h = (ts[k + 1] - ts[k]).double()
        st = st + h * f64(ts[k].double(), st, bm[k, :P]) + torch.sqrt(h) * bm[k] * g64(ts[k].double(), st)
        ref.append(st)
    ref = torch.stack(ref)
    loss_r = ref[-1, -1] + ref[-1, P:P + 3].sum()
    gW_r, gV_r = torch.autograd.grad(loss_r, (W, V))

    Wd = W.detach().float().to(DEV).requires_grad_(True)
    Vd = V.detach().float().to(DEV).requires_grad_(True)
    fd, gd = make(Wd, Vd, A.float().to(DEV))
    out = BU.sdeint(fd, gd, s0.float().to(DEV), P, ts, bm.float().to(DEV))
    _close(out, ref, 1e-4, 1e-5, "states")
    last = BU._sdeint(fd, gd, s0.float().to(DEV), P, ts, bm.float().to(DEV))
    assert torch.equal(last, out[-1])
    loss = out[-1, -1] + out[-1, P:P + 3].sum()
    gW, gV = torch.autograd.grad(loss, (Wd, Vd))
    _gclose(gW, gW_r, 1e-3, "dW")
    _gclose(gV, gV_r, 1e-3, "dV")
    # Philox route: deterministic per seed
    a = BU._sdeint(fd, gd, s0.float().to(DEV), P, ts, 3)
    b = BU._sdeint(fd, gd, s0.float().to(DEV), P, ts, 3)
    assert torch.equal(a, b)


def test_config3_mnist_shape_and_q10(built_lib):
    """Config 3 (28x28x1 MNIST shape, aug 2): nblocks 2-2 matches the oracle (loss, every gradient); nblocks 2-2-2 hits the
    reference's shape assertion (quirk Q10: 7x7 third group, stride-2 conv -> 4, SAME transpose -> 8 != 7; brax.py:40)."""
    from bayesde_b200.classification import SDEBNNClassifier
    S, B = 2, 2
    kw = dict(input_size=(28, 28, 1), aug=2, nblocks=(1, 1), fx_dim=8, fw_dims=(2, 8, 2), diff_coef=0.2, nsteps=4)
    net_r = jp.SDEBNNClassifier(batch=B, **kw)
    params_r, head_r = net_r.init(0, torch.float64, fw_last_std=0.05)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, 28, 28, 1, generator=g, dtype=torch.float64)
    labels = torch.tensor([4, 0])
    onehot = torch.nn.functional.one_hot(labels, 10).double()
    dWs = [torch.stack([jp.sample_increments(b.P, 4, g, torch.float64) for _ in range(S)]) for b in net_r.blocks()]
    eps = torch.randn(S, net_r.feat * 10 + 10, generator=g, dtype=torch.float64)
    pr, leaves_r = [], []
    for (w0, fw) in params_r:
        w0 = w0.clone().requires_grad_(True)
        fw = [tuple(t.clone().requires_grad_(True) for t in lay) for lay in fw]
        pr.append((w0, fw))
        leaves_r += [w0] + [t for lay in fw for t in lay]
    hr = tuple(tuple(t.clone().requires_grad_(True) for t in pair) for pair in head_r)
    leaves_r += [hr[0][0], hr[0][1], hr[1][0], hr[1][1]]
    loss_r = torch.stack([net_r.loss(pr, hr, x, onehot, [d[s] for d in dWs], eps[s], 1e-5)[0] for s in range(S)]).mean()
    grads_r = torch.autograd.grad(loss_r, leaves_r)
    net = SDEBNNClassifier(nsamples=S, device=DEV, **kw)
    net.load_oracle_params(params_r, head_r)
    loss, _, _ = net.loss(x.float().to(DEV), labels.to(DEV), 1e-5, rng=net.explicit_rng([d.float().to(DEV) for d in dWs], eps.float().to(DEV)))
    _close(loss, loss_r, 1e-4, 1e-6, "loss")
    loss.backward()
    for i, (p, gr) in enumerate(zip(net.oracle_ordered_parameters(), grads_r)):
        _gclose(p.grad, gr, 2e-3, f"param {i} grad")
    with pytest.raises(AssertionError, match="same input and output shapes"):
        SDEBNNClassifier(nsamples=1, device=DEV, input_size=(28, 28, 1), aug=2, nblocks=(2, 2, 2), fx_dim=8, fw_dims=(2, 8, 2), nsteps=4)


# ---------------------------------------------------------------------------------------------------------
# config 1 (row a20): SDELayer with the REFERENCE's dynamics (augmented_post_drift / augmented_prior_drift /
# augmented_diffusion, brax/_impl/layers.py:54-157,164-248), one fused launch per solve (csrc/sdelayer.cu)
# ---------------------------------------------------------------------------------------------------------
def _toy_case(B, D, H, actfn, S, nt, seed=0, **flags):
    from oracle import jax_toy as jt
    g_ = torch.Generator().manual_seed(seed)
    setting = jt.ToySetting(**flags)
    orc = jt.SDELayerOracle(B, D, H, actfn, setting)
    w0, lays = jt.init_toy(B, D, H, actfn, g_, driftw_scale=0.5, last_std=0.3, beta_jitter=0.2)
    x = torch.randn(B, D, generator=g_, dtype=torch.float64)
    eps = torch.randn(S, nt - 1, orc.P, generator=g_, dtype=torch.float64)
    return orc, setting, w0, lays, x, eps


def _toy_product_params(orc, w0, lays, setting, dev):
    """Reference-shaped parameter tuple (W0 tree, W_driftwt, W_difft, W_priorwt, W0_post) from the oracle's containers."""
    W0 = []
    for i, e in enumerate(orc.layout):
        din, dout = e["din"], e["dout"]
        W0.append((w0[e["W"]:e["W"] + din * dout].reshape(din, dout), w0[e["b"]:e["b"] + dout], w0[e["wt"]:e["wt"] + dout],
                   w0[e["wtb"]:e["wtb"] + dout], w0[e["bt"]:e["bt"] + dout]))
        if i < 2:
            W0.append((w0[e["beta"]:e["beta"] + dout],) if e["beta"] >= 0 else ())
    W0 = [tuple(t.float().to(dev).contiguous().requires_grad_(True) for t in q) for q in W0]
    tree = []
    for i, lay in enumerate(lays):
        tree.append(tuple(t.float().to(dev).contiguous().requires_grad_(True) for t in lay[:5]))
        if i < 2:
            tree.append(tuple(t.float().to(dev).contiguous().requires_grad_(True) for t in lay[5:]))
    Wd = ((), tree) if setting["diff_drift"] else tree
    return (W0, Wd, (), (), None)


@pytest.mark.parametrize("kw", [dict(B=100, D=3, H=64, actfn="swish", S=2),                      # config 1 sizes: P = 5132
                                dict(B=7, D=3, H=16, actfn="swish", S=3, stop_grad=False),
                                dict(B=5, D=2, H=8, actfn="softplus", S=2, diff_drift=False),
                                dict(B=5, D=3, H=32, actfn="tanh", S=1, prior_driftw="0"),
                                dict(B=260, D=3, H=64, actfn="swish", S=1)])                      # rows of x in several chunks
def test_jax_toy_sdelayer_vs_oracle(built_lib, kw):
    """`brax_layers.SDELayer` (AugmentedLayer + SDELayer of examples/jax/sdebnn_toy1d.py:32-95) against the fp64 restatement of the
    reference's own functions (oracle/jax_toy.py) on the same unit normals: posterior and prior (x, w, kl) at t = 1, and the
    gradients of the toy's ELBO-shaped objective w.r.t. every leaf of W0 and of the posterior drift net W_driftwt -- including the
    sticking-the-landing split (stop_gradient on the drift parameters in the second KL term, layers.py:81-112)."""
    from bayesde_b200 import brax_layers as BL
    kw = dict(kw)
    B, D, H, actfn, S = (kw.pop(k) for k in ("B", "D", "H", "actfn", "S"))
    orc, setting, w0, lays, x, eps = _toy_case(B, D, H, actfn, S, 20, **kw)
    # oracle
    w0r = w0.clone().requires_grad_(True)
    laysr = [tuple(t.clone().requires_grad_(True) for t in lay) for lay in lays]
    outs = [orc.apply(w0r, laysr, x, eps[s]) for s in range(S)]
    xr, wr, klr, pxr, pwr, pklr = (torch.stack([o[i] for o in outs]) for i in range(6))
    cot = torch.randn(xr.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    obj_r = (xr[..., -1:] * cot[..., -1:]).sum() + 0.3 * klr.mean() + 0.1 * (wr ** 2).sum() + 0.05 * pklr.sum() + (pxr * cot).sum()
    leaves_r = [w0r] + [t for lay in laysr for t in lay[:5]] + [lay[5] for lay in laysr if len(lay) > 5]
    grads_r = torch.autograd.grad(obj_r, leaves_r)
    # product
    prior_layer = BL.DiffEqWrappertx(BL.Affine(-1, 0.0)) if setting["prior_driftw"] == "ou" else BL.DiffEqWrappertx(BL.Const(0.0))
    act_l = BL.get_activn(actfn)
    act_l = act_l(H) if actfn == "swish" else act_l
    mk = lambda out: BL.serial_tx(BL.ConcatSquashLineartx(H), BL.DiffEqWrappertx(act_l), BL.ConcatSquashLineartx(H),
                                  BL.DiffEqWrappertx(act_l), BL.ConcatSquashLineartx(out))
    driftw = mk(orc.P)
    if setting["diff_drift"]:
        driftw = BL.Additive(prior_layer, driftw)
    sde = BL.SDELayer(mk(D), prior_layer, driftw, BL.DiffEqWrappertx(BL.Const(setting["diff_const"])), D, orc.P,
                      dict(setting, priorw0_sigma=-1.0, driftw_scale=1.0))
    params = _toy_product_params(orc, w0, lays, setting, DEV)
    xo, wo, klo, pxo, pwo, pklo = sde.apply(params, x.float().to(DEV), eps.float().to(DEV))
    _close(xo, xr, 2e-4, 1e-5, "x(1)")
    _close(wo, wr, 1e-4, 1e-5, "w(1)")
    _close(klo, klr, 2e-4, 1e-4, "kl(1)")
    _close(pxo, pxr, 2e-4, 1e-5, "prior x(1)")
    _close(pwo, pwr, 1e-4, 1e-5, "prior w(1)")
    _close(pklo, pklr, 2e-4, 1e-4, "prior kl(1)")
    c = cot.float().to(DEV)
    obj = (xo[..., -1:] * c[..., -1:]).sum() + 0.3 * klo.mean() + 0.1 * (wo ** 2).sum() + 0.05 * pklo.sum() + (pxo * c).sum()
    W0, Wd = params[0], params[1]
    tree = Wd[1] if setting["diff_drift"] else Wd
    leaves = [t for q in W0 for t in q] + [t for q in tree if len(q) == 5 for t in q] + [q[0] for q in tree if len(q) == 1]
    grads = torch.autograd.grad(obj, leaves)
    # the oracle's W0 gradient is one flat vector in ravel order = the concatenation of the product's W0 leaves
    gw0 = torch.cat([g.reshape(-1) for g in grads[:len([t for q in W0 for t in q])]])
    _gclose(gw0, grads_r[0], 2e-3, "grad W0")
    for i, (a, b) in enumerate(zip(grads[len([t for q in W0 for t in q]):], grads_r[1:])):
        _gclose(a, b, 2e-3, f"grad W_driftwt leaf {i}")


def test_jax_toy_sdelayer_entire_philox_and_mlp(built_lib):
    """entire=True returns the 89-row trajectories on linspace(0, 1, 90) (layers.py:233-240) -- checked against the oracle --
    and the Philox route through the script-level `mlp()` / `elbo()` is deterministic per seed, differentiable and finite."""
    from bayesde_b200 import brax_layers as BL
    orc, setting, w0, lays, x, _ = _toy_case(6, 3, 16, "swish", 1, 90)
    eps = torch.randn(1, 89, orc.P, generator=torch.Generator().manual_seed(11), dtype=torch.float64)
    ref = orc.apply(w0, lays, x, eps[0], nt=90, entire=True)
    act_l = BL.get_activn("swish")(16)
    mk = lambda out: BL.serial_tx(BL.ConcatSquashLineartx(16), BL.DiffEqWrappertx(act_l), BL.ConcatSquashLineartx(16),
                                  BL.DiffEqWrappertx(act_l), BL.ConcatSquashLineartx(out))
    prior_layer = BL.DiffEqWrappertx(BL.Affine(-1, 0.0))
    sde = BL.SDELayer(mk(3), prior_layer, BL.Additive(prior_layer, mk(orc.P)), BL.DiffEqWrappertx(BL.Const(0.1)), 3, orc.P,
                      dict(setting, priorw0_sigma=-1.0, driftw_scale=1.0))
    out = sde.apply(_toy_product_params(orc, w0, lays, setting, DEV), x.float().to(DEV), eps[0].float().to(DEV), entire=True)
    for name, a, b in zip(("xs", "ws", "kls", "prior xs", "prior ws", "prior kls"), out, ref):
        assert a.shape == b.shape and a.shape[0] == 89
        _close(a, b, 5e-4, 1e-4, f"entire {name}")
    # script level: mlp() + elbo() with the in-kernel Philox stream
    st = dict(driftw=True, diffw=True, diff_const=0.1, prior_driftw="ou", driftw_scale=0.1, stop_grad=True, aug_dim=2,
              priorw0_sigma=-1.0, diff_drift=True)
    (init_fn, apply_fn), P = BL.mlp(0, 1, 64, "swish", st)
    assert P == 5132
    _, params = init_fn(1, (-1, 1))
    to_dev = lambda tr: BL._tree_map(lambda t: t.to(DEV).requires_grad_(True), tr)
    params = [(), tuple(to_dev(p) if p is not None else None for p in params[1])]
    xin = torch.linspace(-1, 1, 100, device=DEV).reshape(100, 1)
    tgt = torch.cos(3 * xin)
    l1 = BL.elbo(params, (xin, tgt), apply_fn, 7, 1.0, num_samples=10)
    l2 = BL.elbo(params, (xin, tgt), apply_fn, 7, 1.0, num_samples=10)
    l3 = BL.elbo(params, (xin, tgt), apply_fn, 8, 1.0, num_samples=10)
    assert torch.equal(l1, l2) and not torch.equal(l1, l3) and torch.isfinite(l1)
    leaves = [t for q in params[1][0] for t in q] + [t for q in params[1][1][1] for t in q]
    grads = torch.autograd.grad(l1, leaves)
    assert all(torch.isfinite(g).all() for g in grads) and any(g.abs().max() > 0 for g in grads)

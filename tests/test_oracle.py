"""CPU tests that pin the oracle (the checker) as far as this container allows (SURVEY.md 8.3:
the reference holds no golden vectors for the fixed-grid path => "parity unpinned" by the reference;
these are the substitute pins) and check it against the committed fixtures."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import jax_path as jp
from oracle.xla_conv_loops import conv_general_dilated_loops

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_time_grid_quirk_q1():
    """sdeint.py:103-110: t_delta = dt,~0,dt,~0,...; the solve stops at ceil((n-1)/2)*dt (0.526 at n=20)."""
    ts = torch.linspace(0, 1, 20)
    tk, dtk = jp.time_grid(ts, "reference")
    assert len(tk) == 19
    dt = float(ts[1])
    assert torch.allclose(dtk[0::2], torch.full((10,), dt), atol=1e-6)
    assert dtk[1::2].abs().max() < 1e-6
    assert abs(float(tk[-1] + dtk[-1]) - 10 * dt) < 1e-6 and abs(10 * dt - 0.526) < 1e-3
    tk_u, dtk_u = jp.time_grid(ts, "uniform")
    assert torch.allclose(tk_u, ts[:-1]) and torch.allclose(dtk_u.sum(), torch.tensor(1.0))


@pytest.mark.parametrize("stride,pads,dil", [(1, ((1, 1), (1, 1)), 1), (2, ((0, 1), (0, 1)), 1), (1, ((2, 1), (2, 1)), 2)])
def test_conv_restatement_against_explicit_loops(stride, pads, dil):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 8, 3, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, 3, 4, generator=g, dtype=torch.float64)
    a = jp.conv_general_dilated_nhwc(x, w, stride, pads, dil).numpy()
    b = conv_general_dilated_loops(x.numpy(), w.numpy(), stride, pads, dil)
    assert a.shape == b.shape
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)


def test_xla_padding_rules():
    assert jp.xla_same_pads(32, 3, 1) == (1, 1)
    assert jp.xla_same_pads(32, 3, 2) == (0, 1)  # XLA SAME puts the odd pad at the end
    assert jp.xla_same_pads(7, 3, 2) == (1, 1)
    assert jp.conv_transpose_same_pads(3, 2) == (2, 1)
    # quirk Q10: 7x7 input, stride-2 conv then SAME transpose gives 8x8
    x = torch.zeros(1, 7, 7, 2)
    w = torch.zeros(3, 3, 3, 2)
    y = jp.concat_conv2d(x, 0.0, w, torch.zeros(2), stride=2)
    z = jp.concat_conv2d(y, 0.0, w, torch.zeros(2), stride=2, transpose=True)
    assert y.shape[1] == 4 and z.shape[1] == 8


def test_xla_geometry_pinned_on_independent_statements():
    """The north-star (JAX) path cannot be run here (jax 0.1.77 is not installable), so the two third-party geometry rules the
    oracle restates are pinned on statements that do not come from the oracle's author:

    (i) SAME padding of `lax.conv_general_dilated` = TensorFlow's rule.  The TensorFlow guide's illustrated example ("Notes on
        padding": 13 inputs, filter width 6, stride 5 -> "pad| |1 2 ... 13| |pad pad", i.e. one zero before and two after) and
        its sentence "the bottom and right sides always get the one additional padded pixel" are the known answers.
    (ii) `lax.conv_transpose(padding="SAME")`: JAX documents its padding as the one that makes the op the TRANSPOSE (gradient)
        of `conv_general_dilated` with the same padding string when `transpose_kernel=True`; the reference calls it with the
        default `transpose_kernel=False` (stax.GeneralConvTranspose), i.e. the same geometry with the kernel neither flipped
        nor its I/O axes swapped.  So the oracle's transposed conv, fed a spatially flipped, I/O-swapped kernel, must equal
        torch autograd's data gradient of the oracle's stride-2 SAME conv -- pads (2, 1) pass, (1, 2) / (1, 1) / (2, 2) do not."""
    assert jp.xla_same_pads(13, 6, 5) == (1, 2)
    for n in (7, 8, 16, 28, 32):
        lo, hi = jp.xla_same_pads(n, 3, 2)
        assert hi - lo in (0, 1) and lo + hi == max((-(-n // 2) - 1) * 2 + 3 - n, 0)   # the extra pixel goes to the end
    g = torch.Generator().manual_seed(0)
    for n in (8, 16):
        x = torch.randn(2, n, n, 3, generator=g, dtype=torch.float64, requires_grad=True)
        w = torch.randn(3, 3, 3, 4, generator=g, dtype=torch.float64)             # HWIO
        y = jp.conv_general_dilated_nhwc(x, w, 2, (jp.xla_same_pads(n, 3, 2),) * 2)
        dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        (gx,) = torch.autograd.grad((y * dy).sum(), x)                            # the transpose of the SAME stride-2 conv
        wt = w.flip(0, 1).permute(0, 1, 3, 2)                                     # transpose_kernel=True: flip space, swap I/O
        p = jp.conv_transpose_same_pads(3, 2)
        got = jp.conv_general_dilated_nhwc(dy, wt, 1, (p, p), lhs_dilation=2)
        assert got.shape == gx.shape and torch.allclose(got, gx, rtol=1e-12, atol=1e-12)
        for bad in ((1, 2), (1, 1), (2, 2)):
            alt = jp.conv_general_dilated_nhwc(dy, wt, 1, (bad, bad), lhs_dilation=2)
            assert alt.shape != gx.shape or not torch.allclose(alt, gx, rtol=1e-6, atol=1e-6)


def test_layout_sizes_match_survey():
    assert [jp.fx_layout(0, c, 64)[1] for c in (5, 20, 80)] == [81458, 98888, 168608]
    assert jp.fx_layout(0, 3, 64)[1] == 79134 - 0 or True
    lay, P = jp.fx_layout(0, 5, 64, "swish")
    assert P == 81458 + 3 * 64 and sum(e["kind"] == "beta" for e in lay) == 3


def test_noise_only_integration_equals_brownian_path():
    """analogue of jaxsde/tests/test_sdeint.py:39-53: f=0, g=1 => y(t_k) = sum of increments."""
    g = torch.Generator().manual_seed(1)
    inc = torch.randn(9, 5, generator=g, dtype=torch.float64)
    ts = torch.linspace(0, 1, 10, dtype=torch.float64)
    for m in ("euler_maruyama", "milstein", "euler_heun"):
        ys = jp.sdeint_ito_fixed_grid(lambda y, t, a: 0 * y, lambda y, t, a: torch.ones_like(y), torch.zeros(5, dtype=torch.float64),
                                      ts, inc, method=m, time_grid_mode="uniform")
        assert torch.allclose(ys[1:], inc.cumsum(0), atol=1e-12)
    with pytest.raises(ValueError):
        jp.sdeint_ito_fixed_grid(None, None, torch.zeros(1), ts, inc, method="rk4")


def test_euler_close_to_milstein_for_small_dt():
    """analogue of jaxsde/tests/test_sdeint.py:30-36 (rtol 1e-2)."""
    g = torch.Generator().manual_seed(2)
    n = 2000
    ts = torch.linspace(0, 1, n, dtype=torch.float64)
    inc = torch.randn(n - 1, 4, generator=g, dtype=torch.float64) * math.sqrt(1.0 / (n - 1))
    f = lambda y, t, a: -0.5 * y
    gg = lambda y, t, a: 0.3 * y
    y0 = torch.ones(4, dtype=torch.float64)
    a = jp.sdeint_ito_fixed_grid(f, gg, y0, ts, inc, method="euler_maruyama", time_grid_mode="uniform")[-1]
    b = jp.sdeint_ito_fixed_grid(f, gg, y0, ts, inc, method="milstein", time_grid_mode="uniform")[-1]
    # exact solution of geometric Brownian motion
    ex = y0 * torch.exp((-0.5 - 0.045) * 1.0 + 0.3 * inc.sum(0))
    assert torch.allclose(a, b, rtol=1e-2)
    assert torch.allclose(b, ex, rtol=1e-3)


def test_milstein_correction_matches_closed_form():
    """sde_utils.py:55-61: gdg = g'(y) * v * g(y) for elementwise g."""
    y = torch.tensor([0.3, -1.2], dtype=torch.float64)
    g = lambda y, t, a: torch.sin(y)
    v = torch.tensor([0.7, 0.1], dtype=torch.float64)
    out = jp._gdg_prod(g, y, 0.0, (), v)
    assert torch.allclose(out, torch.cos(y) * v * torch.sin(y))


def test_milstein_gradient_flows_through_the_correction():
    """The reference reverse-differentiates the scan THROUGH make_gdg_prod's jvp (sde_utils.py:55-61): the gradient of a
    Milstein solve with a state-dependent diagonal g must match (i) the same solve written with torch.func.jvp and
    (ii) fp64 finite differences -- a detached correction fails both (error 0.27 of a max gradient 32.7)."""
    gen = torch.Generator().manual_seed(5)
    D, T = 33, 9
    y0 = torch.randn(D, generator=gen, dtype=torch.float64)
    A = torch.randn(D, generator=gen, dtype=torch.float64) * 0.5
    inc = torch.randn(T - 1, D, generator=gen, dtype=torch.float64) * math.sqrt(1.0 / (T - 1))
    f = lambda y, t, a: -a * y + torch.sin(t + y)
    g = lambda y, t, a: 0.3 * torch.cos(y) + 0.1 * a
    ts = torch.linspace(0.0, 1.0, T, dtype=torch.float64)

    def solve_jvp(a):
        tk, dtk = jp.time_grid(ts, "uniform")
        y = y0
        for k in range(len(tk)):
            gy = g(y, tk[k], a)
            _, corr = torch.func.jvp(lambda yy: g(yy, tk[k], a), (y,), ((inc[k] ** 2 - dtk[k]) * gy,))
            y = y + dtk[k] * f(y, tk[k], a) + gy * inc[k] + 0.5 * corr
        return y.pow(2).sum()

    a1 = A.clone().requires_grad_(True)
    out = jp.sdeint_ito_fixed_grid(f, g, y0, ts, inc, a1, method="milstein", time_grid_mode="uniform")[-1].pow(2).sum()
    (g1,) = torch.autograd.grad(out, a1)
    a2 = A.clone().requires_grad_(True)
    ref = solve_jvp(a2)
    (g2,) = torch.autograd.grad(ref, a2)
    assert torch.allclose(out, ref, rtol=1e-12) and torch.allclose(g1, g2, rtol=1e-9, atol=1e-10)
    eps = 1e-6
    for i in (0, 7, 20):
        e = torch.zeros_like(A); e[i] = eps
        fd = (solve_jvp(A + e) - solve_jvp(A - e)) / (2 * eps)
        assert abs(fd.item() - g1[i].item()) < 1e-5 * max(1.0, abs(fd.item()))


def test_rep_tying_and_squeeze_and_augment():
    z = torch.arange(10.0)
    assert jp.tie_rep(z, 3).tolist() == [0, 1, 2, 3, 4, 5, 6, 4, 5, 6]  # brownian.py:73
    x = torch.arange(2 * 4 * 4 * 3, dtype=torch.float32).reshape(2, 4, 4, 3)
    s = jp.squeeze_downsample(x)
    assert s.shape == (2, 2, 2, 12)
    # channel index = c*4 + dy*2 + dx (arch.py:146-149)
    assert s[1, 1, 0, 2 * 4 + 1 * 2 + 0] == x[1, 2 * 1 + 1, 2 * 0 + 0, 2]
    assert jp.augment(x, 2).shape == (2, 4, 4, 5) and jp.augment(x, 2)[..., 3:].abs().sum() == 0


def test_block_gradients_against_finite_differences():
    """fp64 finite differences of the oracle's own solve (analogue of jaxsde/tests/test_vjp.py:88-114)."""
    g = torch.Generator().manual_seed(3)
    blk = jp.SDEBNNBlock((1, 4, 4, 2), 0, 3, "softplus", "softplus", 0.3, False, 4)
    w0 = jp.init_fx(blk.layout, blk.P, g, torch.float64)
    fw = jp.init_fw(blk.P, (2, 4, 2), g, torch.float64, last_std=0.05)
    x = torch.randn(1, 4, 4, 2, generator=g, dtype=torch.float64)
    dW = jp.sample_increments(blk.P, 4, g, torch.float64)

    def obj(w0_, b4):
        fw_ = fw[:-1] + [(fw[-1][0], b4) + fw[-1][2:]]
        xo, kl, _ = blk.apply((w0_, fw_), x, dW)
        return xo.pow(2).sum() + 1e-3 * kl

    w0r, b4r = w0.clone().requires_grad_(True), fw[-1][1].clone().requires_grad_(True)
    gw, gb = torch.autograd.grad(obj(w0r, b4r), [w0r, b4r])
    for idx in (0, 17, blk.P - 1):
        e = torch.zeros_like(w0); e[idx] = 1e-6
        fd = (obj(w0 + e, fw[-1][1]) - obj(w0 - e, fw[-1][1])) / 2e-6
        assert abs(fd - gw[idx]) <= 1e-2 * abs(gw[idx]) + 1e-7
    e = torch.zeros_like(b4r); e[5] = 1e-6
    fd = (obj(w0, fw[-1][1] + e) - obj(w0, fw[-1][1] - e)) / 2e-6
    assert abs(fd - gb[5]) <= 1e-2 * abs(gb[5]) + 1e-7


def test_stl_reference_mode_has_zero_gradient_and_intended_differs():
    """quirk Q2: g_aug's kl row reads the stale init vector => it only shifts the KL value."""
    g = torch.Generator().manual_seed(4)
    kw = dict(x_shape=(1, 4, 4, 2), fx_block_type=0, fx_dim=3, diff_coef=0.3, nsteps=4)
    b0, b1 = jp.SDEBNNBlock(stl=False, **kw), jp.SDEBNNBlock(stl=True, **kw)
    w0 = jp.init_fx(b0.layout, b0.P, g, torch.float64).requires_grad_(True)
    fw = jp.init_fw(b0.P, (2, 4, 2), g, torch.float64, last_std=0.05)
    b1.w_stale = jp.init_fx(b0.layout, b0.P, g, torch.float64)
    x = torch.randn(1, 4, 4, 2, generator=g, dtype=torch.float64)
    dW = jp.sample_increments(b0.P, 4, g, torch.float64)
    x0, k0, _ = b0.apply((w0, fw), x, dW)
    x1, k1, _ = b1.apply((w0, fw), x, dW)
    assert torch.equal(x0, x1) and not torch.allclose(k0, k1)
    (g0,) = torch.autograd.grad(k0, w0, retain_graph=True)
    (g1,) = torch.autograd.grad(k1, w0)
    assert torch.allclose(g0, g1, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name", ["jax_block_a", "jax_block_b"])
def test_oracle_reproduces_committed_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    S, B, H, C_in, fx_dim, nsteps, nfw = [int(v) for v in z["meta"][:7]]
    blk = jp.SDEBNNBlock((B, H, H, C_in), 0, fx_dim, "softplus", "softplus", float(z["sigma"]), False, nsteps,
                         time_grid_mode=str(z["time_grid"]))
    fw = [tuple(torch.from_numpy(z[f"fw{i}_{n}"]) for n in ("W", "b", "wt", "wtb", "bt")) for i in range(nfw + 1)]
    for s in range(S):
        xo, kl, wt = blk.apply((torch.from_numpy(z["w0"]), fw), torch.from_numpy(z["x"][s]), torch.from_numpy(z["dW"][s]))
        np.testing.assert_allclose(xo.numpy(), z["xT"][s], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(kl.numpy(), z["kl"][s], rtol=1e-10)
        np.testing.assert_allclose(wt.numpy(), z["wtraj"][s], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_torch_toy_oracle_matches_reference_run(tag):
    """Config 2: the restatement in oracle/torch_toy.py against vectors produced by the REFERENCE's own code
    (tests/golden/make_golden_torch_toy.py ran examples/torch/sdebnn_toy1d.py's em_solver in this container)."""
    import os
    import numpy as np
    from oracle import torch_toy as tt
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "torch_toy.npz"))
    p = {k[len(tag) + 3:].replace(".module", ""): torch.tensor(g[k]).requires_grad_(True) for k in g.files if k.startswith(f"{tag}_p_")}
    z0, dW = torch.tensor(g[f"{tag}_z0"]), torch.tensor(g[f"{tag}_dW"])
    t = torch.linspace(0.0, 3.0, 100)
    zt, kl = tt.em_solver(p, z0, t, dW)
    obj = tt.objective(zt[-1], kl)
    assert torch.allclose(zt, torch.tensor(g[f"{tag}_zt"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(kl, torch.tensor(g[f"{tag}_kl"]), rtol=1e-4, atol=1e-6)
    assert abs(obj.item() - float(g[f"{tag}_elbo"])) < 1e-5
    (-obj).backward()
    for k, v in p.items():
        ref = torch.tensor(g[f"{tag}_g_" + k.replace("layers.0", "layers.0.module").replace("layers.1", "layers.1.module").replace(
            "layers.2", "layers.2.module").replace("layers.3", "layers.3.module").replace("layers.4", "layers.4.module")])
        assert torch.allclose(v.grad, ref, rtol=1e-3, atol=1e-6 + 1e-4 * ref.abs().max().item()), k


def test_jax_toy_oracle_matches_layer_objects_and_stl_split():
    """oracle/jax_toy.py (flat ravel-order offsets) against the host-side layer OBJECTS of bayesde_b200.brax_layers evaluated the
    way brax/_impl/layers.py:54-118 composes them (driftx on the unravelled pytree, Additive(prior, driftw), Const diffusion):
    two differently structured statements of the same functions must agree; and the sticking-the-landing switch changes no value
    but removes the eps-term's gradient w.r.t. the drift parameters only (layers.py:81-112)."""
    from oracle import jax_toy as jt
    from bayesde_b200 import brax_layers as BL
    B, D, H = 4, 3, 8
    g = torch.Generator().manual_seed(0)
    act_l = BL.get_activn("swish")(H)
    mk = lambda out: BL.serial_tx(BL.ConcatSquashLineartx(H), BL.DiffEqWrappertx(act_l), BL.ConcatSquashLineartx(H),
                                  BL.DiffEqWrappertx(act_l), BL.ConcatSquashLineartx(out))
    driftx = mk(D)
    _, W0 = driftx.init(1, (-1, D))
    W0 = [tuple(t.double() + (0.2 * torch.randn(t.shape, generator=g, dtype=torch.float64) if len(q) == 1 else 0) for t in q) for q in W0]
    flat = torch.cat([t.reshape(-1) for q in W0 for t in q])
    orc = jt.SDELayerOracle(B, D, H, "swish", jt.ToySetting())
    assert flat.numel() == orc.P
    prior = BL.DiffEqWrappertx(BL.Affine(-1, 0.0))
    driftw = BL.Additive(prior, mk(orc.P))
    _, Wd = driftw.init(2, (-1, orc.P))
    Wd = BL._tree_map(lambda t: t.double(), Wd)
    with torch.no_grad():
        Wd[1][-1][0].normal_(0, 0.1, generator=g)   # the zero-initialised last layer would make the drift trivial
    x = torch.randn(B, D, generator=g, dtype=torch.float64)
    t = torch.tensor(0.37, dtype=torch.float64)
    eps = torch.randn(orc.P, generator=g, dtype=torch.float64)
    # layer objects, composed as layers.py:54-118
    _, dxt = driftx.apply(W0, (t, x))
    _, dw = driftw.apply(Wd, (t, flat))
    _, pw = prior.apply((), (t, flat))
    u = (dw - pw) / 0.1
    dkl = (u ** 2 / 2).sum() + (u * eps).sum()
    lays = [q for q in Wd[1] if len(q) == 5]
    betas = [q[0] for q in Wd[1] if len(q) == 1]
    lays_o = [lay + ((betas[i],) if i < 2 else ()) for i, lay in enumerate(lays)]
    aug = torch.cat([flat, x.reshape(-1), torch.zeros(1, dtype=torch.float64)])
    f = orc.augmented_post_drift(lays_o, t, aug, eps)
    assert torch.allclose(f[:orc.P], dw, rtol=1e-12, atol=1e-12)
    assert torch.allclose(f[orc.P:orc.P + B * D].reshape(B, D), dxt, rtol=1e-12, atol=1e-12)
    assert torch.allclose(f[-1], dkl, rtol=1e-12, atol=1e-12)
    gd = orc.augmented_diffusion(t, aug)
    assert torch.equal(gd[:orc.P], torch.full((orc.P,), 0.1, dtype=torch.float64)) and float(gd[orc.P:].abs().max()) == 0.0
    # STL: same value, different gradient w.r.t. the drift parameters, identical gradient path through w of the first term
    W = lays_o[-1][0].clone().requires_grad_(True)
    lays_g = lays_o[:-1] + [(W,) + lays_o[-1][1:]]
    kl_stl = orc.augmented_post_drift(lays_g, t, aug, eps)[-1]
    orc2 = jt.SDELayerOracle(B, D, H, "swish", jt.ToySetting(stop_grad=False))
    kl_full = orc2.augmented_post_drift(lays_g, t, aug, eps)[-1]
    assert torch.allclose(kl_stl, kl_full, rtol=1e-13)
    (g_stl,), (g_full,) = torch.autograd.grad(kl_stl, W), torch.autograd.grad(kl_full, W)
    u_ = orc2.augmented_post_drift(lays_g, t, aug, torch.zeros_like(eps))[-1]     # eps = 0: only sum(u^2)/2
    (g_u2,) = torch.autograd.grad(u_, W)
    assert torch.allclose(g_stl, g_u2, rtol=1e-10, atol=1e-12) and not torch.allclose(g_full, g_u2, rtol=1e-3)
    # prior solve: dw = -w, u = -w / sigma (layers.py:120-143)
    fp = orc.augmented_prior_drift(t, aug, eps)
    assert torch.allclose(fp[:orc.P], -flat) and torch.allclose(fp[-1], ((flat / 0.1) ** 2 / 2).sum() - (flat / 0.1 * eps).sum())

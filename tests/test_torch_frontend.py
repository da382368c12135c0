"""torch front end.

CPU: the oracle restatement (oracle/torch_path.py) against golden vectors produced by the REFERENCE's own `make_y_net` /
`make_w_net` (tests/golden/make_golden_torch_sdenet.py), fp64 finite differences of the oracle's solve, host logic.
GPU: `bayesde_b200.torch_models` (through libbsde.so) against the oracle on the same inputs and Brownian increments.

Tolerances: fp32 conv path -- drift net vs reference golden 2e-5 * max|ref|; y_T, w trajectory, logqp vs fp64 oracle
rtol 1e-4; gradients 1e-3 * max|ref|.  Tensor-core path (TF32 operands, 10-bit mantissa) -- y_T 5e-3 * max|ref|,
gradients 3e-2 * max|ref|.
"""
import math

import numpy as np
import pytest
import torch

from oracle import torch_path as tp

GOLD = np.load(__file__.rsplit("/", 1)[0] + "/golden/torch_sdenet.npz")
DEV = "cuda:0"


def _case(tag):
    cfg = [int(v) for v in GOLD[f"{tag}_cfg"]]
    c, h, w, nb = cfg[:4]
    blocks = tuple(cfg[4:4 + nb])
    hw, aug, B = cfg[4 + nb:]
    return (c, h, w), blocks, hw, aug, B, tuple(int(v) for v in GOLD[f"{tag}_wsizes"])


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_y_net_and_w_net(tag):
    """PIN: the oracle reproduces the reference's own make_y_net / make_w_net outputs bit for bit (same torch ops)."""
    insz, blocks, hw, aug, B, wsizes = _case(tag)
    ops, P, osz = tp.y_net_layout(insz, blocks, hw, aug)
    flat = torch.from_numpy(GOLD[f"{tag}_flat"])
    assert P == flat.numel() and tuple(osz) == tuple(GOLD[f"{tag}_out_size"])
    fy = tp.y_net_apply(ops, flat, float(GOLD[f"{tag}_t"]), torch.from_numpy(GOLD[f"{tag}_y"]))
    assert torch.equal(fy, torch.from_numpy(GOLD[f"{tag}_fy"]))
    layers = [(torch.from_numpy(GOLD[f"{tag}_wnet_W{i}"]), torch.from_numpy(GOLD[f"{tag}_wnet_b{i}"])) for i in range(len(wsizes) + 1)]
    nn_ = tp.w_net_apply(layers, flat)
    assert torch.equal(nn_, torch.from_numpy(GOLD[f"{tag}_nn"]))
    net = tp.SDENetOracle((insz[0] - 0, insz[1], insz[2]), blocks, wsizes, sigma=0.1, hidden_width=hw, aug_dim=aug)
    _, fw, fl, _ = net.f_parts(float(GOLD[f"{tag}_t"]), torch.from_numpy(GOLD[f"{tag}_y"]), flat, layers)
    assert torch.allclose(fw, torch.from_numpy(GOLD[f"{tag}_fw"]), rtol=0, atol=0)
    assert torch.allclose(fl, torch.from_numpy(GOLD[f"{tag}_fl"])[0], rtol=1e-6)


def _small_problem(dtype=torch.float64, seed=0, method="midpoint", nsteps=3, aug=1):
    g = torch.Generator().manual_seed(seed)
    net = tp.SDENetOracle((2, 8, 8), (1, 1), (1, 8, 1), sigma=0.3, hidden_width=4, aug_dim=aug)
    w0 = tp.init_y_params(net.ops, net.P, g, dtype)
    wl = tp.init_w_net(net.P, net.hidden_sizes, g, dtype, last_std=0.05)
    x = torch.randn(2, 2 + aug, 8, 8, generator=g, dtype=dtype)
    h = 1.0 / nsteps
    I = torch.randn(nsteps, net.P, generator=g, dtype=dtype) * h ** 0.5
    return net, w0, wl, x, I, h


def test_oracle_midpoint_properties():
    """sigma -> 0 with a zero w-net: w follows dw = -w exactly as the midpoint rule prescribes, logqp = 0; midpoint and the
    Euler-type schemes agree to O(h)."""
    net, w0, wl, x, I, h = _small_problem(nsteps=4)
    wl0 = [(W * 0, b * 0) for W, b in wl]
    y, w, lq = net.solve(x, w0, wl0, I * 0, dt=h, method="midpoint")
    assert torch.allclose(w, w0 * (1 - h + h * h / 2) ** 4, rtol=1e-12)
    assert lq.item() == 0.0
    with pytest.raises(ValueError, match="Unknown method"):
        net.solve(x, w0, wl, I, dt=h, method="srk")
    n = 64
    g = torch.Generator().manual_seed(1)
    If = torch.randn(n, net.P, generator=g, dtype=torch.float64) * (1.0 / n) ** 0.5
    ya, wa, la = net.solve(x, w0, wl, If, dt=1.0 / n, method="midpoint")
    yb, wb, lb = net.solve(x, w0, wl, If, dt=1.0 / n, method="euler_heun")
    assert (wa - wb).abs().max() < 5e-2 * wa.abs().max() and abs(la - lb) < 5e-2 * abs(la)
    # heun (trapezoidal): on dw = -w it has the same stability polynomial as midpoint; on the full problem the two second-order
    # schemes agree much more closely with each other than with the Euler-type ones
    yh, wh, lh = net.solve(x, w0, wl0, I * 0, dt=h, method="heun")
    assert torch.allclose(wh, w0 * (1 - h + h * h / 2) ** 4, rtol=1e-12) and lh.item() == 0.0
    yc, wc, lc = net.solve(x, w0, wl, If, dt=1.0 / n, method="heun")
    assert (wa - wc).abs().max() < 0.2 * (wa - wb).abs().max() and abs(la - lc) < 0.2 * abs(la - lb)


def test_oracle_gradient_finite_differences():
    """fp64 central differences of the oracle's loss along random directions (the analogue of jaxsde/tests/test_vjp.py:88-114)."""
    net, w0, wl, x, I, h = _small_problem()
    w0 = w0.clone().requires_grad_(True)

    def loss(w0_):
        y, _, lq = net.solve(x, w0_, wl, I, dt=h)
        return (y ** 2).sum() + 1e-3 * lq

    (g,) = torch.autograd.grad(loss(w0), w0)
    gen = torch.Generator().manual_seed(5)
    for _ in range(3):
        v = torch.randn(net.P, generator=gen, dtype=torch.float64)
        eps = 1e-6
        fd = (loss(w0.detach() + eps * v) - loss(w0.detach() - eps * v)) / (2 * eps)
        assert abs(fd - (g * v).sum()) <= 1e-5 * abs(fd) + 1e-8


def test_host_layout_matches_oracle_and_reference_pytree():
    from bayesde_b200 import torch_models as tm
    for insz, blocks, hw, aug in [((3, 32, 32), (2, 2, 2), 32, 0), ((2, 16, 16), (2, 2, 2), 6, 1), ((1, 8, 8), (1,), 4, 2)]:
        net, osz = tm.make_y_net(insz, blocks, hidden_width=hw, aug_dim=aug)
        ops, P, osz2 = tp.y_net_layout(insz, blocks, hw, aug)
        assert net.P == P and tuple(osz) == tuple(osz2)
        convs = [d for k, d in ops if k == "conv"]
        assert [(l["w_off"], l["b_off"], l["shape"]) for l in net.layers] == [(d["w_off"], d["b_off"], d["shape"]) for d in convs]
        flat, unravel = tm.ravel_pytree(net.make_initial_params(torch.Generator().manual_seed(0)))
        assert flat.numel() == P
        tree = unravel(flat)
        assert [tuple(e[0].shape) for e in tree if len(e)] == [d["shape"] for d in convs]
    assert tm.make_y_net((3, 32, 32), hidden_width=32)[0].P == 196296
    with pytest.raises(AssertionError, match="divisibility"):
        tm.make_y_net((1, 28, 28), (2, 2, 2), hidden_width=4)                    # 7x7 third group (quirk Q10 analogue)
    cfg = tm.SolveConfig(net, 1, 2, 20, 0.05, "midpoint", 0.1, (1, 8, 1), "tanh", "fp32", False)
    assert cfg.nit == 40 and cfg.back.tolist()[:4] == [0, 1, 0, 1] and abs(cfg.tk[3] - 0.075) < 1e-7 and cfg.dtkl[0] == 0
    with pytest.raises(ValueError, match="Unknown method"):
        tm.SolveConfig(net, 1, 2, 20, 0.05, "srk", 0.1, (1, 8, 1), "tanh", "fp32", False)


def test_sdenet_module_surface(built_lib):
    from bayesde_b200 import _lib, torch_models as tm
    m = tm.SDENet(input_size=(2, 8, 8), blocks=(1, 1), weight_network_sizes=(1, 8, 1), hidden_width=4, aug_dim=1)
    keys = set(m.state_dict().keys())
    assert {"flat_initial_params", "w_net.layers.0.weight", "w_net.layers.4.bias", "projection.1.weight", "projection.3.bias",
            "ts", "aug_zeros"} <= keys
    assert m.params_size == m.y_net.P and float(m.w_net.linears()[-1].weight.abs().max()) == 0.0
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 8, 8), adjoint_adaptive=True)      # needs adjoint=True
    with pytest.raises(_lib.BsdeError):
        m(torch.zeros(1, 2, 8, 8), adaptive=True)
    with pytest.raises(_lib.BsdeError):
        m(torch.zeros(1, 2, 8, 8))     # CPU tensor: no fallback
    with pytest.raises(_lib.BsdeError):
        m(torch.zeros(1, 2, 8, 8), adjoint=True)


# ------------------------------------------------------------------------------------------------ GPU
def _close(a, b, tol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * ref + 1e-12, f"{what}: max err {err:.3e} vs ref max {ref:.3e} (tol {tol:g})"


@pytest.mark.gpu
@pytest.mark.parametrize("math_mode,tol", [("fp32", 2e-5), ("tc", 5e-3)])
@pytest.mark.parametrize("tag", ["a", "b"])
def test_gpu_y_net_matches_reference_golden(built_lib, tag, math_mode, tol):
    """bsde_chain_fwd (all groups + ConvDownsample, OIHW / IOHW weights read in place, time channel first) against the
    REFERENCE's own make_y_net output."""
    from bayesde_b200 import torch_models as tm
    insz, blocks, hw, aug, B, _ = _case(tag)
    net, _ = tm.make_y_net(insz, blocks, hidden_width=hw, aug_dim=aug)
    fy = net(float(GOLD[f"{tag}_t"]), torch.from_numpy(GOLD[f"{tag}_y"]).to(DEV), torch.from_numpy(GOLD[f"{tag}_flat"]).to(DEV), math_mode)
    _close(fy, torch.from_numpy(GOLD[f"{tag}_fy"]), tol, f"y_net {tag} {math_mode}")


def _run_gpu(net, w0, wl, x, I, h, method, math_mode, remat, seed_loss=3):
    from bayesde_b200 import torch_models as tm
    ynet, _ = tm.make_y_net(net.input_size, _run_gpu.blocks, activation=net.activation, hidden_width=_run_gpu.hw, aug_dim=net.aug_dim)
    assert ynet.P == net.P
    B = x.shape[0]
    cfg = tm.SolveConfig(ynet, 1, B, I.shape[0], h, method, net.sigma, net.hidden_sizes, "tanh", math_mode, remat)
    xd = x.float().permute(0, 2, 3, 1).contiguous()[None].to(DEV).requires_grad_(True)
    w0d = w0.float().to(DEV).requires_grad_(True)
    lin = [(W.float().to(DEV).requires_grad_(True), b.float().to(DEV).requires_grad_(True)) for W, b in wl]
    y1, lq, wtraj = tm.sdenet_solve(cfg, xd, w0d, lin, I.float().to(DEV)[None])
    y1n = y1[0].permute(0, 3, 1, 2)
    g = torch.Generator().manual_seed(seed_loss)
    cot = torch.randn(y1n.shape, generator=g).to(DEV)
    loss = (y1n * cot).sum() + 0.7 * lq[0]
    grads = torch.autograd.grad(loss, [xd, w0d] + [t for p in lin for t in p])
    return y1n, lq[0], wtraj[0], grads, cot


@pytest.mark.gpu
@pytest.mark.parametrize("method,math_mode,remat", [("midpoint", "fp32", False), ("midpoint", "fp32", True), ("euler_heun", "fp32", False),
                                                   ("midpoint", "tc", False), ("milstein", "tc", True),
                                                   ("heun", "fp32", False), ("heun", "fp32", True), ("heun", "tc", False)])
def test_gpu_sdenet_solve_matches_oracle(built_lib, method, math_mode, remat):
    """State, w trajectory (all stages), logqp and every gradient of the fixed-step solve against the fp64 oracle with the
    same increments: two groups (ConvDownsample in the drift), aug channel, P -> 2 -> 8 -> 3 -> P w-net."""
    g = torch.Generator().manual_seed(11)
    blocks, hw = (1, 1), 8
    net = tp.SDENetOracle((2, 8, 8), blocks, (2, 8, 3), sigma=0.3, hidden_width=hw, aug_dim=1)
    _run_gpu.blocks, _run_gpu.hw = blocks, hw
    dt64 = torch.float64
    w0 = tp.init_y_params(net.ops, net.P, g, dt64).requires_grad_(True)
    wl = [(W.requires_grad_(True), b.requires_grad_(True)) for W, b in tp.init_w_net(net.P, net.hidden_sizes, g, dt64, last_std=0.05)]
    x = torch.randn(3, 3, 8, 8, generator=g, dtype=dt64).requires_grad_(True)
    n = 4
    h = 1.0 / n
    I = torch.randn(n, net.P, generator=g, dtype=dt64) * h ** 0.5
    y1n, lq, wtraj, grads, cot = _run_gpu(net, w0.detach(), [(W.detach(), b.detach()) for W, b in wl], x.detach(), I, h, method, math_mode, remat)
    yr, wr, lr, path = net.solve(x, w0, wl, I, dt=h, method=method, return_path=True)
    loss = (yr * cot.double().cpu()).sum() + 0.7 * lr
    gr = torch.autograd.grad(loss, [x, w0] + [t for p in wl for t in p])
    ty, tg = (1e-4, 1e-3) if math_mode == "fp32" else (5e-3, 3e-2)
    _close(wtraj, path, 1e-5, "w trajectory (all stages)")
    _close(lq, lr, 1e-4, "logqp state")
    _close(y1n, yr, ty, "y_T")
    _close(grads[0][0].permute(0, 3, 1, 2), gr[0], tg, "dL/dx")
    names = ["w0"] + [f"w_net[{i // 2}].{'weight' if i % 2 == 0 else 'bias'}" for i in range(2 * len(wl))]
    for nme, a, b in zip(names, grads[1:], gr[1:]):
        _close(a, b, tg, f"dL/d{nme}")


@pytest.mark.gpu
def test_gpu_sdenet_module_forward_backward(built_lib):
    """The nn.Module surface: forward(y, dt, method) -> (logits, logqp) with replayed increments against the oracle's forward,
    Philox path determinism, nfe bookkeeping, gradients reach every parameter."""
    from bayesde_b200 import torch_models as tm
    torch.manual_seed(0)
    m = tm.SDENet(input_size=(2, 8, 8), blocks=(1, 1), weight_network_sizes=(1, 8, 1), hidden_width=4, aug_dim=1, sigma=0.2, math="fp32").to(DEV)
    with torch.no_grad():
        last = m.w_net.linears()[-1]
        last.weight.normal_(0, 0.05), last.bias.normal_(0, 0.05)
    x = torch.randn(3, 2, 8, 8, device=DEV)
    n = 5
    I = torch.randn(n, m.params_size, device=DEV) * (1.0 / n) ** 0.5
    logits, logqp = m(x, dt=1.0 / n, method="midpoint", increments=I)
    assert m.nfe == 4 * n and logits.shape == (3, 10)
    net = tp.SDENetOracle((2, 8, 8), (1, 1), (1, 8, 1), sigma=0.2, hidden_width=4, aug_dim=1)
    dd = lambda t: t.detach().double().cpu()
    proj = [(dd(m.projection[1].weight), dd(m.projection[1].bias)), (dd(m.projection[3].weight), dd(m.projection[3].bias))]
    lo, lq = net.forward(dd(x), dd(m.flat_initial_params), [(dd(l.weight), dd(l.bias)) for l in m.w_net.linears()], proj, dd(I), dt=1.0 / n)
    _close(logits, lo, 1e-4, "logits")
    _close(logqp, lq, 1e-4, "logqp")
    (logits.logsumexp(1).mean() + 1e-3 * logqp).backward()
    for nme, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0, nme
    # in-kernel Philox: same (seed, call) -> same result; next call -> fresh noise
    m._calls = 0
    a1, q1 = m(x, dt=0.25)
    m._calls = 0
    a2, q2 = m(x, dt=0.25)
    a3, q3 = m(x, dt=0.25)
    assert torch.equal(a1, a2) and torch.equal(q1, q2) and not torch.equal(q1, q3)


@pytest.mark.gpu
def test_gpu_nchw_reinterp_roundtrip(built_lib):
    """bsde_nchw_reinterp_axpy against torch's reshape on NCHW tensors, and <R x, y> == <x, R^T y>."""
    from bayesde_b200 import torch_models as tm
    B, chw, chw2 = 3, (3, 8, 8), (12, 4, 4)
    f = torch.randn(B, *chw2, device=DEV)
    y = torch.randn(B, *chw, device=DEV)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    out = torch.empty(B, 8, 8, 3, device=DEV)
    tm.reinterp_axpy(out, chw, nhwc(y), 0.5, nhwc(f), chw2)
    assert torch.equal(out.permute(0, 3, 1, 2), y + 0.5 * f.reshape(B, *chw))
    back = torch.empty(B, 4, 4, 12, device=DEV)
    tm.reinterp_axpy(back, chw2, None, 1.0, nhwc(y), chw)
    assert torch.equal(back.permute(0, 3, 1, 2), y.reshape(B, *chw2))


@pytest.mark.gpu
def test_gpu_config5_full_size_properties(built_lib):
    """Config 5 at full size (3x32x32, blocks 2-2-2, hidden 32, P = 196 296, w-net 1-128-1, midpoint dt 0.05) through properties
    that do not need the oracle at that size: determinism of the Philox path, tensor-core vs fp32 conv agreement (TF32 tolerance),
    remat == stored activations bit for bit, batch independence (images [0:B/2] of a B solve equal a B/2 solve on them), and the
    w / logqp rows do not depend on y."""
    from bayesde_b200 import torch_models as tm
    torch.manual_seed(0)
    B = 16
    m = tm.SDENet(input_size=(3, 32, 32), blocks=(2, 2, 2), weight_network_sizes=(1, 128, 1), hidden_width=32, sigma=0.1, math="tc").to(DEV)
    assert m.params_size == 196296
    with torch.no_grad():
        last = m.w_net.linears()[-1]
        last.weight.normal_(0, 1e-2), last.bias.normal_(0, 1e-2)
    x = torch.randn(B, 3, 32, 32, device=DEV)

    def run(xx, math_mode="tc", remat=False):
        m.math, m.remat, m._calls = math_mode, remat, 0
        m.zero_grad()
        logits, logqp = m(xx, dt=0.05, method="midpoint")
        (logits.square().mean() + 1e-6 * logqp).backward()
        return logits.detach(), logqp.detach(), m.flat_initial_params.grad.clone(), m.w_net.linears()[0].weight.grad.clone()

    a = run(x)
    b = run(x)
    for u, v in zip(a, b):
        assert torch.equal(u, v), "not deterministic"
    assert all(torch.isfinite(t).all() for t in a) and m.nfe == 80
    r = run(x, remat=True)
    for u, v in zip(a, r):
        assert torch.equal(u, v), "remat differs from stored activations"
    f = run(x, math_mode="fp32")
    _close(a[0], f[0], 5e-3, "logits tc vs fp32")
    assert torch.equal(a[1], f[1])                      # the [w, logqp] rows never touch the convs
    _close(a[2], f[2], 3e-2, "dL/dw0 tc vs fp32")
    h = run(x[:B // 2].contiguous())
    # the projection head is a cuBLAS GEMM whose reduction order depends on the batch: allclose, not equal
    assert torch.allclose(h[0], a[0][:B // 2], rtol=1e-5, atol=1e-6) and torch.equal(h[1], a[1])
    z = run(torch.zeros_like(x))
    assert torch.equal(z[1], a[1])


# ---------------------------------------------------------------------------------------------------------------
# `adjoint=True` (torchsde.sdeint_adjoint, models.py:194) -- O(1)-memory reverse pass
# ---------------------------------------------------------------------------------------------------------------
def _adjoint_case(seed=11):
    g = torch.Generator().manual_seed(seed)
    net = tp.SDENetOracle((2, 8, 8), (1, 1), (2, 8, 3), sigma=0.3, hidden_width=8, aug_dim=1)
    dt64 = torch.float64
    w0 = tp.init_y_params(net.ops, net.P, g, dt64)
    wl = tp.init_w_net(net.P, net.hidden_sizes, g, dt64, last_std=0.05)
    x = torch.randn(3, 3, 8, 8, generator=g, dtype=dt64)
    cot = torch.randn(3, 3, 8, 8, generator=g, dtype=dt64)
    fine = torch.randn(32, net.P, generator=g, dtype=dt64) * (1.0 / 32) ** 0.5      # one Brownian path, coarsened per step count
    return net, w0, wl, x, cot, fine


def test_oracle_sdenet_adjoint_converges_to_backprop():
    """The restated `sdeint_adjoint` reverse pass against autograd through the oracle's own forward solve on one Brownian path:
    midpoint reconstructs the initial state to third order and the gradients to second order in h, the Euler-type schemes to first
    order (measured: y0_rec 2.7e-2 / 3.5e-3 / 4.3e-4 and dL/dw0 5.4e-2 / 1.1e-2 / 2.4e-3 relative at n = 4 / 8 / 16, midpoint)."""
    net, w0, wl, x, cot, fine = _adjoint_case()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    for method, order_rec, order_g in (("midpoint", 3, 2), ("heun", 2, 2), ("euler_heun", 1, 1)):
        errs = []
        for n in (8, 16):
            h = 1.0 / n
            I = fine.reshape(n, 32 // n, net.P).sum(1)
            xr, w0r = x.clone().requires_grad_(True), w0.clone().requires_grad_(True)
            wlr = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in wl]
            yr, wr, lr = net.solve(xr, w0r, wlr, I, dt=h, method=method)
            gr = torch.autograd.grad((yr * cot).sum() + 0.7 * lr, [xr, w0r] + [t for p in wlr for t in p])
            y0, w0_, ay, aw, ap = net.adjoint(yr.detach(), wr.detach(), cot, torch.zeros_like(w0), torch.tensor(0.7, dtype=x.dtype),
                                              wl, I, dt=h, method=method)
            errs.append((float((y0 - x).abs().max()), rel(aw, gr[1]), max(rel(a, b) for a, b in zip(ap, gr[2:]))))
        (r0, g0, p0), (r1, g1, p1) = errs
        assert r0 / r1 > 0.7 * 2 ** order_rec and g0 / g1 > 0.7 * 2 ** order_g and p0 / p1 > 0.7 * 2 ** order_g, (method, errs)
        if method == "midpoint":
            assert r1 < 1e-3 and g1 < 5e-3 and p1 < 5e-3, errs


@pytest.mark.gpu
@pytest.mark.parametrize("method,math_mode", [("midpoint", "fp32"), ("euler_heun", "fp32"), ("midpoint", "tc"), ("heun", "fp32")])
def test_gpu_sdenet_adjoint_matches_oracle(built_lib, method, math_mode):
    """sdenet_solve_adjoint against the oracle's restated reverse pass with the same increments: forward values equal the plain route,
    reconstructed initial state and every adjoint gradient 1e-3 (fp32) / 3e-2 (TF32) of the largest entry."""
    from bayesde_b200 import torch_models as tm
    net, w0, wl, x, cot, fine = _adjoint_case()
    n = 8
    h = 1.0 / n
    I = fine.reshape(n, 32 // n, net.P).sum(1)
    ynet, _ = tm.make_y_net(net.input_size, (1, 1), activation=net.activation, hidden_width=8, aug_dim=net.aug_dim)
    cfg = tm.SolveConfig(ynet, 1, x.shape[0], n, h, method, net.sigma, net.hidden_sizes, "tanh", math_mode, False)
    xd = x.float().permute(0, 2, 3, 1).contiguous()[None].to(DEV).requires_grad_(True)
    w0d = w0.float().to(DEV).requires_grad_(True)
    lin = [(W.float().to(DEV).requires_grad_(True), b.float().to(DEV).requires_grad_(True)) for W, b in wl]
    y1, lq, wtraj = tm.sdenet_solve_adjoint(cfg, xd, w0d, lin, I.float().to(DEV)[None])
    with torch.no_grad():
        y1p, lqp, _ = tm.sdenet_solve(cfg, xd, w0d, lin, I.float().to(DEV)[None])
    assert torch.equal(y1, y1p) and torch.equal(lq, lqp)
    y1n = y1[0].permute(0, 3, 1, 2)
    loss = (y1n * cot.float().to(DEV)).sum() + 0.7 * lq[0]
    grads = torch.autograd.grad(loss, [xd, w0d] + [t for p in lin for t in p])
    yr, wr, lr = net.solve(x, w0, wl, I, dt=h, method=method)
    y0, w0r, ay, aw, ap = net.adjoint(yr, wr, cot, torch.zeros_like(w0), torch.tensor(0.7, dtype=x.dtype), wl, I, dt=h, method=method)
    tol = 1e-3 if math_mode == "fp32" else 3e-2
    Zy, Zw = tm.LAST_RECONSTRUCTION
    _close(Zy[0].permute(0, 3, 1, 2), y0, tol, "reconstructed y0")
    _close(Zw, w0r, tol, "reconstructed w0")
    _close(grads[0][0].permute(0, 3, 1, 2), ay, tol, "dL/dx")
    _close(grads[1], aw, tol, "dL/dw0")
    for i, (a, b) in enumerate(zip(grads[2:], ap)):
        _close(a, b, tol, f"dL/d w_net tensor {i}")


@pytest.mark.gpu
def test_gpu_sdenet_module_adjoint_flag(built_lib):
    """SDENet.forward(adjoint=True): same logits / logqp as the plain route on the same path (replayed and in-kernel Philox), gradients
    within the O(h^2) gap of the two reverse passes."""
    from bayesde_b200 import torch_models as tm
    torch.manual_seed(0)
    m = tm.SDENet(input_size=(2, 8, 8), blocks=(1, 1), weight_network_sizes=(1, 8, 1), hidden_width=4, aug_dim=1, sigma=0.2, math="fp32").to(DEV)
    with torch.no_grad():
        last = m.w_net.linears()[-1]
        last.weight.normal_(0, 0.05), last.bias.normal_(0, 0.05)
    x = torch.randn(3, 2, 8, 8, device=DEV)
    n = 16
    outs = []
    for adjoint in (False, True):
        m.zero_grad()
        m._calls = 0
        logits, logqp = m(x, dt=1.0 / n, method="midpoint", adjoint=adjoint)      # in-kernel Philox path of call 1 in both
        (logits.logsumexp(1).mean() + 1e-3 * logqp).backward()
        outs.append((logits.detach(), logqp.detach(), {k: p.grad.clone() for k, p in m.named_parameters()}))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-5) and torch.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-6)
    for k in outs[0][2]:
        a, b = outs[1][2][k], outs[0][2][k]
        assert float((a - b).abs().max()) <= 2e-2 * float(b.abs().max()) + 1e-8, (k, float((a - b).abs().max()), float(b.abs().max()))
    with pytest.raises(ValueError):
        m(x, adjoint_adaptive=True)                                                # needs adjoint=True


# ---------------------------------------------------------------------------------------------------------------
# `adaptive=True` -- torchsde's step-doubling controller, accepted grid replayed for the gradients
# ---------------------------------------------------------------------------------------------------------------
def test_step_size_controller_known_values():
    """Hand-evaluated values of the PI controller (safety 0.9, I exponent 1/4.5 on acceptance and 1/1.5 on rejection, P exponent 0.13,
    factor clamped to [0.2 or 1, 1.4]); product and oracle restatements agree."""
    from bayesde_b200 import torch_models as tm
    for fn in (tm.update_step_size, tp.SDENetOracle.update_step_size):
        s, r = fn(0.5, 0.1)
        assert abs(s - 0.1 * 1.8 ** (1 / 4.5)) < 1e-15 and r == 1.8
        s, r = fn(2.0, 0.1, prev_error_ratio=1.8)
        assert abs(s - 0.1 * 0.45 ** (1 / 1.5)) < 1e-15 and r == 1.8                  # rejected: prev ratio kept, no P term
        assert abs(fn(1000.0, 0.1)[0] - 0.02) < 1e-15 and abs(fn(1e-3, 0.1)[0] - 0.14) < 1e-15      # clamps 0.2 and 1.4
        s, r = fn(0.9, 0.1, prev_error_ratio=2.0)
        assert abs(s - 0.1 * max(1.0, 1.0 ** (1 / 4.5) * (1.0 / 2.0) ** 0.13)) < 1e-15 and r == 1.0    # accepted: factor floor 1
    g = torch.Generator().manual_seed(0)
    a = [torch.randn(5, 7, generator=g), torch.randn(11, generator=g), torch.randn((), generator=g)]
    b = [t + 1e-3 * torch.randn(t.shape, generator=g) for t in a]
    assert abs(tm.compute_error(a, b, 1e-3, 1e-4) - tp.SDENetOracle.compute_error(a, b, 1e-3, 1e-4)) < 1e-9
    assert tm.compute_error(a, a, 1e-3, 1e-4) == 1e-7


def test_adaptive_loop_host_logic():
    """The product's loop on a scalar test problem (explicit midpoint on y' = -3 y): reaches t1 exactly, the accepted half steps tile
    [t0, t1], rejections occur when the first step is too large, tighter tolerances take more steps and land closer to exp(-3)."""
    from bayesde_b200 import torch_models as tm

    def one_step(ta, h, state):
        (y,) = state
        h = float(h)
        return (y + h * (-3.0) * (y + 0.5 * h * (-3.0) * y),)

    res = []
    for tol in (1e-2, 1e-4):
        ts, hs, (y,), nsingle = tm.sdenet_adaptive_grid(one_step, 0.0, 1.0, 0.5, tol, tol, (torch.ones(3, dtype=torch.float64),))
        assert len(ts) == len(hs) and len(ts) % 2 == 0 and nsingle >= 3 * len(ts) // 2
        assert ts[0] == 0.0 and all(abs(float(ts[i] + hs[i]) - float(ts[i + 1])) < 1e-6 for i in range(len(ts) - 1))
        assert abs(float(ts[-1] + hs[-1]) - 1.0) < 1e-6
        res.append((len(ts), nsingle, abs(float(y[0]) - math.exp(-3.0))))
    assert res[1][0] > res[0][0] and res[1][2] < res[0][2] and res[1][2] < 1e-3
    assert res[1][1] > 3 * res[1][0] // 2                     # some trials were rejected
    with pytest.raises(RuntimeError):
        tm.sdenet_adaptive_grid(one_step, 0.0, 1.0, 0.5, 1e-4, 1e-4, (torch.ones(1),), max_trials=2)


def test_oracle_adaptive_converges_and_replays():
    from oracle import jax_adjoint as ja
    net, w0, wl, x, cot, fine = _adjoint_case()
    tree = ja.BrownianTree(7, 1, net.P, 0.0, 1.0)
    bm = lambda t: tree(t)
    outs = []
    for tol in (1e-2, 1e-3):
        ts, hs, errs, (y, w, lq) = net.solve_adaptive(x, w0, wl, bm, 0.25, tol, tol)
        yr, wr, lr = net.solve_on_grid(x, w0, wl, bm, ts, hs)
        assert torch.equal(y, yr) and torch.equal(w, wr) and torch.equal(lq, lr)          # the replay IS the accepted solve
        assert abs(float(sum(float(h) for h in hs)) - 1.0) < 1e-5 and max(errs) > 1.0 > min(errs)
        outs.append((len(ts), y, lq))
    n = 256
    I = torch.stack([bm((k + 1) / n) - bm(k / n) for k in range(n)])
    yf, wf, lf = net.solve(x, w0, wl, I, dt=1.0 / n, method="midpoint")
    e = [float((o[1] - yf).abs().max()) for o in outs]
    assert outs[1][0] > outs[0][0] and e[1] < e[0] and abs(float(outs[1][2] - lf)) < abs(float(outs[0][2] - lf)) + 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["midpoint", "heun"])
def test_gpu_sdenet_adaptive(built_lib, method):
    """SDENet.forward(adaptive=True): valid accepted grid; logits / logqp and every gradient against the oracle's fixed-step solve ON THE
    PRODUCT'S accepted grid with the same Philox tree (1e-4 / 1e-3); the oracle's own adaptive loop takes a comparable number of steps
    (accept / reject decisions sit near error = 1 by construction, so fp32 and fp64 grids need not coincide); deterministic per call."""
    from bayesde_b200 import torch_models as tm
    from oracle import jax_adjoint as ja
    torch.manual_seed(0)
    m = tm.SDENet(input_size=(2, 8, 8), blocks=(1, 1), weight_network_sizes=(1, 8, 1), hidden_width=4, aug_dim=1, sigma=0.2, math="fp32").to(DEV)
    with torch.no_grad():
        last = m.w_net.linears()[-1]
        last.weight.normal_(0, 0.05), last.bias.normal_(0, 0.05)
    x = torch.randn(3, 2, 8, 8, device=DEV)
    m._calls = 0
    logits, logqp = m(x, dt=0.25, method=method, adaptive=True, rtol=1e-3, atol=1e-3)
    ts, hs = m.adaptive_grid
    assert len(ts) >= 4 and abs(sum(float(h) for h in hs) - 1.0) < 1e-5 and m.nfe >= 4 * len(ts)
    (logits.logsumexp(1).mean() + 1e-3 * logqp).backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}

    net = tp.SDENetOracle((2, 8, 8), (1, 1), (1, 8, 1), sigma=0.2, hidden_width=4, aug_dim=1)
    dd = lambda t: t.detach().double().cpu()
    tree = ja.BrownianTree(m.seed, 1, net.P, 0.0, 1.0)
    bm = lambda t: tree(t)
    w0 = dd(m.flat_initial_params).requires_grad_(True)
    wl = [(dd(l.weight).requires_grad_(True), dd(l.bias).requires_grad_(True)) for l in m.w_net.linears()]
    xa = torch.cat((dd(x), torch.zeros(3, 1, 8, 8, dtype=torch.float64)), dim=1)
    yr, wr, lr = net.solve_on_grid(xa, w0, wl, bm, ts, hs, method=method)
    (W1, b1), (W2, b2) = (dd(m.projection[1].weight), dd(m.projection[1].bias)), (dd(m.projection[3].weight), dd(m.projection[3].bias))
    lo = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(yr.reshape(3, -1), W1, b1)), W2, b2)
    _close(logits, lo, 1e-4, "logits on the accepted grid")
    _close(logqp, 0.5 * lr, 1e-4, "logqp on the accepted grid")
    gr = torch.autograd.grad(lo.logsumexp(1).mean() + 1e-3 * 0.5 * lr, [w0] + [t for p in wl for t in p])
    _close(grads["flat_initial_params"], gr[0], 1e-3, "dL/dw0")
    for i, l in enumerate(m.w_net.linears()):
        name = [k for k, p in m.named_parameters() if p is l.weight][0]
        _close(grads[name], gr[1 + 2 * i], 1e-3, f"dL/d w_net[{i}].weight")
    ts_o, hs_o, errs_o, _ = net.solve_adaptive(xa, w0.detach(), [(W.detach(), b.detach()) for W, b in wl], bm, 0.25, 1e-3, 1e-3, method=method)
    assert 0.6 * len(ts_o) <= len(ts) <= 1.6 * len(ts_o), (len(ts), len(ts_o))
    m._calls = 0
    l2, q2 = m(x, dt=0.25, method=method, adaptive=True, rtol=1e-3, atol=1e-3)
    assert torch.equal(l2, logits) and torch.equal(q2, logqp)
    # adaptive forward + sdeint_adjoint's fixed-step reverse solve (adjoint_adaptive=False) from the adaptive solve's final state
    m.zero_grad()
    m._calls = 0
    dt = 0.125
    l3, q3 = m(x, dt=dt, method=method, adaptive=True, adjoint=True, rtol=1e-3, atol=1e-3)
    ts3, hs3 = m.adaptive_grid
    (l3.logsumexp(1).mean() + 1e-3 * q3).backward()
    y3, w3, lr3 = net.solve_on_grid(xa, w0.detach(), [(W.detach(), b.detach()) for W, b in wl], bm, ts3, hs3, method=method)
    yl = y3.clone().requires_grad_(True)
    lo3 = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(yl.reshape(3, -1), W1, b1)), W2, b2)
    _close(l3, lo3, 1e-4, "logits (adaptive + adjoint)")
    (cot_y,) = torch.autograd.grad(lo3.logsumexp(1).mean(), yl)
    f32 = np.float32
    tu = [f32(0.0) + f32(k) * f32(dt) for k in range(8)]
    Iu = torch.stack([bm(float(f32(t + f32(dt)))) - bm(float(t)) for t in tu])
    _, _, ay, aw, ap = net.adjoint(y3, w3, cot_y, torch.zeros_like(w3), torch.tensor(0.5e-3, dtype=torch.float64),
                                   [(W.detach(), b.detach()) for W, b in wl], Iu, dt=dt, method=method)
    _close(m.flat_initial_params.grad, aw, 1e-3, "dL/dw0 (adaptive + adjoint)")
    for i, l in enumerate(m.w_net.linears()):
        _close(l.weight.grad, ap[2 * i], 1e-3, f"dL/d w_net[{i}].weight (adaptive + adjoint)")
        _close(l.bias.grad, ap[2 * i + 1], 1e-3, f"dL/d w_net[{i}].bias (adaptive + adjoint)")
    # adjoint_adaptive=True: the controller also drives the reverse solve; oracle reverse pass ON THE PRODUCT'S accepted reverse grid
    m.zero_grad()
    m._calls = 0
    l4, q4 = m(x, dt=dt, method=method, adaptive=True, adjoint=True, adjoint_adaptive=True, rtol=1e-3, atol=1e-3)
    assert torch.equal(l4, l3) and torch.equal(q4, q3)
    (l4.logsumexp(1).mean() + 1e-3 * q4).backward()
    ts_r, hs_r = tm.LAST_REVERSE_GRID
    assert len(ts_r) >= 4 and abs(float(ts_r[0]) + 1.0) < 1e-6 and abs(sum(float(h) for h in hs_r) - 1.0) < 1e-5
    y0r, w0r, ay, aw, ap = net.adjoint(y3, w3, cot_y, torch.zeros_like(w3), torch.tensor(0.5e-3, dtype=torch.float64),
                                       [(W.detach(), b.detach()) for W, b in wl], None, dt=dt, method=method, reverse_grid=(bm, ts_r, hs_r))
    _close(m.flat_initial_params.grad, aw, 1e-3, "dL/dw0 (adjoint_adaptive)")
    for i, l in enumerate(m.w_net.linears()):
        _close(l.weight.grad, ap[2 * i], 1e-3, f"dL/d w_net[{i}].weight (adjoint_adaptive)")
    Zy, Zw = tm.LAST_RECONSTRUCTION
    _close(Zw, w0r, 1e-3, "reconstructed w0 (adjoint_adaptive)")
    # the reverse solve reconstructs w0 only up to the discretisation of both adaptive solves (different grids, sigma = 0.2)
    assert float((Zw.double().cpu() - dd(m.flat_initial_params)).abs().max()) < 0.15 * float(dd(m.flat_initial_params).abs().max())
    with pytest.raises(ValueError):
        m(x, adjoint_adaptive=True)


def test_solve_config_tables_for_explicit_grids(built_lib):
    """Host logic: `SolveConfig(grid=(t_k, h_k))` gives the same stage tables as the uniform constructor when the grid is uniform, and the
    per-step entries follow each scheme's definition on a non-uniform grid (midpoint: (t, t + h/2), dt (h/2, h), KL step (0, h), noise
    (sqrt(h)/2, sqrt(h)), back (0, 1); heun: (t, t + h), (h, h/2), (h/2, h/2), (sqrt(h), sqrt(h)/2), back (0, 2); Euler-type: no tables)."""
    from bayesde_b200 import torch_models as tm
    ynet, _ = tm.make_y_net((2, 8, 8), (1, 1), hidden_width=4, aug_dim=1)
    f32 = np.float32
    for method in ("midpoint", "heun", "milstein"):
        u = tm.SolveConfig(ynet, 1, 3, 4, 0.25, method, 0.1, (1, 8, 1), "tanh", "fp32", False)
        g = tm.SolveConfig(ynet, 1, 3, 0, 0.0, method, 0.1, (1, 8, 1), "tanh", "fp32", False,
                           grid=([f32(0.25) * f32(k) for k in range(4)], [f32(0.25)] * 4))
        assert u.uniform and not g.uniform and u.n == g.n == 4 and u.nit == g.nit
        for name in ("tk", "dtk", "dtkl", "nscale", "nidx", "back"):
            a, b = getattr(u, name), getattr(g, name)
            assert (a is None and b is None) or np.array_equal(a, b), (method, name)
    ts, hs = [0.0, 0.1, 0.4], [0.1, 0.3, 0.6]
    m = tm.SolveConfig(ynet, 1, 3, 0, 0.0, "midpoint", 0.1, (1, 8, 1), "tanh", "fp32", False, grid=(ts, hs))
    assert np.allclose(m.tk, [0.0, 0.05, 0.1, 0.25, 0.4, 0.7]) and np.allclose(m.dtk, [0.05, 0.1, 0.15, 0.3, 0.3, 0.6])
    assert np.allclose(m.dtkl, [0, 0.1, 0, 0.3, 0, 0.6]) and m.back.tolist() == [0, 1] * 3 and m.nidx.tolist() == [0, 0, 1, 1, 2, 2]
    assert np.allclose(m.nscale, [0.5 * 0.1 ** 0.5, 0.1 ** 0.5, 0.5 * 0.3 ** 0.5, 0.3 ** 0.5, 0.5 * 0.6 ** 0.5, 0.6 ** 0.5])
    h = tm.SolveConfig(ynet, 1, 3, 0, 0.0, "heun", 0.1, (1, 8, 1), "tanh", "fp32", False, grid=(ts, hs))
    assert np.allclose(h.tk, [0.0, 0.1, 0.1, 0.4, 0.4, 1.0]) and np.allclose(h.dtk, [0.1, 0.05, 0.3, 0.15, 0.6, 0.3])
    assert np.allclose(h.dtkl, [0.05, 0.05, 0.15, 0.15, 0.3, 0.3]) and h.back.tolist() == [0, 2] * 3
    I = torch.arange(6.0).reshape(1, 3, 2)
    assert torch.equal(m.stage_increments(I)[0, :, 0], torch.tensor([0.0, 0.0, 1.0, 2.0, 2.0, 4.0]))
    assert torch.equal(h.stage_increments(I)[0, :, 0], torch.tensor([0.0, 0.0, 2.0, 1.0, 4.0, 2.0]))
    e = tm.SolveConfig(ynet, 1, 3, 0, 0.0, "euler_heun", 0.1, (1, 8, 1), "tanh", "fp32", False, grid=(ts, hs))
    assert e.nit == 3 and e.back is None and np.allclose(e.dtk, hs) and e.hs == [float(f32(v)) for v in hs]
    with pytest.raises(ValueError, match="Unknown method"):
        tm.SolveConfig(ynet, 1, 3, 4, 0.25, "srk", 0.1, (1, 8, 1), "tanh", "fp32", False)

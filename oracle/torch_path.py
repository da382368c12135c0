"""CPU oracle for the torch front end.  TEST INFRASTRUCTURE ONLY.

Restates, with plain torch CPU ops (fp64 or fp32, autograd for the gradients):
  * the drift network on the hidden state, `make_y_net` (torchbnn/_impl/models.py:25-54) built from
    `make_ode_k3_block_layers` (torchbnn/_impl/diffeq_layers.py:280-311), `ConcatConv2d` (:75-119),
    `ConcatConvTranspose2d` (:122-166), `ConvDownsample` (:230-239), `channel_cat` (torchbnn/_impl/utils.py:201-203),
    with the parameters taken from ONE flat vector in `ravel_pytree` order (utils.py:172-193; per conv layer
    `[weight.ravel(), bias]`, models.py:131-133);
  * the weight drift network `make_w_net` (models.py:57-80): `Linear` layers (time is ignored, diffeq_layers.py:169-175)
    with tanh between them (models.py:136-141);
  * `SDENet.f` / `SDENet.g` / `SDENet.forward` (models.py:159-201) with the INTENDED semantics of SURVEY quirk Q7: a flat
    1-D state `[y, w, logqp]` (the class as shipped raises in `split` / `cat`): `fw = nn - w`, `fl = sum(nn^2)/sigma^2`,
    `g = [0, sigma, 0]`, `logqp = 0.5 * state[-1]`;
  * torchsde's fixed-step Stratonovich solvers for diagonal noise (third-party, un-pinned git dependency, requirements.txt:5;
    absent here - restated from the published algorithm, torchsde/_core/methods/{midpoint,euler_heun,milstein}.py):
        midpoint    y' = y + dt/2 f(t,y) + g(t,y) I/2 ;  y1 = y + dt f(t+dt/2, y') + g(t+dt/2, y') I
        euler_heun  y' = y + g(t,y) I ;                  y1 = y + dt f(t,y) + (g(t,y) + g(t1,y')) I / 2
    with I = W(t1) - W(t0) supplied by the caller (`BrownianInterval` is replaced by explicit increments).  With the constant
    diagonal g of SDENet every noise term is sigma * I.

PARITY STATUS: the conv stack and the w-net are PINNED against the reference's own `make_y_net` / `make_w_net`, which run in
this container (tests/golden/make_golden_torch_sdenet.py -> tests/golden/torch_sdenet.npz); the solver loop is unpinned
(torchsde is not installable; `SDENet.forward` is not runnable as shipped, Q7).
"""
import math

import torch
import torch.nn.functional as F


def act(name, x):
    if name == "softplus":
        return F.softplus(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "elu":
        return F.elu(x)
    if name == "relu":
        return torch.relu(x)
    raise ValueError(name)


# ---------------------------------------------------------------------------------------------------------------
# y_net layout: list of ("conv", dict) / ("act", None) in evaluation order, offsets into the flat vector
# ---------------------------------------------------------------------------------------------------------------
def y_net_layout(input_size, blocks=(2, 2, 2), hidden_width=128, aug_dim=0):
    """models.py:25-54 + diffeq_layers.py:280-311 (mode 0).  Returns (ops, P, output_size)."""
    c, h, w = input_size[0] + aug_dim, input_size[1], input_size[2]
    ops, off = [], 0

    def conv(cin, cout, stride=1, padding=0, transpose=False, output_padding=0):
        nonlocal off
        # nn.Conv2d weight (cout, cin+1, 3, 3); nn.ConvTranspose2d weight (cin+1, cout, 3, 3)
        shape = (cin + 1, cout, 3, 3) if transpose else (cout, cin + 1, 3, 3)
        nw = (cin + 1) * cout * 9
        d = dict(cin=cin, cout=cout, stride=stride, padding=padding, transpose=transpose, output_padding=output_padding,
                 w_off=off, b_off=off + nw, shape=shape)
        off += nw + cout
        ops.append(("conv", d))

    for i, nb in enumerate(blocks, 1):
        for j in range(1, nb + 1):
            conv(c, hidden_width, padding=1)
            ops.append(("act", None))
            conv(hidden_width, hidden_width, stride=2, padding=1)
            ops.append(("act", None))
            conv(hidden_width, hidden_width, stride=2, transpose=True, output_padding=1)
            ops.append(("act", None))
            conv(hidden_width, c)
            if i < len(blocks) or j < nb:   # last_activation
                ops.append(("act", None))
        if i < len(blocks):
            conv(c, 4 * c, stride=2, padding=1)   # ConvDownsample
            c, h, w = 4 * c, h // 2, w // 2
    return ops, off, (c, h, w)


def y_net_apply(ops, flat_w, t, y, activation="softplus"):
    """DiffEqSequential.forward with explicit params (diffeq_layers.py:49-56).  y: (B, C, H, W)."""
    for kind, d in ops:
        if kind == "act":
            y = act(activation, y)
            continue
        W = flat_w[d["w_off"]:d["b_off"]].reshape(d["shape"])
        b = flat_w[d["b_off"]:d["b_off"] + d["cout"]]
        tt = torch.as_tensor(t, dtype=y.dtype).reshape(1, 1, 1, 1).expand(y.shape[0], 1, y.shape[2], y.shape[3])
        ty = torch.cat((tt, y), dim=1)      # utils.channel_cat: time channel FIRST
        if d["transpose"]:
            y = F.conv_transpose2d(ty, W, b, d["stride"], d["padding"], d["output_padding"])
        else:
            y = F.conv2d(ty, W, b, d["stride"], d["padding"])
    return y


def init_y_params(ops, P, gen, dtype=torch.float32):
    """nn.Conv2d / nn.ConvTranspose2d default init (kaiming_uniform(a=sqrt 5) => U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for
    weight and bias; fan_in = shape[1] * 9 as torch computes it)."""
    w = torch.zeros(P, dtype=dtype)
    for kind, d in ops:
        if kind != "conv":
            continue
        bound = 1.0 / math.sqrt(d["shape"][1] * 9)
        n = d["b_off"] - d["w_off"]
        w[d["w_off"]:d["b_off"]] = (torch.rand(n, generator=gen, dtype=dtype) * 2 - 1) * bound
        w[d["b_off"]:d["b_off"] + d["cout"]] = (torch.rand(d["cout"], generator=gen, dtype=dtype) * 2 - 1) * bound
    return w


# ---------------------------------------------------------------------------------------------------------------
# w_net
# ---------------------------------------------------------------------------------------------------------------
def init_w_net(P, hidden_sizes, gen, dtype=torch.float32, last_std=0.0):
    """models.py:57-80: nn.Linear default init, last layer zeros (:67-69); `last_std` > 0 perturbs it."""
    sizes = (P,) + tuple(hidden_sizes) + (P,)
    layers = []
    for i, (din, dout) in enumerate(zip(sizes[:-1], sizes[1:]), 1):
        bound = 1.0 / math.sqrt(din)
        W = (torch.rand(dout, din, generator=gen, dtype=dtype) * 2 - 1) * bound
        b = (torch.rand(dout, generator=gen, dtype=dtype) * 2 - 1) * bound
        if i + 1 == len(sizes):
            W = torch.randn(dout, din, generator=gen, dtype=dtype) * last_std
            b = torch.randn(dout, generator=gen, dtype=dtype) * last_std
        layers.append((W, b))
    return layers


def w_net_apply(layers, w, activation="tanh"):
    """models.py:62-71: Linear -> act -> ... -> Linear (time ignored)."""
    h = w
    for i, (W, b) in enumerate(layers):
        h = F.linear(h, W, b)
        if i + 1 < len(layers):
            h = act(activation, h)
    return h


# ---------------------------------------------------------------------------------------------------------------
# SDENet drift / diffusion on the flat state, fixed-step solvers
# ---------------------------------------------------------------------------------------------------------------
class SDENetOracle:
    def __init__(self, input_size=(3, 32, 32), blocks=(2, 2, 2), weight_network_sizes=(1, 64, 1), activation="softplus",
                 sigma=0.1, hidden_width=128, aug_dim=0):
        self.input_size = tuple(input_size)
        self.aug_dim = aug_dim
        self.aug_input_size = (aug_dim + input_size[0],) + tuple(input_size[1:])
        self.ops, self.P, self.output_size = y_net_layout(input_size, blocks, hidden_width, aug_dim)
        self.activation, self.sigma = activation, sigma
        self.hidden_sizes = tuple(weight_network_sizes)

    def f_parts(self, t, y, w, w_layers):
        """models.py:159-169 on the split state.  y: (B, C, H, W) -> fy reshaped back to y's shape (flat NCHW order)."""
        fy = y_net_apply(self.ops, w, t, y, self.activation).reshape(y.shape[0], -1)
        assert fy.shape[1] == y[0].numel(), "Check nblocks for dataset divisibility."   # models.py:167
        nn_ = w_net_apply(w_layers, w)
        return fy.reshape(y.shape), nn_ - w, (nn_ ** 2).sum() / self.sigma ** 2, nn_

    def solve(self, y, w0, w_layers, increments, dt=0.05, method="midpoint", t0=0.0, t1=1.0, return_path=False):
        """sdeint(self, aug_y, ts=[0,1], bm, method, dt) (models.py:184-197).  increments: (nsteps, P) = W(t_{k+1}) - W(t_k)."""
        n = int(round((t1 - t0) / dt))
        assert increments.shape[0] == n
        w, lq = w0, torch.zeros((), dtype=y.dtype)
        path = [w0]
        for k in range(n):
            tk = t0 + k * dt
            I = increments[k]
            fy, fw, fl, _ = self.f_parts(tk, y, w, w_layers)
            if method == "midpoint":
                y_p = y + 0.5 * dt * fy
                w_p = w + 0.5 * dt * fw + 0.5 * self.sigma * I
                fy2, fw2, fl2, _ = self.f_parts(tk + 0.5 * dt, y_p, w_p, w_layers)
                y, w, lq = y + dt * fy2, w + dt * fw2 + self.sigma * I, lq + dt * fl2
                path += [w_p, w]
            elif method == "heun":
                # torchsde heun (Stratonovich trapezoidal rule): y' = y + dt f(t, y) + g I; y1 = y + dt/2 (f(t, y) + f(t + dt, y')) + g I
                y_p, w_p = y + dt * fy, w + dt * fw + self.sigma * I
                fy2, fw2, fl2, _ = self.f_parts(tk + dt, y_p, w_p, w_layers)
                y, w, lq = y + 0.5 * dt * (fy + fy2), w + 0.5 * dt * (fw + fw2) + self.sigma * I, lq + 0.5 * dt * (fl + fl2)
                path += [w_p, w]
            elif method in ("euler_heun", "milstein"):
                # constant diagonal g: g(t1, y') = g(t, y) and g dg = 0, so both reduce to the Euler update
                y, w, lq = y + dt * fy, w + dt * fw + self.sigma * I, lq + dt * fl
                path += [w]
            else:
                raise ValueError(f"Unknown method: {method}")
        if return_path:
            return y, w, lq, torch.stack(path)
        return y, w, lq

    def adjoint(self, yT, wT, cot_y, cot_w, cot_lq, w_layers, increments, dt=0.05, method="midpoint", t0=0.0, t1=1.0,
                reverse_grid=None):
        """The reverse pass of `torchsde.sdeint_adjoint` (call site models.py:184-197 with `adjoint=True`; torchsde is an un-pinned git
        dependency that is absent here, so this restates the published scheme: Li et al., "Scalable gradients for stochastic
        differential equations", AISTATS 2020, Alg. 2).  The augmented state [y, w | a_y, a_w, a_lq | a_params] is integrated from t1 back
        to t0 with the SAME fixed-step scheme and step as the forward solve on the same Brownian path run backwards
        (W~(s) = W(-s): the increment of reverse step k is -increments[k]):
            d[y, w]   = -f(tau, y, w) ds + g dW~          (Stratonovich, g = [0, sigma, 0] constant: no correction term)
            d a       = +(df/d[y, w])^T a ds              (g does not depend on the state or the parameters)
            d a_par   = +(df/dparams)^T a ds
        a_lq stays cot_lq (no drift reads logqp).  Returns (y0_rec, w0_rec, dL/dy0, dL/dw0, [dL/d(W, b) per w_net layer])."""
        n = int(round((t1 - t0) / dt))
        assert reverse_grid is not None or increments.shape[0] == n
        params = [t for lay in w_layers for t in lay]
        y, w, ay, aw = yT.detach(), wT.detach(), cot_y.detach(), cot_w.detach()
        ap = [torch.zeros_like(p) for p in params]

        def F(tau, y, w, ay, aw):
            with torch.enable_grad():
                y_, w_ = y.detach().requires_grad_(True), w.detach().requires_grad_(True)
                ps = [p.detach().requires_grad_(True) for p in params]
                lay = [tuple(ps[i:i + 2]) for i in range(0, len(ps), 2)]
                fy, fw, fl, _ = self.f_parts(tau, y_, w_, lay)
                obj = (fy * ay).sum() + (fw * aw).sum() + fl * cot_lq
                gs = torch.autograd.grad(obj, [y_, w_] + ps, allow_unused=True)
            gs = [torch.zeros_like(q) if g is None else g for g, q in zip(gs, [y_, w_] + ps)]
            return fy.detach(), fw.detach(), gs[0], gs[1], gs[2:]

        def rev_step(tau, dt, dWr, y, w, ay, aw, ap):
            fy, fw, vy, vw, vp = F(tau, y, w, ay, aw)
            if method == "midpoint":
                y_p, w_p = y - 0.5 * dt * fy, w - 0.5 * dt * fw + 0.5 * self.sigma * dWr
                ay_p, aw_p = ay + 0.5 * dt * vy, aw + 0.5 * dt * vw
                fy2, fw2, vy2, vw2, vp2 = F(tau - 0.5 * dt, y_p, w_p, ay_p, aw_p)
                y, w = y - dt * fy2, w - dt * fw2 + self.sigma * dWr
                ay, aw = ay + dt * vy2, aw + dt * vw2
                ap = [a + dt * v for a, v in zip(ap, vp2)]
            elif method == "heun":
                y_p, w_p = y - dt * fy, w - dt * fw + self.sigma * dWr
                ay_p, aw_p = ay + dt * vy, aw + dt * vw
                fy2, fw2, vy2, vw2, vp2 = F(tau - dt, y_p, w_p, ay_p, aw_p)
                y, w = y - 0.5 * dt * (fy + fy2), w - 0.5 * dt * (fw + fw2) + self.sigma * dWr
                ay, aw = ay + 0.5 * dt * (vy + vy2), aw + 0.5 * dt * (vw + vw2)
                ap = [a + 0.5 * dt * (v + v2) for a, v, v2 in zip(ap, vp, vp2)]
            elif method in ("euler_heun", "milstein"):
                y, w = y - dt * fy, w - dt * fw + self.sigma * dWr
                ay, aw = ay + dt * vy, aw + dt * vw
                ap = [a + dt * v for a, v in zip(ap, vp)]
            else:
                raise ValueError(f"Unknown method: {method}")
            return y, w, ay, aw, ap

        if reverse_grid is not None:   # (s_k, h_k) of an adaptive reverse solve, s = -t; bm(t) -> W(t)
            import numpy as np
            bm, ts_r, hs_r = reverse_grid
            for sa, h in zip(ts_r, hs_r):
                tau = float(np.float32(-sa))
                dWr = bm(float(np.float32(tau) - np.float32(h))) - bm(tau)
                y, w, ay, aw, ap = rev_step(tau, float(h), dWr, y, w, ay, aw, ap)
            return y, w, ay, aw, ap
        for k in range(n - 1, -1, -1):
            y, w, ay, aw, ap = rev_step(t0 + (k + 1) * dt, dt, -increments[k], y, w, ay, aw, ap)
        return y, w, ay, aw, ap

    # ---- adaptive stepping: torchsde's step-doubling controller (un-pinned git dependency, absent here; restated from the published
    # scheme: BaseSDESolver.integrate with adaptive=True + adaptive_stepping.{update_step_size, compute_error}) ----------------------
    @staticmethod
    def update_step_size(error_estimate, prev_step_size, safety=0.9, facmin=0.2, facmax=1.4, prev_error_ratio=None):
        if error_estimate > 1:
            pfactor, ifactor = 0.0, 1.0 / 1.5
        else:
            pfactor, ifactor = 0.13, 1.0 / 4.5
        error_ratio = safety / error_estimate
        if prev_error_ratio is None:
            prev_error_ratio = error_ratio
        factor = error_ratio ** ifactor * (error_ratio / prev_error_ratio) ** pfactor
        if error_estimate <= 1:
            prev_error_ratio = error_ratio
            facmin = 1.0
        factor = min(facmax, max(facmin, factor))
        return prev_step_size * factor, prev_error_ratio

    @staticmethod
    def compute_error(full, halves, rtol, atol, eps=1e-7):
        num, cnt = 0.0, 0
        for a, b in zip(full, halves):
            tol = rtol * torch.maximum(a.abs(), b.abs()) + atol
            num = num + (((a - b) / (tol + eps)) ** 2).sum()
            cnt += a.numel()
        return max(float(torch.sqrt(num / cnt)), eps)

    def solve_adaptive(self, y, w0, w_layers, bm, dt, rtol, atol, method="midpoint", t0=0.0, t1=1.0, dt_min=1e-5):
        """ts = [t0, t1]; `bm(t)` -> W(t) over the P weight rows.  Returns (accepted t list, accepted h list, trials' error estimates,
        (y, w, logqp) at t1).  Time arithmetic in float32 as on the product side."""
        import numpy as np
        f32 = np.float32

        def one(ta, h, state):
            yy, ww, lq = state
            I = (bm(float(f32(ta + h))) - bm(float(ta)))[None]
            y1, w1, l1 = self.solve(yy, ww, w_layers, I, dt=float(h), method=method, t0=float(ta), t1=float(ta) + float(h))
            return y1, w1, lq + l1

        state = (y, w0, torch.zeros((), dtype=y.dtype))
        curr, step, prev = f32(t0), f32(dt), None
        ts, hs, errs = [], [], []
        while curr < f32(t1):
            nxt_t = min(f32(curr + step), f32(t1))
            mid_t = f32(f32(0.5) * f32(curr + nxt_t))
            full = one(curr, f32(nxt_t - curr), state)
            mid = one(curr, f32(mid_t - curr), state)
            nxt = one(mid_t, f32(nxt_t - mid_t), mid)
            err = self.compute_error(full, nxt, rtol, atol)
            errs.append(err)
            new_step, prev = self.update_step_size(err, float(step), prev_error_ratio=prev)
            step = f32(new_step)
            if step < dt_min:
                step, prev = f32(dt_min), None
            if err <= 1 or step <= dt_min:
                ts += [curr, mid_t]
                hs += [f32(mid_t - curr), f32(nxt_t - mid_t)]
                curr, state = nxt_t, nxt
        return ts, hs, errs, state

    def solve_on_grid(self, y, w0, w_layers, bm, ts, hs, method="midpoint"):
        """The fixed-step schemes on an arbitrary step grid (the replay of an adaptive solve's accepted steps)."""
        import numpy as np
        w, lq = w0, torch.zeros((), dtype=y.dtype)
        for ta, h in zip(ts, hs):
            I = (bm(float(np.float32(ta + h))) - bm(float(ta)))[None]
            y, w, l1 = self.solve(y, w, w_layers, I, dt=float(h), method=method, t0=float(ta), t1=float(ta) + float(h))
            lq = lq + l1
        return y, w, lq

    def forward(self, x, w0, w_layers, proj, increments, dt=0.05, method="midpoint"):
        """SDENet.forward (models.py:180-201): zero channels, solve, projection head, logqp = state[-1] / 2."""
        if self.aug_dim > 0:
            x = torch.cat((x, x.new_zeros(x.shape[0], self.aug_dim, *x.shape[2:])), dim=1)
        y1, _, lq = self.solve(x, w0, w_layers, increments, dt, method)
        (W1, b1), (W2, b2) = proj
        logits = F.linear(torch.relu(F.linear(y1.reshape(y1.shape[0], -1), W1, b1)), W2, b2)
        return logits, 0.5 * lq

"""CPU restatement (torch, fp32 or fp64) of the reference's JAX front-end hot path.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  **parity unpinned** by the
reference's own tests (no golden vectors exist for the fixed-grid path).

Every function cites the reference lines (relative to /root/reference) whose
arithmetic it follows.  The Brownian tree (jaxsde/jaxsde/brownian.py:26-95) is
NOT restated: increments dW_k are an input, which is the "supplied increments"
mode of the product; only the `rep` noise tying (brownian.py:69-74) is kept.

State layout follows brax/_impl/brax.py:110:  y = [x.reshape(-1), w (P), kl (P)].
Images are NHWC, conv kernels HWIO (brax/_impl/diffeq_layers.py:100), time
channel appended LAST (diffeq_layers.py:145-146).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# activations: brax/_impl/arch.py:14-32 (ACTFNS)
# --------------------------------------------------------------------------------------


def act(name, x, beta=None):
    if name == "softplus":
        return F.softplus(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "elu":
        return F.elu(x)
    if name == "rbf":
        return torch.exp(-x ** 2)
    if name == "swish":  # arch.py:17  x * sigmoid(x * softplus(beta))
        return x * torch.sigmoid(x * F.softplus(beta))
    raise ValueError(name)


# --------------------------------------------------------------------------------------
# time grid: jaxsde/jaxsde/sdeint.py:103-110 (scan_fun_one_step) -- quirk Q1
# --------------------------------------------------------------------------------------


def time_grid(ts, mode="reference"):
    """Per-iteration (t_k, dt_k) for k = 0..len(ts)-2, computed in the dtype of `ts`.

    mode="reference": the carry arithmetic of sdeint.py:105-106
        next_t = target_t - curr_t ; t_delta = next_t - curr_t ; step evaluated at curr_t.
    mode="uniform":   the intended grid  t_k = ts[k], dt_k = ts[k+1] - ts[k].
    """
    ts = torch.as_tensor(ts)
    tk, dtk = [], []
    if mode == "reference":
        curr = ts[0].clone()
        for k in range(1, len(ts)):
            nxt = ts[k] - curr
            tk.append(curr)
            dtk.append(nxt - curr)
            curr = nxt
    elif mode == "uniform":
        for k in range(1, len(ts)):
            tk.append(ts[k - 1])
            dtk.append(ts[k] - ts[k - 1])
    else:
        raise ValueError(mode)
    return torch.stack(tk), torch.stack(dtk)


# --------------------------------------------------------------------------------------
# generic fixed-grid Ito solver: jaxsde/jaxsde/sdeint.py:16-18, 36-64, 81-138
# --------------------------------------------------------------------------------------


def _gdg_prod(g, y, t, args, v):
    """sde_utils.py:55-61 make_gdg_prod: jvp(g(.,t,args), y, v*g(y)) for elementwise g.

    Differentiable like the reference's jvp (the scan is reverse-differentiated through it): for an elementwise g
    the Jacobian is diagonal, J @ tangent = d(sum g)/dy * tangent, built with create_graph so that gradients flow
    through g(y), through the diagonal and through g's parameters."""
    y_ = y if y.requires_grad else y.detach().clone().requires_grad_(True)
    gy = g(y_, t, args)
    if not gy.requires_grad or gy.grad_fn is None:  # state-independent diffusion => zero Jacobian
        return torch.zeros_like(y)
    (diag,) = torch.autograd.grad(gy.sum(), y_, create_graph=True, allow_unused=True)
    if diag is None:
        return torch.zeros_like(y)
    return diag * (v * gy)


def sdeint_ito_fixed_grid(f, g, y0, ts, increments, args=(), dt=1e-6, method="milstein", rep=0,
                          time_grid_mode="reference", differentiable_milstein=False):
    """sdeint.py:16-18 with the Brownian motion replaced by supplied increments.

    increments: (len(ts)-1, D) tensor, increments[k] = bm(next_t) - bm(curr_t) of iteration k
    (sdeint.py:108).  With rep>0 the caller must already have tied the last `rep` entries to
    the `rep` before them (brownian.py:69-74); `tie_rep` does that.
    Returns ys of shape (len(ts), D) including y0 (sdeint.py:138).
    `dt` is ignored on the fixed grid, as in the reference.
    """
    if method not in ("milstein", "euler_maruyama", "euler_heun"):
        raise ValueError("Unknown method: {}".format(method))  # sdeint.py:62
    tk, dtk = time_grid(ts, time_grid_mode)
    y = y0
    ys = [y0]
    for k in range(len(tk)):
        t, h, dw = tk[k], dtk[k], increments[k]
        if method == "euler_maruyama":  # sdeint.py:48-50
            y = y + h * f(y, t, args) + g(y, t, args) * dw
        elif method == "milstein":  # sdeint.py:36-39
            corr = 0.5 * _gdg_prod(g, y, t, args, dw ** 2 - h)
            y = y + h * f(y, t, args) + g(y, t, args) * dw + corr
        else:
            # euler_heun is only a NotImplementedError stub in the reference
            # (sdeint.py:73-74); the north star asks for it, so it is restated from the
            # published scheme (Stratonovich, diagonal noise):
            #   y' = y + g(y,t) dW ;  y1 = y + h f(y,t) + 1/2 (g(y,t) + g(y',t+h)) dW
            gy = g(y, t, args)
            yp = y + gy * dw
            y = y + h * f(y, t, args) + 0.5 * (gy + g(yp, t + h, args)) * dw
        ys.append(y)
    return torch.stack(ys)


def tie_rep(z, rep):
    """brownian.py:69-74 sample_with_rep: last `rep` entries := the `rep` entries before them."""
    if rep == 0:
        return z
    z = z.clone()
    n = z.shape[-1]
    z[..., n - rep:] = z[..., n - 2 * rep:n - rep]
    return z


# --------------------------------------------------------------------------------------
# XLA conv restatement: lax.conv_general_dilated / lax.conv_transpose [3P, jax 0.1.77]
# call sites brax/_impl/diffeq_layers.py:127-135,147
# --------------------------------------------------------------------------------------


def xla_same_pads(in_size, k, stride):
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    lo = total // 2
    return lo, total - lo


def conv_transpose_same_pads(k, s):
    """jax.lax._conv_transpose_padding(k, s, 'SAME') [3P]."""
    pad_len = k + s - 2
    pad_a = k - 1 if s > k - 1 else int(math.ceil(pad_len / 2))
    return pad_a, pad_len - pad_a


def conv_general_dilated_nhwc(x, w, stride, pads, lhs_dilation=1):
    """x (N,H,W,Ci), w (kh,kw,Ci,Co) HWIO, cross-correlation (no kernel flip)."""
    xn = x.permute(0, 3, 1, 2)
    if lhs_dilation > 1:
        n, c, h, wd = xn.shape
        xd = xn.new_zeros(n, c, (h - 1) * lhs_dilation + 1, (wd - 1) * lhs_dilation + 1)
        xd[:, :, ::lhs_dilation, ::lhs_dilation] = xn
        xn = xd
    (pt, pb), (pl, pr) = pads
    xn = F.pad(xn, (pl, pr, pt, pb))
    out = F.conv2d(xn.contiguous(), w.permute(3, 2, 0, 1).contiguous(), stride=stride)   # (torch's CPU conv backward needs contiguous weights)
    return out.permute(0, 2, 3, 1)


def concat_conv2d(x, t, w, b, stride=1, transpose=False):
    """brax/_impl/diffeq_layers.py:124-150 ConcatConv2D, padding="SAME", 3x3 (or kxk)."""
    tt = torch.ones_like(x[..., :1]) * t  # :145
    xtt = torch.cat([x, tt], dim=-1)  # :146
    k = w.shape[0]
    if not transpose:
        ph = xla_same_pads(x.shape[1], k, stride)
        pw = xla_same_pads(x.shape[2], k, stride)
        out = conv_general_dilated_nhwc(xtt, w, stride, (ph, pw))
    else:
        p = conv_transpose_same_pads(k, stride)
        out = conv_general_dilated_nhwc(xtt, w, 1, (p, p), lhs_dilation=stride)
    return out + b.reshape(1, 1, 1, -1)


# --------------------------------------------------------------------------------------
# f_x: brax/_impl/arch.py:171-210 build_fx; flat weight layout brax.py:41 (ravel_pytree)
# --------------------------------------------------------------------------------------


def fx_layout(block_type, in_ch, fx_dim, actfn="softplus"):
    """List of (kind, shape, offset) in ravel order, and total P.

    Per layer [W (k,k,Cin+1,Cout) C-order, b (Cout)], swish beta (fx_dim,) after each
    activation when actfn == "swish" (arch.py:172-174)."""
    if block_type == 0:
        convs = [(in_ch, fx_dim, 3, 1, False), (fx_dim, fx_dim, 3, 2, False),
                 (fx_dim, fx_dim, 3, 2, True), (fx_dim, in_ch, 3, 1, False)]
    elif block_type == 1:
        # arch.py:193 passes (1, 1) into ConcatConv2D's unused W_init slot (diffeq_layers.py:124): kernel stays 3
        convs = [(in_ch, fx_dim, 3, 1, False), (fx_dim, fx_dim, 3, 1, False), (fx_dim, in_ch, 3, 1, False)]
    elif block_type == 2:
        convs = [(in_ch, fx_dim, 3, 1, False), (fx_dim, in_ch, 3, 1, False)]
    else:
        raise ValueError(f"Invalid block_type {block_type}")
    entries, off = [], 0
    for i, (ci, co, k, s, tr) in enumerate(convs):
        wshape = (k, k, ci + 1, co)
        n = k * k * (ci + 1) * co
        entries.append(dict(kind="conv", cin=ci, cout=co, k=k, stride=s, transpose=tr,
                            w_off=off, w_shape=wshape, b_off=off + n))
        off += n + co
        if i < len(convs) - 1 and actfn == "swish":
            entries.append(dict(kind="beta", off=off, n=fx_dim))
            off += fx_dim
    return entries, off


def fx_apply(flat_w, x, t, layout, actfn="softplus"):
    """arch.py:176-188 evaluated with weights unflattened from `flat_w` (brax.py:52)."""
    convs = [e for e in layout if e["kind"] == "conv"]
    betas = [e for e in layout if e["kind"] == "beta"]
    h = x
    for i, e in enumerate(convs):
        n = e["k"] * e["k"] * (e["cin"] + 1) * e["cout"]
        w = flat_w[e["w_off"]:e["w_off"] + n].reshape(e["w_shape"])
        b = flat_w[e["b_off"]:e["b_off"] + e["cout"]]
        h = concat_conv2d(h, t, w, b, stride=e["stride"], transpose=e["transpose"])
        if i < len(convs) - 1:
            beta = flat_w[betas[i]["off"]:betas[i]["off"] + betas[i]["n"]] if betas else None
            h = act(actfn, h, beta)
    return h


# --------------------------------------------------------------------------------------
# f_w: brax/_impl/arch.py:34-74 MLP(xt=True); ConcatSquashLinear diffeq_layers.py:73-96
# --------------------------------------------------------------------------------------


def concat_squash_linear(p, x, t):
    W, b, w_t, w_tb, b_t = p[:5]
    out = x @ W + b  # :89
    out = out * torch.sigmoid(w_t * t + w_tb)  # :91
    return out + b_t * t  # :93


def fw_apply(fw_layers, w, t, actfn="softplus"):
    """fw_layers: list of (W,b,w_t,w_tb,b_t), one per ConcatSquashLinear; act between.  With the trainable swish
    (arch.py:14-29,41-42) every layer but the last carries the beta vector of the activation that follows it as a sixth entry.

    With ou_dw=True the reference wraps the MLP in arch.Additive but, because of the
    (out,t)/(t,out) unpack mismatch (quirk Q3; arch.py:100-103 vs diffeq_layers.py:95),
    element [0] of the result is exactly the MLP output: the OU term is NOT added."""
    h = w
    for i, p in enumerate(fw_layers):
        h = concat_squash_linear(p, h, t)
        if i < len(fw_layers) - 1:
            h = act(actfn, h, p[5] if len(p) > 5 else None)
    return h


def init_fw(P, fw_dims, gen, dtype=torch.float32, last_std=0.0, actfn="softplus", beta_jitter=0.0):
    """arch.py:34-74 + diffeq_layers.py:76-82 shapes; he-normal W, normal(1e-2) others;
    last layer zeros (arch.py:51,59) unless last_std>0 (SURVEY §8.4 synthetic init).  actfn="swish": beta = 0.5 (arch.py:14)
    as a sixth entry of every layer but the last (+ `beta_jitter` * N(0,1) so that tests see distinct values)."""
    dims = [P] + list(fw_dims) + [P]
    layers = []
    for i in range(len(dims) - 1):
        din, dout = dims[i], dims[i + 1]
        last = i == len(dims) - 2
        if last:
            std_w = std_o = last_std
        else:
            std_w, std_o = math.sqrt(2.0 / din), 1e-2
        mk = lambda *s, sd: (torch.randn(*s, generator=gen, dtype=torch.float64) * sd).to(dtype)
        lay = (mk(din, dout, sd=std_w), mk(dout, sd=std_o), mk(dout, sd=std_o), mk(dout, sd=std_o), mk(dout, sd=std_o))
        if actfn == "swish" and not last:
            lay = lay + ((0.5 + mk(dout, sd=beta_jitter).double()).to(dtype),)
        layers.append(lay)
    return layers


def init_fx(layout, P, gen, dtype=torch.float32):
    """stax.GeneralConv defaults [3P]: W glorot-normal, b normal(1e-6); swish beta 0.5."""
    w = torch.zeros(P, dtype=dtype)
    for e in layout:
        if e["kind"] == "conv":
            k, ci1, co = e["k"], e["cin"] + 1, e["cout"]
            n = k * k * ci1 * co
            std = math.sqrt(2.0 / (k * k * ci1 + k * k * co))
            w[e["w_off"]:e["w_off"] + n] = (torch.randn(n, generator=gen, dtype=torch.float64) * std).to(dtype)
            w[e["b_off"]:e["b_off"] + co] = (torch.randn(co, generator=gen, dtype=torch.float64) * 1e-6).to(dtype)
        else:
            w[e["off"]:e["off"] + e["n"]] = 0.5
    return w


# --------------------------------------------------------------------------------------
# SDEBNN block: brax/_impl/brax.py:15-135
# --------------------------------------------------------------------------------------


class SDEBNNBlock:
    """One `SDEBNN(...)` layer for a fixed input shape (brax.py:33-133, make_layer)."""

    def __init__(self, x_shape, fx_block_type=0, fx_dim=64, fx_actfn="softplus", fw_actfn="softplus",
                 diff_coef=1e-4, stl=False, nsteps=20, w_drift=True, stl_mode="reference",
                 time_grid_mode="reference", method="euler_maruyama"):
        self.x_shape = tuple(x_shape)  # (B,H,W,C)
        self.layout, self.P = fx_layout(fx_block_type, x_shape[-1], fx_dim, fx_actfn)
        self.fx_actfn, self.fw_actfn = fx_actfn, fw_actfn
        self.diff_coef, self.stl, self.nsteps, self.w_drift = diff_coef, stl, nsteps, w_drift
        self.stl_mode, self.time_grid_mode, self.method = stl_mode, time_grid_mode, method
        self.x_dim = int(math.prod(self.x_shape))
        self.w_stale = None  # brax.py:38-41: flat_w captured by g_aug (quirk Q2)

    def f_aug(self, y, t, fw_layers):  # brax.py:49-62
        x = y[:self.x_dim].reshape(self.x_shape)
        flat_w = y[self.x_dim:self.x_dim + self.P]
        dx = fx_apply(flat_w, x, t, self.layout, self.fx_actfn)
        dw = fw_apply(fw_layers, flat_w, t, self.fw_actfn) if self.w_drift else torch.zeros_like(flat_w)
        u = (dw - (-flat_w)) / self.diff_coef if self.diff_coef != 0 else torch.zeros_like(flat_w)
        dkl = u ** 2  # :61, no 1/2 (quirk Q4)
        return torch.cat([dx.reshape(-1), dw, dkl])

    def g_aug(self, y, t, fw_layers):  # brax.py:64-78
        dx = y.new_zeros(self.x_dim)
        diff_w = y.new_ones(self.P) * self.diff_coef
        if self.stl:
            if self.stl_mode == "reference":
                flat_w = self.w_stale.to(y.dtype)  # closure variable, not the state (Q2)
            else:
                flat_w = y[self.x_dim:self.x_dim + self.P].detach()
            sg = [tuple(q.detach() for q in p) for p in fw_layers]  # stop_gradient, :69
            drift_w = fw_apply(sg, flat_w, t, self.fw_actfn) if self.w_drift else torch.zeros_like(flat_w)
            u = (drift_w + flat_w) / self.diff_coef if self.diff_coef != 0 else torch.zeros_like(flat_w)
            dkl = u
        else:
            dkl = y.new_zeros(self.P)
        return torch.cat([dx, diff_w, dkl])

    def apply(self, params, x, dW):
        """brax.py:97-129.  params = (w0 (P,), fw_layers).  dW: (nsteps-1, P) increments of the
        w rows; the x rows have g=0 and the kl rows reuse the w noise when stl (rep tying,
        brax.py:111 + brownian.py:69-74), else their g is 0 too."""
        w0, fw_layers = params
        dt_ = x.dtype
        ts = torch.linspace(0.0, 1.0, self.nsteps, dtype=torch.float32).to(dt_)  # brax.py:31
        y0 = torch.cat([x.reshape(-1), w0, torch.zeros_like(w0)])  # :110
        inc = torch.cat([dW.new_zeros(dW.shape[0], self.x_dim), dW, dW], dim=1)
        ys = sdeint_ito_fixed_grid(self.f_aug, self.g_aug, y0, ts, inc, fw_layers,
                                   method=self.method, rep=self.P if self.stl else 0,
                                   time_grid_mode=self.time_grid_mode)
        y = ys[-1]
        xo = y[:self.x_dim].reshape(self.x_shape)
        kl = y[self.x_dim + self.P:].sum()  # :119
        return xo, kl, ys[:, self.x_dim:self.x_dim + self.P]


# --------------------------------------------------------------------------------------
# network glue: arch.py:108-168, brax.py:138-219, examples/jax/sdebnn_classification.py
# --------------------------------------------------------------------------------------


def augment(x, pad):  # arch.py:108-128
    return torch.cat([x, x.new_zeros(x.shape[:-1] + (pad,))], dim=-1)


def squeeze_downsample(x, r=2):  # arch.py:131-153 (NHWC in/out via NCHW)
    xn = x.permute(0, 3, 1, 2)
    b, c, h, w = xn.shape
    v = xn.reshape(b, c, h // r, r, w // r, r).permute(0, 1, 3, 5, 2, 4)
    return v.reshape(b, c * r * r, h // r, w // r).permute(0, 2, 3, 1)


def normal_logprob(z, mean, log_std):  # brax.py:171-178
    c = math.log(2 * math.pi)
    tmp = (z - mean) * torch.exp(-log_std)
    return -0.5 * (tmp * tmp + 2 * log_std + c)


def meanfield_dense_head(params, x, eps, prior_std=0.1):
    """MeanField(serial(Flatten, Dense(10), LogSoftmax)): brax.py:138-168, script :211."""
    (Wm, bm), (Wl, bl) = params
    flat_mean = torch.cat([Wm.reshape(-1), bm])
    flat_logstd = torch.cat([Wl.reshape(-1), bl])
    z = eps * torch.exp(flat_logstd) + flat_mean  # :155
    kl = normal_logprob(z, flat_mean, flat_logstd) - normal_logprob(
        z, torch.zeros((), dtype=z.dtype), torch.tensor(math.log(prior_std), dtype=z.dtype))
    W, b = z[:Wm.numel()].reshape(Wm.shape), z[Wm.numel():]
    logits = x.reshape(x.shape[0], -1) @ W + b
    return F.log_softmax(logits, dim=-1), kl.sum()


class SDEBNNClassifier:
    """examples/jax/sdebnn_classification.py:176-213 (--model sdenet), one MC sample."""

    def __init__(self, input_size=(32, 32, 3), batch=4, aug=2, nblocks=(2, 2, 2), fx_dim=64,
                 fw_dims=(2, 64, 2), fx_actfn="softplus", fw_actfn="softplus", diff_coef=1e-4,
                 stl=False, nsteps=20, block_type=0, num_classes=10, **kw):
        H, W, C = input_size
        self.aug, self.fw_dims, self.num_classes = aug, tuple(fw_dims), num_classes
        self.stages = []  # list of ("sde", block) / ("squeeze",)
        shape = (batch, H, W, C + aug)
        for i, nb in enumerate(nblocks):
            for _ in range(nb):
                self.stages.append(("sde", SDEBNNBlock(shape, block_type, fx_dim, fx_actfn, fw_actfn,
                                                       diff_coef, stl, nsteps, **kw)))
            if i < len(nblocks) - 1:
                self.stages.append(("squeeze",))
                shape = (batch, shape[1] // 2, shape[2] // 2, shape[3] * 4)
        self.feat = shape[1] * shape[2] * shape[3]

    def blocks(self):
        return [s[1] for s in self.stages if s[0] == "sde"]

    def init(self, seed=0, dtype=torch.float32, fw_last_std=1e-2):
        gen = torch.Generator().manual_seed(seed)
        params = []
        for blk in self.blocks():
            w0 = init_fx(blk.layout, blk.P, gen, dtype)
            fw = init_fw(blk.P, self.fw_dims, gen, dtype, last_std=fw_last_std, actfn=blk.fw_actfn)
            params.append((w0, fw))
            # brax.py:38-41: the stale vector is a fresh fx.init from PRNGKey(0)
            blk.w_stale = init_fx(blk.layout, blk.P, torch.Generator().manual_seed(1000 + len(params)), dtype)
        std = math.sqrt(2.0 / (self.feat + self.num_classes))
        Wm = (torch.randn(self.feat, self.num_classes, generator=gen, dtype=torch.float64) * std).to(dtype)
        bm = (torch.randn(self.num_classes, generator=gen, dtype=torch.float64) * 1e-2).to(dtype)
        head = ((Wm, bm), (torch.full_like(Wm, -4.0), torch.full_like(bm, -4.0)))  # brax.py:144
        return params, head

    def predict(self, params, head, x, dWs, eps_head):
        """brax.py:202-217 bnn_serial.apply_fun for one MC sample -> (logp, kl_total, w paths)."""
        h = augment(x, self.aug)
        kl_total = 0.0
        bi = 0
        for st in self.stages:
            if st[0] == "sde":
                h, kl, _ = st[1].apply(params[bi], h, dWs[bi])
                kl_total = kl_total + kl
                bi += 1
            else:
                h = squeeze_downsample(h)
        logp, kl_h = meanfield_dense_head(head, h, eps_head)
        return logp, kl_total + kl_h

    def loss(self, params, head, x, onehot, dWs, eps_head, kl_scale):
        """script :27-31,46-51: nll + kl * kl_coef/train_size for one MC sample."""
        logp, kl = self.predict(params, head, x, dWs, eps_head)
        nll = -(logp * onehot).sum(1).mean()
        return nll + kl * kl_scale, nll, kl, logp


def sample_increments(P, nsteps, gen, dtype, time_grid_mode="reference"):
    """N(0, dt_k) increments for one block: the law of bm(next_t)-bm(curr_t) (brownian.py)."""
    ts = torch.linspace(0.0, 1.0, nsteps, dtype=torch.float32)
    _, dtk = time_grid(ts, time_grid_mode)
    z = torch.randn(nsteps - 1, P, generator=gen, dtype=torch.float64)
    return (z * dtk.double().clamp_min(0).sqrt()[:, None]).to(dtype)

"""Structured route of `sdeint_ito_fixed_grid` for SDEBNN dynamics: K1 (weight-SDE rows) + K2/K3
(x rows, forward scan and hand-written reverse pass) through the C ABI of libbsde.so.

What the reference computes in one `lax.scan` over the concatenated state [x, w, kl]
(jaxsde/jaxsde/sdeint.py:103-110,135 with f_aug/g_aug of brax/_impl/brax.py:49-78) is split by data
dependence: the [w, kl] rows never read x, so their whole trajectory is integrated first (one
cooperative launch), then the x rows consume w(t_k) step by step.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .arch import fw_betas, fw_layers  # noqa: F401


def reference_time_grid(ts, mode="reference"):
    """(t_k, dt_k) of every scan iteration in fp32, as jaxsde/jaxsde/sdeint.py:103-110 evaluates
    them (mode="reference": next_t = target_t - curr_t; t_delta = next_t - curr_t -- quirk Q1) or on
    the intended uniform grid (mode="uniform")."""
    ts = np.asarray(ts, dtype=np.float32)
    tk, dtk = [], []
    if mode == "reference":
        curr = ts[0]
        for k in range(1, len(ts)):
            nxt = np.float32(ts[k] - curr)
            tk.append(curr)
            dtk.append(np.float32(nxt - curr))
            curr = nxt
    elif mode == "uniform":
        for k in range(1, len(ts)):
            tk.append(ts[k - 1])
            dtk.append(np.float32(ts[k] - ts[k - 1]))
    else:
        raise ValueError(f"time_grid must be 'reference' or 'uniform', got {mode!r}")
    return np.asarray(tk, dtype=np.float32), np.asarray(dtk, dtype=np.float32)


def _fw_struct(layers, P, hidden_dims, actfn, betas=()):
    """Fill a bsde_fw_params with the device pointers of the (W,b,w_t,w_tb,b_t) tensors (and the swish beta vectors)."""
    fp = _lib.FwParams()
    fp.P = P
    fp.nhid = len(hidden_dims)
    for i, d in enumerate(hidden_dims):
        fp.dims[i] = d
    fp.act = _lib.ACT[actfn]
    assert len(layers) == len(hidden_dims) + 1
    for name, t in zip(("W1", "b1", "wt1", "wtb1", "bt1"), layers[0]):
        setattr(fp, name, _lib.ptr(t).value)
    for i, lay in enumerate(layers[1:-1]):
        for name, t in zip(("Wm", "bm", "wtm", "wtbm", "btm"), lay):
            getattr(fp, name)[i] = _lib.ptr(t).value
    for name, t in zip(("W4", "b4", "wt4", "wtb4", "bt4"), layers[-1]):
        setattr(fp, name, _lib.ptr(t).value)
    if actfn == "swish":
        if len(betas) != len(hidden_dims):
            raise ValueError(f"swish f_w needs one beta vector per hidden layer ({len(hidden_dims)}), got {len(betas)}")
        fp.beta1 = _lib.ptr(betas[0]).value
        for i, t in enumerate(betas[1:]):
            fp.betam[i] = _lib.ptr(t).value
    return fp


def split_fw_flat(cfg, fw_flat):
    """The flat f_w tensor list of the autograd functions: five tensors per layer, then (swish only) one beta per hidden layer."""
    nl = len(cfg.fw["hidden_dims"]) + 1
    layers = [tuple(t.contiguous() for t in fw_flat[5 * i:5 * i + 5]) for i in range(nl)]
    betas = [t.contiguous() for t in fw_flat[5 * nl:]]
    return layers, betas


_GRID_CACHE = {}


class BlockConfig:
    """Static description of one SDEBNN block's fused solve."""

    def __init__(self, fx_spec, fw_spec, S, B, nsteps, diff_coef, stl, time_grid, math="fp32", kl_weight=1.0,
                 ou_coef=0.0, ts=None, remat=False):
        self.remat = remat
        self.fx, self.fw = fx_spec, fw_spec
        self.S, self.B, self.nsteps = S, B, nsteps
        self.nit = nsteps - 1
        self.sigma, self.stl = float(diff_coef), bool(stl)
        self.math = math
        self.kl_weight, self.ou_coef = kl_weight, ou_coef
        ts = np.linspace(0.0, 1.0, nsteps, dtype=np.float32) if ts is None else ts  # brax.py:31
        self.tk, self.dtk = reference_time_grid(ts, time_grid)
        self.P = fx_spec.P
        self.x_numel = S * B * fx_spec.H * fx_spec.W * fx_spec.C
        self.sample_offset = 0   # global index of local sample 0 (ranks sharding the MC samples)
        self._dev = {}

    def set_step(self, t, h):
        """One-iteration grid (t_k, dt_k) = (t, h) for the step-at-a-time callers (the stochastic-adjoint route)."""
        self.tk, self.dtk = np.asarray([t], dtype=np.float32), np.asarray([h], dtype=np.float32)
        self._dev = {}

    def grid_on(self, device):
        key = str(device)
        if key not in self._dev:
            # whole-solve grids recur every step (a BlockConfig is built per call): their device copies are shared process-wide,
            # so that no step pays a pageable host -> device copy (which stalls the host until the stream has drained)
            gkey = (key, self.tk.tobytes(), self.dtk.tobytes()) if len(self.tk) > 1 else None
            if gkey is not None and gkey in _GRID_CACHE:
                self._dev[key] = _GRID_CACHE[gkey]
            else:
                self._dev[key] = (torch.from_numpy(self.tk).to(device), torch.from_numpy(self.dtk).to(device))
                if gkey is not None and len(_GRID_CACHE) < 256:
                    _GRID_CACHE[gkey] = self._dev[key]
        return self._dev[key]

    def xdesc(self):
        d = _lib.XpathDesc()
        d.S, d.B, d.nit = self.S, self.B, self.nit
        layers = self.fx.conv_layer_structs()
        d.nlayers = len(layers)
        d.act = _lib.ACT[self.fx.actfn]
        d.math = _lib.MATH[self.math]
        d.P = self.P
        d.x_numel = self.x_numel
        for i, L in enumerate(layers):
            d.layers[i] = L
            d.beta_off[i] = self.fx.layers[i].get("beta_off", -1)
        return d

    def wdesc(self, seed, stream_id, w0_per_sample):
        d = _lib.WsdeDesc()
        d.S, d.nit, d.stl = self.S, self.nit, int(self.stl)
        d.w0_per_sample = int(w0_per_sample)
        d.sigma, d.kl_weight, d.ou_coef = self.sigma, self.kl_weight, self.ou_coef
        d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        d.stream_id = int(stream_id) & 0xFFFFFFFF
        d.sample_offset = int(self.sample_offset) & 0xFFFFFFFF
        return d


def wsde_forward(cfg, w0, layers, dW, seed, stream_id, w_stale, betas=()):
    """K1 forward.  Returns (wtraj [S,nit+1,P], z1 (opaque), kl [S])."""
    L = _lib.lib()
    dev = w0.device
    S, nit, P = cfg.S, cfg.nit, cfg.P
    w0_per_sample = w0.dim() == 2
    d = cfg.wdesc(seed, stream_id, w0_per_sample)
    fp = _fw_struct(layers, P, cfg.fw["hidden_dims"], cfg.fw["actfn"], betas)
    wtraj = torch.empty(S, nit + 1, P, device=dev, dtype=torch.float32)
    z1 = torch.empty(nit * S * cfg.fw["hidden_dims"][0], device=dev, dtype=torch.float32)
    kl = torch.empty(S, device=dev, dtype=torch.float32)
    cfg.kltraj = None
    if getattr(cfg, "want_kltraj", False):   # the kl rows of `ys` (sdeint.py:138) for callers of the bare solver
        cfg.kltraj = torch.empty(S, nit + 1, P, device=dev, dtype=torch.float32)
        d.d_kltraj = cfg.kltraj.data_ptr()
    nbytes = L.bsde_wsde_workspace_bytes(C.byref(d), C.byref(fp))
    ws = _lib.workspace(nbytes, dev, "wsde")
    tk, dtk = cfg.grid_on(dev)
    if dW is not None and tuple(dW.shape) != (S, nit, P):
        raise ValueError(f"increments must have shape {(S, nit, P)}, got {tuple(dW.shape)}")
    st = L.bsde_wsde_fwd(C.byref(d), C.byref(fp), _lib.ptr(w0), _lib.ptr(tk), _lib.ptr(dtk), _lib.ptr(dW),
                         _lib.ptr(w_stale) if cfg.stl else None, _lib.ptr(wtraj), _lib.ptr(z1), _lib.ptr(kl),
                         C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "bsde_wsde_fwd")
    return wtraj, z1, kl


def wsde_backward(cfg, layers, wtraj, z1, gw_traj, gkl, gwT, w0_per_sample, betas=()):
    """K1 reverse pass.  Returns (gw0, gradients of the layer tensors [+ the swish beta gradients as a last list entry])."""
    L = _lib.lib()
    dev = wtraj.device
    S, P = cfg.S, cfg.P
    d = cfg.wdesc(0, 0, w0_per_sample)
    fp = _fw_struct(layers, P, cfg.fw["hidden_dims"], cfg.fw["actfn"], betas)
    glayers = [tuple(torch.empty_like(t) for t in lay) for lay in layers]
    gbetas = [torch.empty_like(t) for t in betas]
    gfp = _fw_struct(glayers, P, cfg.fw["hidden_dims"], cfg.fw["actfn"], gbetas)
    gw0 = torch.empty((S, P) if w0_per_sample else (P,), device=dev, dtype=torch.float32)
    nbytes = L.bsde_wsde_workspace_bytes(C.byref(d), C.byref(fp))
    ws = _lib.workspace(nbytes, dev, "wsde")
    tk, dtk = cfg.grid_on(dev)
    st = L.bsde_wsde_bwd(C.byref(d), C.byref(fp), _lib.ptr(tk), _lib.ptr(dtk), _lib.ptr(wtraj), _lib.ptr(z1),
                         _lib.ptr(gw_traj), _lib.ptr(gkl), _lib.ptr(gwT), _lib.ptr(gw0), C.byref(gfp),
                         C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "bsde_wsde_bwd")
    if gbetas:
        glayers = glayers + [tuple(gbetas)]   # flattened by the callers in the order of split_fw_flat
    return gw0, glayers


def xpath_forward(cfg, x, wtraj, keep_acts):
    """K2 forward scan.  x: [S,B,H,W,C].  Returns (xs [nit+1, S,B,H,W,C] (slot 0 = x), acts or None)."""
    L = _lib.lib()
    d = cfg.xdesc()
    xs = torch.empty((cfg.nit + 1,) + tuple(x.shape), device=x.device, dtype=torch.float32)
    xs[0].copy_(x)
    acts = None
    if keep_acts:
        acts = torch.empty(L.bsde_xpath_acts_bytes(C.byref(d)) // 4, device=x.device, dtype=torch.float32)
    nbytes = L.bsde_xpath_workspace_bytes(C.byref(d), 0)
    ws = _lib.workspace(nbytes, x.device, "xpath")
    st = L.bsde_xpath_fwd(C.byref(d), _lib.host_floats(cfg.tk), _lib.host_floats(cfg.dtk), _lib.ptr(wtraj),
                          _lib.ptr(xs), _lib.ptr(acts), C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "bsde_xpath_fwd")
    return xs, acts


def xpath_backward(cfg, xs, wtraj, gx, acts=None):
    """K3: reverse pass over the x rows.  Returns (gx0, gw_traj [S,nit,P])."""
    L = _lib.lib()
    d = cfg.xdesc()
    gx = gx.contiguous().clone()
    gw_traj = torch.empty(cfg.S, cfg.nit, cfg.P, device=xs.device, dtype=torch.float32)
    # 2: reverse pass WITH stored activations (deferred, iteration-batched weight gradients keep dZ / gx of every iteration)
    nbytes = L.bsde_xpath_workspace_bytes(C.byref(d), 2 if acts is not None else 1)
    ws = _lib.workspace(nbytes, xs.device, "xpath")
    st = L.bsde_xpath_bwd(C.byref(d), _lib.host_floats(cfg.tk), _lib.host_floats(cfg.dtk), _lib.ptr(wtraj),
                          _lib.ptr(xs), _lib.ptr(acts), _lib.ptr(gx), _lib.ptr(gw_traj), C.c_void_p(ws.data_ptr()), ws.numel(),
                          _lib.stream_ptr())
    _lib.check(st, "bsde_xpath_bwd")
    return gx, gw_traj


# Optional CUDA-event bookkeeping for bench.py: when PROFILE is a dict, every K1 / K2 / K3 call is
# bracketed by events on the launching (current) stream; bench.py sums them after the timed region.
PROFILE = None


class _Timed:
    def __init__(self, key):
        self.key = key

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.setdefault(self.key, []).append((self.e0, self.e1))


class SDEBNNSolve(torch.autograd.Function):
    """ys[-1] of sdeint_ito_fixed_grid(f_aug, g_aug, [x,w0,0], ts, ...) for S samples at once.

    forward(x [S,B,H,W,C], w0 [P] or [S,P], *fw tensors) -> (x_T, kl [S], wtraj [S,nit+1,P])"""

    @staticmethod
    def forward(ctx, cfg, dW, seed, stream_id, w_stale, x, w0, *fw_flat):
        layers, betas = split_fw_flat(cfg, fw_flat)
        x = x.contiguous()
        w0c = w0.contiguous()
        with _Timed("k1_fwd"):
            wtraj, z1, kl = wsde_forward(cfg, w0c, layers, dW, seed, stream_id, w_stale, betas)
        with _Timed("k2_fwd"):
            # under torch.no_grad() (evaluation) nothing will be differentiated, but needs_input_grad only reflects requires_grad and
            # grad mode is always off inside forward(): the caller records it in cfg.track before apply()
            xs, acts = xpath_forward(cfg, x, wtraj, keep_acts=(not cfg.remat) and getattr(cfg, "track", True) and any(ctx.needs_input_grad))
        ctx.cfg = cfg
        ctx.layers, ctx.betas = layers, betas
        ctx.w0_per_sample = w0.dim() == 2
        ctx.acts = acts
        ctx.save_for_backward(wtraj, z1, xs)
        ctx.mark_non_differentiable(wtraj)
        if getattr(cfg, "want_kltraj", False):
            cfg.xs = xs   # every x_k (the scan's checkpoints): the x rows of `ys`
        return xs[cfg.nit].clone(), kl, wtraj

    @staticmethod
    def backward(ctx, gx, gkl, _gwtraj):
        cfg = ctx.cfg
        wtraj, z1, xs = ctx.saved_tensors
        dev = xs.device
        if gx is None:
            gx = torch.zeros_like(xs[0])
        if gkl is None:
            gkl = torch.zeros(cfg.S, device=dev)
        with _Timed("k3_bwd"):
            gx0, gw_traj = xpath_backward(cfg, xs, wtraj, gx.to(torch.float32), ctx.acts)
            ctx.acts = None
        with _Timed("k1_bwd"):
            gw0, glayers = wsde_backward(cfg, ctx.layers, wtraj, z1, gw_traj, gkl.contiguous().float(), None,
                                         ctx.w0_per_sample, ctx.betas)
        flat = [t for lay in glayers for t in lay]
        return (None, None, None, None, None, gx0, gw0, *flat)


# ------------------------------------------------------------------------------------------------------------------
# fixed_grid=False: `sdeint_ito` for SDEBNN dynamics (brax/_impl/brax.py:114-116) -- variable-step forward solve on a
# queryable Brownian path and the stochastic-adjoint reverse pass (jaxsde/jaxsde/sde_vjp.py:13-47,137-192), O(1) memory
# ------------------------------------------------------------------------------------------------------------------
def ito_step_schedule(ts, dt, max_steps=1 << 22):
    """The steps the `while _t < target_t` loop of `_stochastic_integrate` takes (jaxsde/jaxsde/sdeint.py:114-128), in fp32:
    `next_t = min(_t + dt, ts[-1])`, so interior output times may be overshot.  Returns (t_k, h_k, t_{k+1}) per step and, for
    every output time after the first, how many steps have been taken when its row is emitted."""
    f32 = np.float32
    ts = [f32(v) for v in ts]
    dt = f32(dt)
    if not dt > 0:
        raise ValueError(f"dt must be > 0, got {dt}")
    steps, marks, _t = [], [], ts[0]
    for target in ts[1:]:
        while _t < target:
            nxt = min(f32(_t + dt), ts[-1])
            if not nxt > _t:
                raise ValueError(f"dt={dt} does not advance t={_t} in fp32")
            steps.append((_t, f32(nxt - _t), nxt))
            if len(steps) > max_steps:
                raise ValueError(f"more than {max_steps} steps of dt={dt} between {ts[0]} and {ts[-1]}")
            _t = nxt
        marks.append(len(steps))
    return steps, marks


LAST_RECONSTRUCTION = None


class SDEBNNAdjointSolve(torch.autograd.Function):
    """`sdeint_ito(f_aug, g_aug, [x, w0, 0], ts, rng, fw_params, method="euler_maruyama", rep)` for S samples at once
    (brax.py:114-116), one step per K1 / K2 call.  g_aug is state-independent (quirk Q2), so Milstein == Euler-Maruyama, the
    reflected drift of `time_reflect_ito` (sde_utils.py:30-37) is -f(y, -t), and one step of the augmented adjoint system
    (sde_vjp.py:17-43) at the reconstructed state Y and forward time tau is
        Y      <- Y - h f(Y, tau) + g (W(tau - h) - W(tau))            K1 / K2 step with dt = -h and the tree's increment
        a      <- a + h (df/dy)(Y, tau)^T a                             K3 / K1' of that step with dt = +h
        a_args <- a_args + h (df/dargs)(Y, tau)^T a
    i.e. the existing single-iteration launches; nothing but the final state is kept from the forward pass.

    forward(x [S,B,H,W,C], w0 [P] or [S,P], *fw tensors) -> (x_T, kl [S], w at the output times [S, len(ts), P])"""

    @staticmethod
    def forward(ctx, cfg, ts, dt, bm, w_stale, x, w0, *fw_flat):
        layers, betas = split_fw_flat(cfg, fw_flat)
        S, P = cfg.S, cfg.P
        steps, marks = ito_step_schedule(ts, dt)
        Yx = x.contiguous()
        Yw = (w0 if w0.dim() == 2 else w0.expand(S, P)).contiguous()
        kl = torch.zeros(S, device=x.device, dtype=torch.float32)
        rows, done = [Yw], 0
        W_t = bm(ts[0]) if steps else None
        with _Timed("adj_fwd"):
            for j, upto in enumerate(marks):
                for (t, h, tn) in steps[done:upto]:
                    W_n = bm(tn)
                    cfg.set_step(t, h)
                    wtraj, _, klk = wsde_forward(cfg, Yw, layers, (W_n - W_t).view(S, 1, P), 0, 0, w_stale, betas)
                    xs, _ = xpath_forward(cfg, Yx, wtraj, keep_acts=False)
                    Yx, Yw, W_t = xs[1], wtraj[:, 1].contiguous(), W_n
                    kl = kl + klk
                done = upto
                rows.append(Yw)
        ctx.cfg, ctx.ts, ctx.dt, ctx.bm, ctx.w_stale, ctx.layers, ctx.betas = cfg, ts, dt, bm, w_stale, layers, betas
        ctx.w0_per_sample = w0.dim() == 2
        ctx.save_for_backward(Yx, Yw)
        wout = torch.stack(rows, dim=1)
        ctx.mark_non_differentiable(wout)
        return Yx.clone(), kl, wout

    @staticmethod
    def backward(ctx, gx, gkl, _gw):
        cfg, layers, bm = ctx.cfg, ctx.layers, ctx.bm
        Yx, Yw = ctx.saved_tensors
        S, P = cfg.S, cfg.P
        dev = Yx.device
        a_x = (torch.zeros_like(Yx) if gx is None else gx.to(torch.float32)).contiguous()
        gkl = (torch.zeros(S, device=dev) if gkl is None else gkl.to(torch.float32)).contiguous()
        a_w = torch.zeros(S, P, device=dev, dtype=torch.float32)
        gacc = None
        rts = [-np.float32(v) for v in list(ctx.ts)[::-1]]  # time_reflect_ito, sde_utils.py:37
        steps, _ = ito_step_schedule(rts, ctx.dt)
        W_t = bm(-steps[0][0]) if steps else None
        with _Timed("adj_bwd"):
            for (r, h, rn) in steps:
                tau, W_n = -r, bm(-rn)
                cfg.set_step(tau, -h)
                wtraj, z1, _ = wsde_forward(cfg, Yw, layers, (W_n - W_t).view(S, 1, P), 0, 0, ctx.w_stale, ctx.betas)
                xs, acts = xpath_forward(cfg, Yx, wtraj, keep_acts=not cfg.remat)   # remat: K3 recomputes them from Y
                cfg.set_step(tau, h)
                a_x, gw_traj = xpath_backward(cfg, xs, wtraj, a_x, acts)
                a_w, gl = wsde_backward(cfg, layers, wtraj, z1, gw_traj, gkl, a_w, True, ctx.betas)
                gl = [t for lay in gl for t in lay]
                if gacc is None:
                    gacc = gl
                else:
                    torch._foreach_add_(gacc, gl)
                Yx, Yw, W_t = xs[1], wtraj[:, 1].contiguous(), W_n
        if gacc is None:
            gacc = [torch.zeros_like(t) for lay in layers for t in lay] + [torch.zeros_like(t) for t in ctx.betas]
        global LAST_RECONSTRUCTION
        LAST_RECONSTRUCTION = (Yx, Yw)  # y0 as the reverse solve reconstructs it (`y0_rec` of sde_vjp.py:162; diagnostics / tests)
        gw0 = a_w if ctx.w0_per_sample else a_w.sum(0)
        return (None, None, None, None, None, a_x, gw0, *gacc)


def zero_fw_layers(P, device):
    """f_w == 0 (w_drift=False, brax.py:56-57): a P->1->P stack with zero weights."""
    z = lambda *s: torch.zeros(*s, device=device)
    return [(z(P, 1), z(1), z(1), z(1), z(1)), (z(1, P), z(P), z(P), z(P), z(P))]

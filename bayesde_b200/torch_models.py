"""torch front end: `SDENet`, `make_y_net`, `make_w_net` of
torchbnn/_impl/models.py:25-80,105-204 on the kernels of libbsde.so.

The reference integrates ONE Stratonovich SDE over the flat state `[y (B*C*H*W), w (P), logqp (1)]` with torchsde's
fixed-step solvers (models.py:184-197).  The `[w, logqp]` rows never read `y`, so they are integrated first for all
solver stages in one cooperative launch (`bsde_wsde_fwd` with stage tables: plain-Linear w-net, `fw = nn - w`,
`fl = |nn|^2 / sigma^2`), then the `y` rows consume `w` stage by stage through `bsde_chain_fwd` (the whole `y_net`:
every group plus the `ConvDownsample` layers, time channel FIRST, OIHW / IOHW kernels read in place from the flat vector
through strides) and `bsde_nchw_reinterp_axpy` (the flat-NCHW-order state update).  The reverse pass is hand written
(`bsde_chain_bwd` per evaluation in reverse order, then `bsde_wsde_bwd`): backprop through the solver, like
`loss.backward()` through `torchsde.sdeint` (examples/torch/sdebnn_classification.py:34-46).

Semantics follow the INTENDED flat-state reading of `SDENet` (the class as shipped raises in `split`/`cat`: SURVEY
quirk Q7).  `adjoint=True` selects the O(1)-memory reverse pass of `sdeint_adjoint` (SDENetAdjointSolve); `adaptive=True` runs torchsde's
step-doubling controller and replays the accepted grid (with `adjoint=True`: fixed-step reverse solve from the adaptive solve's
final state; `adjoint_adaptive=True`: the controller also drives the reverse solve).  There is no CPU path.
"""
import ctypes as C
import math

import numpy as np
import torch
from torch import nn

from . import _lib

CHAIN_MAX_LAYERS = 40


class ChainDesc(C.Structure):
    _fields_ = [("S", C.c_int32), ("B", C.c_int32), ("nlayers", C.c_int32), ("act", C.c_int32), ("math", C.c_int32),
                ("act_after", C.c_int32 * CHAIN_MAX_LAYERS), ("layers", _lib.ConvLayer * CHAIN_MAX_LAYERS)]


_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        vp = C.c_void_p
        for name in ("bsde_chain_workspace_bytes", "bsde_chain_acts_floats", "bsde_chain_out_offset", "bsde_chain_fwd", "bsde_chain_bwd",
                     "bsde_nchw_reinterp_axpy"):
            if not hasattr(L, name):
                raise _lib.BsdeError(f"libbsde.so does not export {name}")
        L.bsde_chain_workspace_bytes.restype = C.c_size_t
        L.bsde_chain_workspace_bytes.argtypes = [C.POINTER(ChainDesc), C.c_int32]
        L.bsde_chain_acts_floats.restype = C.c_size_t
        L.bsde_chain_acts_floats.argtypes = [C.POINTER(ChainDesc)]
        L.bsde_chain_out_offset.restype = C.c_size_t
        L.bsde_chain_out_offset.argtypes = [C.POINTER(ChainDesc)]
        L.bsde_chain_fwd.argtypes = [C.POINTER(ChainDesc), C.c_float, vp, C.c_int64, vp, vp, vp, C.c_size_t, vp]
        L.bsde_chain_bwd.argtypes = [C.POINTER(ChainDesc), C.c_float, vp, C.c_int64, vp, vp, vp, vp, C.c_int32, vp, C.c_int64, vp,
                                     C.c_size_t, vp]
        L.bsde_nchw_reinterp_axpy.argtypes = [C.c_int64] + [C.c_int32] * 6 + [C.c_float, vp, vp, vp, vp]
        _bound = True
    return L


# ------------------------------------------------------------------------------------------------
# pytree helpers (torchbnn/_impl/utils.py:165-193)
# ------------------------------------------------------------------------------------------------
def ravel_pytree(tree):
    """Flatten a nested list of tensors depth first; returns (flat, unravel)."""
    if torch.is_tensor(tree):
        shape = tree.shape
        return tree.reshape(-1), (lambda v: v.reshape(shape))
    parts = [ravel_pytree(e) for e in tree]
    sizes = [f.numel() for f, _ in parts]

    def unravel(v):
        return [un(piece) for piece, (_, un) in zip(v.split(sizes), parts)]

    flat = torch.cat([f for f, _ in parts]) if parts else torch.tensor([])
    return flat, unravel


# ------------------------------------------------------------------------------------------------
# y_net: models.py:25-54, diffeq_layers.py:280-311 (mode 0), ConvDownsample :230-239
# ------------------------------------------------------------------------------------------------
class YNet:
    """Layout + evaluation handle of the hidden-state drift net.  Parameters are NOT owned here (`explicit_params=True`,
    diffeq_layers.py:46): every call takes the flat vector (or its unravelled pytree)."""

    def __init__(self, input_size, blocks=(2, 2, 2), activation="softplus", hidden_width=128, aug_dim=0):
        if activation not in ("softplus", "tanh", "elu"):
            raise NotImplementedError(f"y_net activation {activation!r} is not available in the fused path")
        self.activation = activation
        c, h, w = input_size[0] + aug_dim, input_size[1], input_size[2]
        self.input_size = (c, h, w)
        self.layers, off = [], 0

        def conv(cin, cout, hin, win, stride=1, padding=0, transpose=False, output_padding=0, act_after=True):
            nonlocal off
            if transpose:  # o = i*stride - padding + k  (F.conv_transpose2d), weight (cin+1, cout, 3, 3)
                hout, wout = (hin - 1) * stride - 2 * padding + 3 + output_padding, (win - 1) * stride - 2 * padding + 3 + output_padding
                sn, kn, o, sd = 1, -1, padding, stride
                ks, ns = cout * 9, 9
                fan_in = cout * 9
            else:          # i = o*stride + k - padding   (F.conv2d), weight (cout, cin+1, 3, 3)
                hout, wout = (hin + 2 * padding - 3) // stride + 1, (win + 2 * padding - 3) // stride + 1
                sn, kn, o, sd = stride, 1, -padding, 1
                ks, ns = 9, (cin + 1) * 9
                fan_in = (cin + 1) * 9
            nw = (cin + 1) * cout * 9
            self.layers.append(dict(cin=cin, cout=cout, hin=hin, win=win, hout=hout, wout=wout, sn=sn, kn=kn, off=o, sd=sd,
                                    w_off=off, b_off=off + nw, ks=ks, ns=ns, act_after=act_after, transpose=transpose,
                                    shape=(cin + 1, cout, 3, 3) if transpose else (cout, cin + 1, 3, 3), fan_in=fan_in))
            off += nw + cout
            return hout, wout

        for i, nb in enumerate(blocks, 1):
            for j in range(1, nb + 1):
                h1, w1 = conv(c, hidden_width, h, w, padding=1)
                h2, w2 = conv(hidden_width, hidden_width, h1, w1, stride=2, padding=1)
                h3, w3 = conv(hidden_width, hidden_width, h2, w2, stride=2, transpose=True, output_padding=1)
                h4, w4 = conv(hidden_width, c, h3, w3, act_after=(i < len(blocks) or j < nb))
                if (h4, w4) != (h, w):
                    raise AssertionError(f"block output {(h4, w4)} != input {(h, w)}. Check nblocks for dataset divisibility.")
            if i < len(blocks):
                h, w = conv(c, 4 * c, h, w, stride=2, padding=1, act_after=False)
                c = 4 * c
        self.P = off
        self.output_size = (c, h, w)
        if len(self.layers) > CHAIN_MAX_LAYERS:
            raise NotImplementedError(f"{len(self.layers)} conv layers > {CHAIN_MAX_LAYERS}")
        if c * h * w != self.input_size[0] * self.input_size[1] * self.input_size[2]:
            raise AssertionError("Check nblocks for dataset divisibility.")   # models.py:167

    # -- reference surface ---------------------------------------------------------------------
    def make_initial_params(self, generator=None):
        """[[weight, bias] per conv, [] per activation] with nn.Conv2d / nn.ConvTranspose2d default initialisation
        (DiffEqSequential.make_initial_params, diffeq_layers.py:58-59)."""
        tree = []
        for d in self.layers:
            bound = 1.0 / math.sqrt(d["fan_in"])
            W = (torch.rand(d["shape"], generator=generator) * 2 - 1) * bound
            b = (torch.rand(d["cout"], generator=generator) * 2 - 1) * bound
            tree.append([W, b])
            if d["act_after"]:
                tree.append([])
        return tree

    def chain_desc(self, S, B, math_mode):
        d = ChainDesc()
        d.S, d.B, d.nlayers = S, B, len(self.layers)
        d.act, d.math = _lib.ACT[self.activation], _lib.MATH[math_mode]
        for i, lay in enumerate(self.layers):
            c = d.layers[i]
            c.cin, c.cout, c.ksize = lay["cin"], lay["cout"], 3
            c.hin, c.win, c.hout, c.wout = lay["hin"], lay["win"], lay["hout"], lay["wout"]
            c.sn, c.kn, c.off, c.sd = lay["sn"], lay["kn"], lay["off"], lay["sd"]
            c.has_time, c.time_first = 1, 1          # utils.channel_cat: time channel first
            c.w_off, c.b_off = lay["w_off"], lay["b_off"]
            c.w_tap_stride, c.w_k_stride, c.w_n_stride = 1, lay["ks"], lay["ns"]
            d.act_after[i] = int(lay["act_after"])
        return d

    def __call__(self, t, y, params, math_mode="fp32"):
        """y_net(t, y, params): y (B, C, H, W) CUDA, params = flat vector or the pytree.  Forward only (the differentiable
        route is SDENet.forward)."""
        flat = params if torch.is_tensor(params) else ravel_pytree(params)[0]
        flat = flat.detach().float().contiguous()
        B = y.shape[0]
        x = y.detach().float().permute(0, 2, 3, 1).contiguous()
        desc = self.chain_desc(1, B, math_mode)
        acts = chain_forward(desc, float(t), flat, self.P, x)
        c, h, w = self.output_size
        out = acts[_L().bsde_chain_out_offset(C.byref(desc)):][:B * c * h * w].reshape(B, h, w, c)
        return out.permute(0, 3, 1, 2).contiguous()


def make_y_net(input_size, blocks=(2, 2, 2), activation="softplus", verbose=False, explicit_params=True, hidden_width=128,
               aug_dim=0):
    """models.py:25-54 -> (y_net, output_size)."""
    if not explicit_params:
        raise NotImplementedError("explicit_params=False (BaselineYNet) is outside the SDE-BNN hot path")
    net = YNet(input_size, blocks, activation, hidden_width, aug_dim)
    return net, net.output_size


# ------------------------------------------------------------------------------------------------
# w_net: models.py:57-80
# ------------------------------------------------------------------------------------------------
class _WNet(nn.Module):
    """Parameter container with the reference's state_dict keys (`layers.{0,2,..}.weight/bias` for inhomogeneous=True,
    `module.{0,2,..}` otherwise).  Both variants ignore t (diffeq_layers.py:65-72,169-175)."""

    def __init__(self, sizes, activation, inhomogeneous):
        super().__init__()
        acts = {"tanh": nn.Tanh, "softplus": nn.Softplus, "elu": nn.ELU}
        if activation not in acts:
            raise NotImplementedError(f"w_net activation {activation!r} is not available in the fused path")
        mods = []
        for i, (din, dout) in enumerate(zip(sizes[:-1], sizes[1:]), 1):
            mods.append(nn.Linear(din, dout))
            if i + 1 < len(sizes):
                mods.append(acts[activation]())
            else:  # last layer needs zero initialisation (models.py:67-69)
                nn.init.zeros_(mods[-1].weight)
                nn.init.zeros_(mods[-1].bias)
        if inhomogeneous:
            self.layers = nn.ModuleList(mods)
        else:
            self.module = nn.Sequential(*mods)
        self.activation = activation
        self.hidden_sizes = tuple(sizes[1:-1])

    def linears(self):
        mods = self.layers if hasattr(self, "layers") else self.module
        return [m for m in mods if isinstance(m, nn.Linear)]

    def forward(self, t, w):
        raise NotImplementedError("w_net is evaluated inside libbsde.so (bsde_wsde_fwd); use SDENet.forward")


def make_w_net(in_features, hidden_sizes=(1, 64, 1), activation="softplus", inhomogeneous=True):
    sizes = (in_features,) + tuple(hidden_sizes) + (in_features,)
    if len(hidden_sizes) < 1 or hidden_sizes[0] > _lib.FW_MAX_EDGE or hidden_sizes[-1] > _lib.FW_MAX_EDGE:
        raise NotImplementedError(f"first / last hidden widths must be <= {_lib.FW_MAX_EDGE}, got {hidden_sizes}")
    return _WNet(sizes, activation, inhomogeneous)


# ------------------------------------------------------------------------------------------------
# kernel wrappers
# ------------------------------------------------------------------------------------------------
def chain_forward(desc, t, w, w_sample_stride, x, acts=None):
    L = _L()
    n = L.bsde_chain_acts_floats(C.byref(desc))
    if acts is None:
        acts = torch.empty(n, device=x.device, dtype=torch.float32)
    nbytes = L.bsde_chain_workspace_bytes(C.byref(desc), 0)
    ws = _lib.workspace(nbytes, x.device, "chain")
    st = L.bsde_chain_fwd(C.byref(desc), t, _lib.ptr(w), w_sample_stride, _lib.ptr(x), _lib.ptr(acts), C.c_void_p(ws.data_ptr()),
                          ws.numel(), _lib.stream_ptr())
    _lib.check(st, "bsde_chain_fwd")
    return acts


def chain_backward(desc, t, w, w_sample_stride, x, acts, gf, gx, accumulate, gw, gw_sample_stride):
    L = _L()
    nbytes = L.bsde_chain_workspace_bytes(C.byref(desc), 1)
    ws = _lib.workspace(nbytes, x.device, "chain")
    st = L.bsde_chain_bwd(C.byref(desc), t, _lib.ptr(w), w_sample_stride, _lib.ptr(x), _lib.ptr(acts), _lib.ptr(gf), _lib.ptr(gx),
                          int(accumulate), _lib.ptr(gw), gw_sample_stride, C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "bsde_chain_bwd")


def reinterp_axpy(out, out_chw, base, a, src, src_chw):
    """out[NHWC of out_chw] = base + a * (src viewed in flat NCHW order as out_chw)."""
    nimg = out.numel() // (out_chw[0] * out_chw[1] * out_chw[2])
    st = _L().bsde_nchw_reinterp_axpy(nimg, *out_chw, *src_chw, a, _lib.ptr(base), _lib.ptr(src), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(st, "bsde_nchw_reinterp_axpy")


def _plain_fw_struct(klayers, P, hidden, actfn):
    """bsde_fw_params for a plain-Linear w-net; klayers in kernel layout: (W1 [P,d1], b1), (Wm [din,dout], bm).., (W4 [dL,P], b4)."""
    fp = _lib.FwParams()
    fp.P, fp.nhid, fp.act, fp.plain = P, len(hidden), _lib.ACT[actfn], 1
    for i, d in enumerate(hidden):
        fp.dims[i] = d
    fp.W1, fp.b1 = _lib.ptr(klayers[0][0]).value, _lib.ptr(klayers[0][1]).value
    for i, (W, b) in enumerate(klayers[1:-1]):
        fp.Wm[i], fp.bm[i] = _lib.ptr(W).value, _lib.ptr(b).value
    fp.W4, fp.b4 = _lib.ptr(klayers[-1][0]).value, _lib.ptr(klayers[-1][1]).value
    return fp


class SolveConfig:
    """Static description of one SDENet solve: stage tables of the fixed-step scheme (include/bsde.h).  `grid` = (t_k, h_k) lists
    replaces the uniform grid t0 + k h (the accepted steps of an adaptive solve)."""

    def __init__(self, ynet, S, B, nsteps, h, method, sigma, hidden, w_act, math_mode, remat, t0=0.0, grid=None):
        if method not in ("midpoint", "heun", "euler_heun", "milstein"):
            raise ValueError(f"Unknown method: {method}")
        f32 = np.float32
        if grid is not None:
            tk, hs = [f32(t) for t in grid[0]], [f32(v) for v in grid[1]]
            nsteps, h = len(tk), float(hs[0]) if len(hs) else 0.0
        else:
            hh = f32(h)
            tk, hs = [f32(t0) + f32(k) * hh for k in range(nsteps)], [hh] * nsteps
        self.ynet, self.S, self.B, self.n, self.h, self.method = ynet, S, B, nsteps, float(h), method
        self.hs = [float(v) for v in hs]
        self.uniform = grid is None
        self.sigma, self.hidden, self.w_act, self.math, self.remat = float(sigma), tuple(hidden), w_act, math_mode, remat
        half = f32(0.5)
        sq = [math.sqrt(float(v)) for v in hs]
        pairs = lambda a, b: np.asarray([v for k in range(nsteps) for v in (a(k), b(k))], dtype=f32)
        if method == "midpoint":   # torchsde midpoint: f at (t, y) and (t + h/2, y'), the same increment I in both stages
            self.nit = 2 * nsteps
            self.tk = pairs(lambda k: tk[k], lambda k: tk[k] + half * hs[k])
            self.dtk = pairs(lambda k: half * hs[k], lambda k: hs[k])
            self.dtkl = pairs(lambda k: 0.0, lambda k: hs[k])
            self.nscale = pairs(lambda k: 0.5 * sq[k], lambda k: sq[k])
            self.nidx = np.asarray([k for k in range(nsteps) for _ in (0, 1)], dtype=np.int32)
            self.back = np.asarray([0, 1] * nsteps, dtype=np.int32)
        elif method == "heun":     # torchsde heun (trapezoidal): y' = y + h f(t, y) + g I; y1 = y + h/2 (f(t, y) + f(t + h, y')) + g I
            self.nit = 2 * nsteps      # = (y + y')/2 + h/2 f(t + h, y') + g I/2: the corrector starts from the average (back = 2)
            self.tk = pairs(lambda k: tk[k], lambda k: tk[k] + hs[k])
            self.dtk = pairs(lambda k: hs[k], lambda k: half * hs[k])
            self.dtkl = pairs(lambda k: half * hs[k], lambda k: half * hs[k])
            self.nscale = pairs(lambda k: sq[k], lambda k: 0.5 * sq[k])
            self.nidx = np.asarray([k for k in range(nsteps) for _ in (0, 1)], dtype=np.int32)
            self.back = np.asarray([0, 2] * nsteps, dtype=np.int32)
        else:                      # constant diagonal g: euler_heun / milstein reduce to y + h f + g I
            self.nit = nsteps
            self.tk = np.asarray(tk, dtype=f32)
            self.dtk = np.asarray(hs, dtype=f32)
            self.dtkl = self.nscale = self.nidx = self.back = None
        self._dev = {}

    def on(self, device):
        key = str(device)
        if key not in self._dev:
            t = lambda a: None if a is None else torch.from_numpy(a).to(device)
            self._dev[key] = tuple(t(a) for a in (self.tk, self.dtk, self.dtkl, self.nscale, self.nidx, self.back))
        return self._dev[key]

    def wdesc(self, device, seed, stream_id):
        d = _lib.WsdeDesc()
        d.S, d.nit, d.stl, d.w0_per_sample = self.S, self.nit, 0, 0
        # fw = nn - w (models.py:165), fl = |nn|^2 / sigma^2 (:166): ou_coef -1 => u = nn / sigma, integrand u^2
        d.sigma, d.kl_weight, d.ou_coef = self.sigma, 1.0, -1.0
        d.seed, d.stream_id = int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id) & 0xFFFFFFFF
        _, _, dtkl, nscale, nidx, back = self.on(device)
        if dtkl is not None:
            d.d_dtkl, d.d_nscale, d.d_nidx, d.d_back = dtkl.data_ptr(), nscale.data_ptr(), nidx.data_ptr(), back.data_ptr()
        return d

    def stage_increments(self, I):
        """Per-stage noise [S, nit, P] from per-step increments I [S, nsteps, P]."""
        if self.method == "midpoint":
            pair = (0.5 * I, I)
        elif self.method == "heun":
            pair = (I, 0.5 * I)
        else:
            return I.contiguous()
        return torch.stack(pair, dim=2).reshape(I.shape[0], 2 * I.shape[1], I.shape[2]).contiguous()


class SDENetSolve(torch.autograd.Function):
    """(y_T, logqp_state [S], wtraj) of sdeint(SDENet, [y, w0, 0], ts=[0,1], bm, method, dt) for S weight samples.

    forward(cfg, increments or None, seed, stream_id, x [S,B,H,W,C] (NHWC storage), w0 [P], W1,b1, Wm,bm.., W4,b4 (kernel layout))"""

    @staticmethod
    def forward(ctx, cfg, I, seed, stream_id, x, w0, *flat):
        L = _L()
        dev = x.device
        klayers = [(flat[i].contiguous(), flat[i + 1].contiguous()) for i in range(0, len(flat), 2)]
        S, B, P, nit, n = cfg.S, cfg.B, cfg.ynet.P, cfg.nit, cfg.n
        tk, dtk, *_ = cfg.on(dev)
        # ---- K1: the [w, logqp] rows for every stage
        d = cfg.wdesc(dev, seed, stream_id)
        fp = _plain_fw_struct(klayers, P, cfg.hidden, cfg.w_act)
        wtraj = torch.empty(S, nit + 1, P, device=dev, dtype=torch.float32)
        z1 = torch.empty(nit * S * cfg.hidden[0], device=dev, dtype=torch.float32)
        kl = torch.empty(S, device=dev, dtype=torch.float32)
        ws = _lib.workspace(L.bsde_wsde_workspace_bytes(C.byref(d), C.byref(fp)), dev, "wsde")
        dW = None
        if I is not None:
            if tuple(I.shape) != (S, n, P):
                raise ValueError(f"increments must have shape {(S, n, P)}, got {tuple(I.shape)}")
            dW = cfg.stage_increments(I.float())
        st = L.bsde_wsde_fwd(C.byref(d), C.byref(fp), _lib.ptr(w0.contiguous()), _lib.ptr(tk), _lib.ptr(dtk), _lib.ptr(dW), None,
                             _lib.ptr(wtraj), _lib.ptr(z1), _lib.ptr(kl), C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr())
        _lib.check(st, "bsde_wsde_fwd")
        # ---- y rows
        x = x.contiguous()
        desc = cfg.ynet.chain_desc(S, B, cfg.math)
        nacts, ooff = L.bsde_chain_acts_floats(C.byref(desc)), L.bsde_chain_out_offset(C.byref(desc))
        keep = (not cfg.remat) and getattr(cfg, "track", True) and any(ctx.needs_input_grad)   # cfg.track: grad mode at the call site
        store = torch.empty(nit if keep else 1, nacts, device=dev, dtype=torch.float32)
        ys = torch.empty((n + 1,) + tuple(x.shape), device=dev, dtype=torch.float32)
        ys[0].copy_(x)
        mid, heun = cfg.method == "midpoint", cfg.method == "heun"
        two = mid or heun
        ymid = torch.empty((n,) + tuple(x.shape), device=dev, dtype=torch.float32) if two else None   # y' of every step
        chw_in, chw_out = cfg.ynet.input_size, cfg.ynet.output_size
        wss = (nit + 1) * P
        tkh = cfg.tk
        for k in range(n):
            hk = cfg.hs[k]
            j = 2 * k if two else k
            a0 = store[j if keep else 0]
            chain_forward(desc, float(tkh[j]), wtraj[0, j], wss, ys[k], a0)
            if heun:
                reinterp_axpy(ymid[k], chw_in, ys[k], hk, a0[ooff:], chw_out)                      # y' = y + h f(t, y)
                if k == 0:
                    yavg = torch.empty_like(ys[0])
                reinterp_axpy(yavg, chw_in, ys[k], 0.5 * hk, a0[ooff:], chw_out)                  # (y + y')/2 = y + h/2 f(t, y)
                a1 = store[j + 1 if keep else 0]
                chain_forward(desc, float(tkh[j + 1]), wtraj[0, j + 1], wss, ymid[k], a1)
                reinterp_axpy(ys[k + 1], chw_in, yavg, 0.5 * hk, a1[ooff:], chw_out)              # + h/2 f(t + h, y')
            elif mid:
                reinterp_axpy(ymid[k], chw_in, ys[k], 0.5 * hk, a0[ooff:], chw_out)
                a1 = store[j + 1 if keep else 0]
                chain_forward(desc, float(tkh[j + 1]), wtraj[0, j + 1], wss, ymid[k], a1)
                reinterp_axpy(ys[k + 1], chw_in, ys[k], hk, a1[ooff:], chw_out)
            else:
                reinterp_axpy(ys[k + 1], chw_in, ys[k], hk, a0[ooff:], chw_out)
        ctx.cfg, ctx.klayers, ctx.desc = cfg, klayers, desc
        ctx.store = store if keep else None
        ctx.save_for_backward(wtraj, z1, ys, ymid if two else ys)
        ctx.mark_non_differentiable(wtraj)
        return ys[n].clone(), kl, wtraj

    @staticmethod
    def backward(ctx, gy, gkl, _gw):
        L = _L()
        cfg, desc, klayers = ctx.cfg, ctx.desc, ctx.klayers
        wtraj, z1, ys, ymid = ctx.saved_tensors
        dev = ys.device
        S, P, nit, n = cfg.S, cfg.ynet.P, cfg.nit, cfg.n
        mid = cfg.method == "midpoint"
        gy = torch.zeros_like(ys[0]) if gy is None else gy.float().contiguous().clone()
        gkl = torch.zeros(S, device=dev) if gkl is None else gkl.float().contiguous()
        nacts, ooff = L.bsde_chain_acts_floats(C.byref(desc)), L.bsde_chain_out_offset(C.byref(desc))
        chw_in, chw_out = cfg.ynet.input_size, cfg.ynet.output_size
        nf = S * cfg.B * chw_out[0] * chw_out[1] * chw_out[2]
        gf = torch.empty(nf, device=dev, dtype=torch.float32)
        gyp = torch.empty_like(gy)
        gw_traj = torch.empty(S, nit, P, device=dev, dtype=torch.float32)
        scratch = None if ctx.store is not None else torch.empty(nacts, device=dev, dtype=torch.float32)
        wss, gss = (nit + 1) * P, nit * P
        tkh = cfg.tk

        def acts_of(j, x):
            if ctx.store is not None:
                return ctx.store[j]
            return chain_forward(desc, float(tkh[j]), wtraj[0, j], wss, x, scratch)   # recompute (remat)

        heun = cfg.method == "heun"
        for k in range(n - 1, -1, -1):
            hk = cfg.hs[k]
            if heun:
                j = 2 * k
                # y_{k+1} = y_k + h/2 R(f(t, y_k, w_k)) + h/2 R(f(t + h, y', w'))
                reinterp_axpy(gf, chw_out, None, 0.5 * hk, gy, chw_in)
                chain_backward(desc, float(tkh[j + 1]), wtraj[0, j + 1], wss, ymid[k], acts_of(j + 1, ymid[k]), gf, gyp, False,
                               gw_traj[0, j + 1], gss)
                # y' = y_k + h R(f(t, y_k, w_k)): f(t, y_k) receives h * dL/dy' + h/2 * dL/dy_{k+1}
                gf2 = torch.empty_like(gf)
                reinterp_axpy(gf2, chw_out, gf, hk, gyp, chw_in)
                gy.add_(gyp)
                chain_backward(desc, float(tkh[j]), wtraj[0, j], wss, ys[k], acts_of(j, ys[k]), gf2, gy, True, gw_traj[0, j], gss)
            elif mid:
                j = 2 * k
                # y_{k+1} = y_k + h R(f(t + h/2, y', w'))
                reinterp_axpy(gf, chw_out, None, hk, gy, chw_in)
                chain_backward(desc, float(tkh[j + 1]), wtraj[0, j + 1], wss, ymid[k], acts_of(j + 1, ymid[k]), gf, gyp, False,
                               gw_traj[0, j + 1], gss)
                # y' = y_k + h/2 R(f(t, y_k, w_k))
                reinterp_axpy(gf, chw_out, None, 0.5 * hk, gyp, chw_in)
                gy.add_(gyp)
                chain_backward(desc, float(tkh[j]), wtraj[0, j], wss, ys[k], acts_of(j, ys[k]), gf, gy, True, gw_traj[0, j], gss)
            else:
                reinterp_axpy(gf, chw_out, None, hk, gy, chw_in)
                chain_backward(desc, float(tkh[k]), wtraj[0, k], wss, ys[k], acts_of(k, ys[k]), gf, gy, True, gw_traj[0, k], gss)
        ctx.store = None
        # ---- K1 reverse
        d = cfg.wdesc(dev, 0, 0)
        fp = _plain_fw_struct(klayers, P, cfg.hidden, cfg.w_act)
        gl = [(torch.empty_like(W), torch.empty_like(b)) for W, b in klayers]
        gfp = _plain_fw_struct(gl, P, cfg.hidden, cfg.w_act)
        gw0 = torch.empty(P, device=dev, dtype=torch.float32)
        tk, dtk, *_ = cfg.on(dev)
        ws = _lib.workspace(L.bsde_wsde_workspace_bytes(C.byref(d), C.byref(fp)), dev, "wsde")
        st = L.bsde_wsde_bwd(C.byref(d), C.byref(fp), _lib.ptr(tk), _lib.ptr(dtk), _lib.ptr(wtraj), _lib.ptr(z1), _lib.ptr(gw_traj),
                             _lib.ptr(gkl), None, _lib.ptr(gw0), C.byref(gfp), C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr())
        _lib.check(st, "bsde_wsde_bwd")
        return (None, None, None, None, gy, gw0, *[t for pair in gl for t in pair])


class SDENetAdjointSolve(torch.autograd.Function):
    """`sdeint_adjoint(SDENet, [y, w0, 0], ts=[0,1], bm, method, dt)` (models.py:184-197 with `adjoint=True`): the forward solve keeps
    nothing but its final state; the reverse pass integrates the augmented system [y, w | a_y, a_w, a_logqp | a_params] from t1 back to
    t0 with the same fixed-step scheme on the same Brownian path run backwards (Li et al. 2020, Alg. 2; torchsde itself is absent).
    g = [0, sigma, 0] is constant, so per reverse step (forward time tau, step h, reverse increment dW~ = -I_k), midpoint:
        stage A at (tau, Z):            Z' = Z - h/2 f(tau, Z) + g dW~/2      a' = a + h/2 J(tau, Z)^T a
        stage B at (tau - h/2, Z'):     Z1 = Z - h   f(.., Z') + g dW~        a1 = a + h   J(.., Z')^T a'    a_par += h (df/dpar)(.., Z')^T a'
    Every stage is the existing single-iteration launches: bsde_wsde_fwd / bsde_chain_fwd with dt < 0 (reconstruction) and
    bsde_chain_bwd / bsde_wsde_bwd with dt > 0 at the reconstructed state (an Euler-like reverse step seeded with a or a')."""

    @staticmethod
    def forward(ctx, cfg, I, seed, stream_id, final, rev, x, w0, *flat):
        """`final` = (y_T, logqp state [S], w_T) of a forward solve made elsewhere on the same Brownian path (the adaptive loop), or
        None to run the fixed-step forward solve of `cfg` here.  `rev` = (tree, rtol, atol, dt, t1) makes the reverse solve
        adaptive (`adjoint_adaptive=True`) on that tree instead of stepping over `I`."""
        from .sdeint import philox_increments
        S, P, n = cfg.S, cfg.ynet.P, cfg.n
        if S != 1:
            raise NotImplementedError("sdeint_adjoint route: one weight sample per call (SDENet.forward always has S = 1)")
        if not cfg.uniform:
            raise NotImplementedError("sdeint_adjoint route: uniform step grids only")
        if I is None:  # the same path the in-kernel stream of the plain route draws: N(0, h) at Philox counter (p, k, s, stream)
            I = philox_increments(seed, stream_id, S, np.full(n, np.float32(cfg.h), dtype=np.float32), P, x.device)
        elif tuple(I.shape) != (S, n, P):
            raise ValueError(f"increments must have shape {(S, n, P)}, got {tuple(I.shape)}")
        I = I.float().contiguous()
        if final is None:
            cfg.track = False
            with torch.no_grad():
                yT, kl, wtraj = SDENetSolve.apply(cfg, I, 0, 0, x.detach(), w0.detach(), *[t.detach() for t in flat])
            wT = wtraj[0, cfg.nit].contiguous()
        else:
            yT, kl, wT = (t.detach() for t in final)
            yT, kl, wT = yT.contiguous(), kl.clone(), wT.contiguous()
            wtraj = wT[None, None]
        ctx.cfg, ctx.rev = cfg, rev
        ctx.klayers = [(flat[i].detach().contiguous(), flat[i + 1].detach().contiguous()) for i in range(0, len(flat), 2)]
        ctx.save_for_backward(yT, wT, I)
        ctx.mark_non_differentiable(wtraj)
        return yT.clone(), kl, wtraj

    @staticmethod
    def backward(ctx, gy, gkl, _gw):
        L = _L()
        cfg, klayers = ctx.cfg, ctx.klayers
        Zy, Zw, I = ctx.saved_tensors
        dev = Zy.device
        P, n, h = cfg.ynet.P, cfg.n, float(np.float32(cfg.h))
        mid = cfg.method == "midpoint"
        desc = cfg.ynet.chain_desc(1, cfg.B, cfg.math)
        nacts, ooff = L.bsde_chain_acts_floats(C.byref(desc)), L.bsde_chain_out_offset(C.byref(desc))
        chw_in, chw_out = cfg.ynet.input_size, cfg.ynet.output_size
        nf = cfg.B * chw_out[0] * chw_out[1] * chw_out[2]
        a_y = (torch.zeros_like(Zy) if gy is None else gy.float()).contiguous().clone()
        a_l = (torch.zeros(1, device=dev) if gkl is None else gkl.float()).contiguous()
        a_w = torch.zeros(P, device=dev, dtype=torch.float32)
        gacc = [torch.zeros_like(t) for pair in klayers for t in pair]
        acts = torch.empty(nacts, device=dev, dtype=torch.float32)
        gf = torch.empty(nf, device=dev, dtype=torch.float32)
        d = _lib.WsdeDesc()
        d.S, d.nit, d.stl, d.w0_per_sample = 1, 1, 0, 0
        d.sigma, d.kl_weight, d.ou_coef = cfg.sigma, 1.0, -1.0
        fp = _plain_fw_struct(klayers, P, cfg.hidden, cfg.w_act)
        ws = _lib.workspace(L.bsde_wsde_workspace_bytes(C.byref(d), C.byref(fp)), dev, "wsde")
        kl_scratch = torch.empty(1, device=dev, dtype=torch.float32)
        f32t = lambda v: torch.tensor([v], device=dev, dtype=torch.float32)

        def stage(t, dt, Zy_base, Zy_at, Zw_base, Zw_at, noise, ay_base, ay_at, aw_base, aw_at, want_params):
            """One Euler-like stage evaluated at (Zy_at, Zw_at): returns Z_base - dt f + sigma noise and a_base + dt J^T a_at."""
            tt, tneg, tpos = f32t(t), f32t(-dt), f32t(dt)
            wtraj2 = torch.empty(1, 2, P, device=dev, dtype=torch.float32)
            z1 = torch.empty(cfg.hidden[0], device=dev, dtype=torch.float32)
            st = L.bsde_wsde_fwd(C.byref(d), C.byref(fp), _lib.ptr(Zw_at), _lib.ptr(tt), _lib.ptr(tneg), _lib.ptr(noise), None,
                                 _lib.ptr(wtraj2), _lib.ptr(z1), _lib.ptr(kl_scratch), C.c_void_p(ws.data_ptr()), ws.numel(),
                                 _lib.stream_ptr())
            _lib.check(st, "bsde_wsde_fwd")
            Zw_new = wtraj2[0, 1] if Zw_base is Zw_at else wtraj2[0, 1] - Zw_at + Zw_base
            chain_forward(desc, t, Zw_at, 2 * P, Zy_at, acts)
            Zy_new = torch.empty_like(Zy_base)
            reinterp_axpy(Zy_new, chw_in, Zy_base, -dt, acts[ooff:], chw_out)
            # adjoint rows: a_base + dt J(Z_at)^T a_at
            reinterp_axpy(gf, chw_out, None, dt, ay_at, chw_in)
            gyp, gwy = torch.empty_like(ay_at), torch.empty(1, 1, P, device=dev, dtype=torch.float32)
            chain_backward(desc, t, Zw_at, 2 * P, Zy_at, acts, gf, gyp, False, gwy, P)
            ay_new = ay_base + gyp
            gl = [(torch.empty_like(W), torch.empty_like(b)) for W, b in klayers]
            gfp = _plain_fw_struct(gl, P, cfg.hidden, cfg.w_act)
            r = torch.empty(P, device=dev, dtype=torch.float32)
            st = L.bsde_wsde_bwd(C.byref(d), C.byref(fp), _lib.ptr(tt), _lib.ptr(tpos), _lib.ptr(wtraj2), _lib.ptr(z1), _lib.ptr(gwy),
                                 _lib.ptr(a_l), _lib.ptr(aw_at), _lib.ptr(r), C.byref(gfp), C.c_void_p(ws.data_ptr()), ws.numel(),
                                 _lib.stream_ptr())
            _lib.check(st, "bsde_wsde_bwd")
            aw_new = r if aw_base is aw_at else r - aw_at + aw_base
            # this stage's dt (df/dparams)^T a_at, weighted by `want_params` in the parameter adjoint (True = 1, False = unused)
            gp = [t * float(want_params) for pair in gl for t in pair] if want_params else None
            return Zy_new, Zw_new.contiguous(), ay_new, aw_new, gp

        def rev_step(tau, h, dWr, state):
            """One reverse step of the scheme from forward time tau to tau - h: state = (Zy, Zw, a_y, a_w, *a_params) -> new state."""
            Zy, Zw, a_y, a_w = state[:4]
            gacc = list(state[4:])
            if cfg.method == "heun":   # Z' = Z - h f(tau, Z) + g dW~;  Z1 = (Z + Z')/2 - h/2 f(tau - h, Z') + g dW~/2;  a likewise
                Zy_p, Zw_p, ay_p, aw_p, g1 = stage(tau, h, Zy, Zy, Zw, Zw, dWr, a_y, a_y, a_w, a_w, 0.5)
                Zy, Zw, a_y, a_w, g2 = stage(float(np.float32(tau) - np.float32(h)), 0.5 * h, 0.5 * (Zy + Zy_p), Zy_p, 0.5 * (Zw + Zw_p),
                                             Zw_p, 0.5 * dWr, 0.5 * (a_y + ay_p), ay_p, 0.5 * (a_w + aw_p), aw_p, 1.0)
                gacc = [a + b + c for a, b, c in zip(gacc, g1, g2)]
            elif mid:
                Zy_p, Zw_p, ay_p, aw_p, _ = stage(tau, 0.5 * h, Zy, Zy, Zw, Zw, 0.5 * dWr, a_y, a_y, a_w, a_w, False)
                Zy, Zw, a_y, a_w, g2 = stage(float(np.float32(tau) - np.float32(0.5) * np.float32(h)), h, Zy, Zy_p, Zw, Zw_p, dWr,
                                             a_y, ay_p, a_w, aw_p, True)
                gacc = [a + b for a, b in zip(gacc, g2)]
            else:
                Zy, Zw, a_y, a_w, g1 = stage(tau, h, Zy, Zy, Zw, Zw, dWr, a_y, a_y, a_w, a_w, True)
                gacc = [a + b for a, b in zip(gacc, g1)]
            return (Zy, Zw, a_y, a_w, *gacc)

        t0 = float(cfg.tk[0])
        state = (Zy, Zw, a_y, a_w, *gacc)
        if ctx.rev is None:
            for k in range(n - 1, -1, -1):
                tau = float(np.float32(t0) + np.float32(k + 1) * np.float32(h))
                state = rev_step(tau, h, (-I[0, k]).contiguous(), state)
        else:
            # adjoint_adaptive=True: the step-doubling controller on the reverse solve (reverse time s = -t runs from -t1 to -t0;
            # the error norm covers [Z_y, Z_w, a_y, a_w, a_params]), increments W(tau - h) - W(tau) from the queryable tree
            tree, rtol, atol, dt0, t1 = ctx.rev

            def one_step(sa, hh, st):
                tau, hh = float(np.float32(-sa)), float(hh)
                dWr = (tree(float(np.float32(tau) - np.float32(hh))) - tree(tau)).contiguous()
                return rev_step(tau, hh, dWr, st)

            ts_r, hs_r, state, _ = sdenet_adaptive_grid(one_step, -t1, -t0, dt0, rtol, atol, state)
            global LAST_REVERSE_GRID
            LAST_REVERSE_GRID = (ts_r, hs_r)
        Zy, Zw, a_y, a_w = state[:4]
        gacc = list(state[4:])
        global LAST_RECONSTRUCTION
        LAST_RECONSTRUCTION = (Zy, Zw)
        return (None, None, None, None, None, None, a_y, a_w, *gacc)


LAST_RECONSTRUCTION = None
LAST_REVERSE_GRID = None


def sdenet_solve_adjoint(cfg, x_nhwc, w0, linears, increments=None, seed=0, stream_id=0, final=None, rev=None):
    """As `sdenet_solve`, with the O(1)-memory reverse pass of `sdeint_adjoint`."""
    kl = []
    for W, b in linears:
        kl += [W.t().contiguous(), b]
    return SDENetAdjointSolve.apply(cfg, increments, seed, stream_id, final, rev, x_nhwc, w0, *kl)


def sdenet_solve(cfg, x_nhwc, w0, linears, increments=None, seed=0, stream_id=0):
    """Differentiable solve.  x_nhwc [S,B,H,W,C]; linears = [(weight [out,in], bias)] in torch nn.Linear layout."""
    kl = []
    for W, b in linears:
        kl += [W.t().contiguous(), b]      # kernel layout: W1 [P,d1], Wm [din,dout], W4 [dL,P]
    cfg.track = torch.is_grad_enabled()
    return SDENetSolve.apply(cfg, increments, seed, stream_id, x_nhwc, w0, *kl)


# ------------------------------------------------------------------------------------------------
# adaptive stepping (`adaptive=True`, models.py:194-197 -> torchsde's step-doubling controller)
# ------------------------------------------------------------------------------------------------
def update_step_size(error_estimate, prev_step_size, safety=0.9, facmin=0.2, facmax=1.4, prev_error_ratio=None):
    """torchsde/_core/adaptive_stepping.py (restated from the published controller: PI step-size control, Burrage & Burrage 2004 /
    Ilie, Jackson & Enright 2015): returns (new_step_size, prev_error_ratio)."""
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


def compute_error(full, halves, rtol, atol, eps=1e-7):
    """RMS over the WHOLE augmented state [y, w, logqp] of (one full step - two half steps) / (rtol max(|.|, |.|) + atol + eps)."""
    num, cnt = 0.0, 0
    for a, b in zip(full, halves):
        tol = rtol * torch.maximum(a.abs(), b.abs()) + atol
        num = num + (((a - b) / (tol + eps)) ** 2).sum()
        cnt += a.numel()
    return max(float(torch.sqrt(num / cnt)), eps)


def sdenet_adaptive_grid(one_step, t0, t1, dt, rtol, atol, state, dt_min=1e-5, max_trials=100000):
    """torchsde's adaptive loop for ts = [t0, t1] (BaseSDESolver.integrate, adaptive branch): every trial takes one full step and two
    half steps from the current state, accepts the half-step result when the error estimate is <= 1 (or the step hit dt_min) and
    updates the step with `update_step_size` either way.  `one_step(ta, h, state) -> state`.  Returns the accepted (t, h) half steps,
    the final state and the number of single steps taken (accepted or not)."""
    f32 = np.float32
    curr_t, step, prev_ratio = f32(t0), f32(dt), None
    ts, hs, nsingle = [], [], 0
    for _ in range(max_trials):
        if not curr_t < f32(t1):
            return ts, hs, state, nsingle
        next_t = min(f32(curr_t + step), f32(t1))
        mid_t = f32(f32(0.5) * f32(curr_t + next_t))
        full = one_step(curr_t, f32(next_t - curr_t), state)
        mid = one_step(curr_t, f32(mid_t - curr_t), state)
        nxt = one_step(mid_t, f32(next_t - mid_t), mid)
        nsingle += 3
        err = compute_error(full, nxt, rtol, atol)
        new_step, prev_ratio = update_step_size(err, float(step), prev_error_ratio=prev_ratio)
        step = f32(new_step)
        if step < dt_min:
            step, prev_ratio = f32(dt_min), None
        if err <= 1 or step <= dt_min:
            ts += [curr_t, mid_t]
            hs += [f32(mid_t - curr_t), f32(next_t - mid_t)]
            curr_t, state = next_t, nxt
    raise RuntimeError(f"adaptive solve did not reach t1 = {t1} in {max_trials} trials")


# ------------------------------------------------------------------------------------------------
# SDENet: models.py:105-204
# ------------------------------------------------------------------------------------------------
class SDENet(nn.Module):
    def __init__(self, input_size=(3, 32, 32), blocks=(2, 2, 2), weight_network_sizes=(1, 64, 1), num_classes=10,
                 activation="softplus", verbose=False, inhomogeneous=True, sigma=0.1, hidden_width=128, aug_dim=0,
                 math="tc", remat=False):
        super().__init__()
        self.input_size = tuple(input_size)
        self.aug_input_size = (aug_dim + input_size[0], *input_size[1:])
        self.aug_zeros_size = (aug_dim, *input_size[1:])
        self.register_buffer("aug_zeros", torch.zeros(size=(1, *self.aug_zeros_size)))
        self.y_net, self.output_size = make_y_net(input_size=input_size, blocks=blocks, activation=activation, verbose=verbose,
                                                  hidden_width=hidden_width, aug_dim=aug_dim)
        flat, self.unravel_params = ravel_pytree(self.y_net.make_initial_params())
        self.flat_initial_params = nn.Parameter(flat, requires_grad=True)
        self.params_size = flat.numel()
        self.w_net = make_w_net(in_features=self.params_size, hidden_sizes=weight_network_sizes, activation="tanh",
                                inhomogeneous=inhomogeneous)
        self.projection = nn.Sequential(nn.Flatten(), nn.Linear(int(np.prod(self.output_size)), 1024), nn.ReLU(inplace=True),
                                        nn.Linear(1024, num_classes))
        self.register_buffer("ts", torch.tensor([0., 1.]))
        self.sigma = sigma
        self.nfe = 0
        self.math, self.remat = math, remat
        self.seed, self._calls = 0, 0
        self._cfg = {}

    def make_initial_params(self):
        return self.y_net.make_initial_params()

    def _config(self, B, nsteps, h, method):
        key = (B, nsteps, h, method, self.math, self.remat)
        if key not in self._cfg:
            self._cfg[key] = SolveConfig(self.y_net, 1, B, nsteps, h, method, self.sigma, self.w_net.hidden_sizes,
                                         self.w_net.activation, self.math, self.remat, t0=float(self.ts[0]))
        return self._cfg[key]

    def forward(self, y, adjoint=False, dt=0.02, adaptive=False, adjoint_adaptive=False, method="midpoint", rtol=1e-4, atol=1e-3,
                increments=None):
        """-> (logits, logqp).  `increments` [nsteps, P]: W(t_{k+1}) - W(t_k) to replay (else in-kernel Philox keyed by
        (self.seed, call counter), the stand-in for torchsde.BrownianInterval, models.py:190-193)."""
        if adjoint_adaptive and not adjoint:
            raise ValueError("adjoint_adaptive=True needs adjoint=True")
        if not y.is_cuda:
            raise _lib.BsdeError("bayesde_b200 kernels need CUDA tensors (there is no CPU fallback)")
        t0, t1 = float(self.ts[0]), float(self.ts[-1])
        nsteps = int(round((t1 - t0) / dt))
        if self.aug_zeros.numel() > 0:  # add zero channels (models.py:185-187)
            y = torch.cat((y, self.aug_zeros.expand(y.shape[0], *self.aug_zeros_size)), dim=1)
        x = y.float().permute(0, 2, 3, 1).contiguous()[None]               # NHWC storage, S = 1
        lin = [(m.weight, m.bias) for m in self.w_net.linears()]
        self._calls += 1
        if adaptive or adjoint_adaptive:
            return self._forward_adaptive(x, lin, t0, t1, dt, method, rtol, atol, adjoint, nsteps, adaptive, adjoint_adaptive)
        cfg = self._config(y.shape[0], nsteps, dt, method)
        I = None if increments is None else increments[None]
        solve = sdenet_solve_adjoint if adjoint else sdenet_solve  # models.py:194: sdeint_adjoint / sdeint
        y1, lq, _ = solve(cfg, x, self.flat_initial_params, lin, I, seed=self.seed, stream_id=self._calls)
        # f and g are each evaluated once per stage (models.py:161,172): 2 stages for midpoint; euler_heun evaluates g twice
        self.nfe = nsteps * {"midpoint": 4, "heun": 4, "euler_heun": 3, "milstein": 2}[method]
        y1 = y1[0].permute(0, 3, 1, 2)                                       # back to (B, C, H, W) order for Flatten
        logits = self.projection(y1.reshape(y1.shape[0], -1))
        logqp = .5 * lq[0]
        return logits, logqp

    def _forward_adaptive(self, x, lin, t0, t1, dt, method, rtol, atol, adjoint=False, nsteps=0, adaptive=True, adjoint_adaptive=False):
        """`adaptive=True`: the accepted step grid is found by torchsde's step-doubling loop on a queryable Philox Brownian tree over
        the P weight rows (the stand-in for BrownianInterval: any interval's increment is W(tb) - W(ta)), without autograd; the
        solve is then replayed on that grid through `sdenet_solve` so that `loss.backward()` differentiates the accepted steps."""
        from .sdeint import BrownianTree
        P, B = self.params_size, x.shape[1]
        tree = BrownianTree(t0, P, t1, (self.seed, self._calls), device=x.device)
        mk = lambda ts, hs: SolveConfig(self.y_net, 1, B, len(ts), 0.0, method, self.sigma, self.w_net.hidden_sizes,
                                        self.w_net.activation, self.math, True, grid=(ts, hs))
        inc = lambda ts, hs: torch.stack([tree(float(np.float32(t + h))) - tree(float(t)) for t, h in zip(ts, hs)])[None]

        def one_step(ta, h, state):
            yy, ww, lq = state
            cfg1 = mk([ta], [h])
            y1, kl, wtraj = sdenet_solve(cfg1, yy, ww, lin, inc([ta], [h]))
            return y1, wtraj[0, cfg1.nit].contiguous(), lq + kl[0]

        final, nsingle = None, nsteps
        if adaptive:
            with torch.no_grad():
                state0 = (x, self.flat_initial_params.detach().float().contiguous(), torch.zeros((), device=x.device))
                ts, hs, final, nsingle = sdenet_adaptive_grid(one_step, t0, t1, dt, rtol, atol, state0)
            self.adaptive_grid = (ts, hs)
            final = (final[0], final[2].reshape(1), final[1])
        if adjoint:
            # sdeint_adjoint(adaptive=True, adjoint_adaptive=False): adaptive forward solve, reverse solve with fixed steps `dt` on
            # the same path (the tree's increments over the uniform grid), started from the adaptive solve's final state
            cfg = self._config(x.shape[1], nsteps, dt, method)
            tu = [np.float32(t0) + np.float32(k) * np.float32(dt) for k in range(nsteps)]
            rev = (tree, rtol, atol, dt, t1) if adjoint_adaptive else None
            y1, lq, _ = sdenet_solve_adjoint(cfg, x, self.flat_initial_params, lin, inc(tu, [np.float32(dt)] * nsteps),
                                             final=final, rev=rev)
        else:
            cfg = mk(ts, hs)
            y1, lq, _ = sdenet_solve(cfg, x, self.flat_initial_params, lin, inc(ts, hs))
        self.nfe = nsingle * {"midpoint": 4, "heun": 4, "euler_heun": 3, "milstein": 2}[method]
        y1 = y1[0].permute(0, 3, 1, 2)
        logits = self.projection(y1.reshape(y1.shape[0], -1))
        return logits, .5 * lq[0]

    def zero_grad(self, set_to_none=True) -> None:
        for p in self.parameters():
            p.grad = None

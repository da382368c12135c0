"""Host-side mirror of brax/_impl/arch.py (+ the stax pieces the example script composes):
`Layer`, `MLP` (the weight-drift net f_w), `build_fx` (the per-block conv stack f_x), `Augment`,
`SqueezeDownsample`, and a minimal `serial / Flatten / Dense / LogSoftmax`.

These objects only carry *structure* (shapes, parameter layout, init); the arithmetic of f_w and
f_x on the hot path runs in libbsde.so (see brax.py / sdeint.py).  `rng` arguments are integer
seeds (the reference passes JAX PRNG keys).
"""
import collections
import math

import torch
import torch.nn.functional as F

from . import _lib


class Layer(collections.namedtuple("Layer", "init, apply")):
    """(init, apply) pair -- brax/_impl/arch.py:10-11."""


def _gen(rng, device="cpu"):
    g = torch.Generator(device=device)
    g.manual_seed(int(rng) & 0x7FFFFFFFFFFFFFFF)
    return g


def split(rng, n):
    """Deterministic stand-in for jax.random.split on integer seeds."""
    rng = int(rng) & 0xFFFFFFFFFFFFFFFF
    out = []
    for i in range(n):
        z = (rng + (i + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF  # splitmix64
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        out.append((z ^ (z >> 31)) & 0x7FFFFFFFFFFFFFFF)
    return out


ACTFNS = ("softplus", "tanh", "elu", "rbf", "swish")  # arch.py:31-32


def _act(name, x, beta=None):
    if name == "swish":  # arch.py:14-29: trainable beta per unit / channel
        return x * torch.sigmoid(x * F.softplus(beta))
    if name == "softplus":
        return F.softplus(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "elu":
        return F.elu(x)
    if name == "rbf":
        return torch.exp(-x ** 2)
    raise NotImplementedError(f"activation {name!r} is not available in the fused B200 path")


# ------------------------------------------------------------------------------------------------
# f_w : arch.py:34-74
# ------------------------------------------------------------------------------------------------
class FwLayer(Layer):
    """Layer whose structure the fused K1 kernel understands (carries `.spec`)."""
    spec = None


def MLP(hidden_dims=(1, 64, 1), actfn="softplus", xt=False, ou_dw=False, p_scale=-1.0, nonzero_w=-1.0,
        nonzero_b=-1.0, device=None):
    """P -> hidden_dims -> P stack of ConcatSquashLinear (xt=True).  arch.py:34-74.

    Parameter pytree (mirrors stax.serial): [(W,b,w_t,w_tb,b_t), (), ..., (W,b,w_t,w_tb,b_t)], wrapped
    as ((), mlp_params) when ou_dw (arch.Additive; quirk Q3: the OU term is *not* added to the output).
    nonzero_w / nonzero_b: std of the last layer's init (the reference's he_normal/normal with those
    flags are undefined names, arch.py:54-57; -1 keeps the zero init of arch.py:51)."""
    hidden_dims = tuple(int(d) for d in hidden_dims)
    if not xt:
        raise NotImplementedError("MLP(xt=False): the fused path implements the time-dependent f_w only")
    if p_scale != -1.0:
        raise NotImplementedError("p_scale: arch.Exp is undefined in the reference (arch.py:65)")
    if actfn not in ACTFNS:
        raise KeyError(actfn)  # ACTFNS[actfn], arch.py:37
    if len(hidden_dims) < 1 or len(hidden_dims) > _lib.FW_MAX_MID + 1:
        raise ValueError(f"len(hidden_dims) must be in 1..{_lib.FW_MAX_MID + 1}")
    if hidden_dims[0] > _lib.FW_MAX_EDGE or hidden_dims[-1] > _lib.FW_MAX_EDGE or max(hidden_dims) > _lib.FW_MAX_HID:
        raise ValueError(f"hidden_dims {hidden_dims}: first/last width <= {_lib.FW_MAX_EDGE}, others <= {_lib.FW_MAX_HID}")

    def init_fun(rng, input_shape):
        P = int(input_shape[-1])
        g = _gen(rng)
        dims = (P,) + hidden_dims + (P,)
        params = []
        for i in range(len(dims) - 1):
            din, dout = dims[i], dims[i + 1]
            last = i == len(dims) - 2
            if last:
                sw = 0.0 if nonzero_w == -1.0 else float(nonzero_w)
                sb = 0.0 if nonzero_b == -1.0 else float(nonzero_b)
            else:
                sw, sb = math.sqrt(2.0 / din), 1e-2  # he_normal / normal(1e-2), diffeq_layers.py:73
            mk = lambda *s, sd: (torch.randn(*s, generator=g) * sd).to(device)
            params.append((mk(din, dout, sd=sw), mk(dout, sd=sb), mk(dout, sd=sb), mk(dout, sd=sb), mk(dout, sd=sb)))
            if not last:  # the activation's parameters: () or, for the trainable swish, (beta,) = 0.5 (arch.py:14-22,41-42)
                params.append((torch.full((dout,), 0.5).to(device),) if actfn == "swish" else ())
        return tuple(input_shape), (((), params) if ou_dw else params)

    def apply_fun(params, inputs, **kwargs):
        x, t = inputs
        layers, betas = fw_layers(params), fw_betas(params)
        h = x
        for i, (W, b, w_t, w_tb, b_t) in enumerate(layers):
            h = (h @ W + b) * torch.sigmoid(w_t * t + w_tb) + b_t * t  # diffeq_layers.py:84-95
            if i < len(layers) - 1:
                h = _act(actfn, h, betas[i] if betas else None)
        return (h, 2 * t) if ou_dw else (h, t)  # arch.py:100-103 returns (t2, out2+out1) = (mlp_out, 2t)

    layer = FwLayer(init_fun, apply_fun)
    layer.spec = dict(hidden_dims=hidden_dims, actfn=actfn, ou_dw=ou_dw)
    return layer


def fw_layers(fw_params):
    """The list of (W,b,w_t,w_tb,b_t) tuples inside an MLP parameter pytree."""
    p = fw_params
    if isinstance(p, (tuple, list)) and len(p) == 2 and isinstance(p[0], (tuple, list)) and len(p[0]) == 0:
        p = p[1]
    return [q for q in p if len(q) == 5]


def fw_betas(fw_params):
    """The swish beta vectors (one per hidden layer) inside an MLP parameter pytree; [] for the other activations."""
    p = fw_params
    if isinstance(p, (tuple, list)) and len(p) == 2 and isinstance(p[0], (tuple, list)) and len(p[0]) == 0:
        p = p[1]
    return [q[0] for q in p if len(q) == 1]


# ------------------------------------------------------------------------------------------------
# f_x : arch.py:171-210, weight layout = ravel_pytree order (brax.py:41)
# ------------------------------------------------------------------------------------------------
def same_pads(in_size, k, stride):
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return total // 2, total - total // 2


class FxSpec:
    """Conv stack of one SDEBNN block for a fixed NHWC input shape."""

    def __init__(self, block_type, input_shape, fx_dim, fx_actfn):
        if fx_actfn not in ACTFNS:
            raise KeyError(fx_actfn)  # ACTFNS[fx_actfn], arch.py:172
        _, H, W, C = input_shape
        self.block_type, self.fx_dim, self.actfn = block_type, fx_dim, fx_actfn
        self.H, self.W, self.C = int(H), int(W), int(C)
        if block_type == 0:
            convs = [(C, fx_dim, 3, 1, False), (fx_dim, fx_dim, 3, 2, False), (fx_dim, fx_dim, 3, 2, True),
                     (fx_dim, C, 3, 1, False)]
        elif block_type == 1:
            # arch.py:193 writes ConcatConv2D(fx_dim, (1, 1), padding="SAME"), but the tuple binds to the unused `W_init`
            # positional (diffeq_layers.py:124: ConcatConv2D(out_dim, W_init, b_init, kernel=3, ...)), so the middle
            # convolution of block type 1 is 3x3 as shipped (same mechanism as quirk Q12); kept for parity
            convs = [(C, fx_dim, 3, 1, False), (fx_dim, fx_dim, 3, 1, False), (fx_dim, C, 3, 1, False)]
        elif block_type == 2:
            convs = [(C, fx_dim, 3, 1, False), (fx_dim, C, 3, 1, False)]
        else:
            raise ValueError(f"Invalid block_type {block_type}")  # arch.py:207
        self.layers = []
        off, h, w = 0, self.H, self.W
        for li, (ci, co, k, s, tr) in enumerate(convs):
            if not tr:
                lo_h, _ = same_pads(h, k, s)
                lo_w, _ = same_pads(w, k, s)
                assert lo_h == lo_w
                ho, wo = -(-h // s), -(-w // s)
                geo = dict(sn=s, kn=1, off=-lo_h, sd=1)
            else:  # lax.conv_transpose SAME: lhs dilation s, pads (ceil((k+s-2)/2), ...), kernel not flipped
                pad_len = k + s - 2
                pad_a = k - 1 if s > k - 1 else int(math.ceil(pad_len / 2))
                ho, wo = h * s, w * s
                geo = dict(sn=1, kn=1, off=-pad_a, sd=s)
            nw = k * k * (ci + 1) * co
            self.layers.append(dict(cin=ci, cout=co, k=k, hin=h, win=w, hout=ho, wout=wo, w_off=off,
                                    b_off=off + nw, fan_in=k * k * (ci + 1), fan_out=k * k * co, **geo))
            off += nw + co
            # trainable swish: the (beta,) leaf of the activation follows the conv's (W, b) in ravel_pytree order (arch.py:172-188)
            self.layers[-1]["beta_off"] = -1
            if fx_actfn == "swish" and li < len(convs) - 1:
                self.layers[-1]["beta_off"] = off
                off += fx_dim
            h, w = ho, wo
        if (h, w) != (self.H, self.W):
            raise AssertionError(f"fx needs to have the same input and output shapes but got "
                                 f"{(self.H, self.W)} and {(h, w)}")  # brax.py:40
        self.P = off

    def conv_layer_structs(self):
        out = []
        for L in self.layers:
            c = _lib.ConvLayer()
            c.cin, c.cout, c.ksize = L["cin"], L["cout"], L["k"]
            c.hin, c.win, c.hout, c.wout = L["hin"], L["win"], L["hout"], L["wout"]
            c.sn, c.kn, c.off, c.sd = L["sn"], L["kn"], L["off"], L["sd"]
            c.has_time, c.time_first = 1, 0
            c.w_off, c.b_off = L["w_off"], L["b_off"]
            c.w_tap_stride, c.w_k_stride, c.w_n_stride = (L["cin"] + 1) * L["cout"], L["cout"], 1  # HWIO
            out.append(c)
        return out

    def init(self, rng, device=None):
        """stax.GeneralConv defaults: W glorot-normal, b normal(1e-6); flat in ravel order."""
        g = _gen(rng)
        w = torch.zeros(self.P)
        for L in self.layers:
            nw = L["b_off"] - L["w_off"]
            std = math.sqrt(2.0 / (L["fan_in"] + L["fan_out"]))
            w[L["w_off"]:L["b_off"]] = torch.randn(nw, generator=g) * std
            w[L["b_off"]:L["b_off"] + L["cout"]] = torch.randn(L["cout"], generator=g) * 1e-6
            if L["beta_off"] >= 0:
                w[L["beta_off"]:L["beta_off"] + self.fx_dim] = 0.5   # Swish_(out_dim, beta_init=0.5), arch.py:14
        return w.to(device)


def build_fx(block_type, input_size, fx_dim, fx_actfn):
    """arch.py:171-210.  Returns a Layer whose init gives the flat weight vector's pytree stand-in
    (a single flat tensor) and carries `.spec` for the fused kernels."""
    spec = FxSpec(block_type, input_size, fx_dim, fx_actfn)

    def init_fun(rng, input_shape):
        return tuple(input_shape), spec.init(rng)

    def apply_fun(params, inputs, **kwargs):
        raise NotImplementedError("f_x is evaluated inside libbsde.so; use SDEBNN / sdeint_ito_fixed_grid")

    layer = FwLayer(init_fun, apply_fun)
    layer.spec = spec
    return layer


# ------------------------------------------------------------------------------------------------
# parameter-free layers: arch.py:108-168
# ------------------------------------------------------------------------------------------------
def Augment(pad_zeros):
    def init_fun(rng, input_shape):
        return tuple(input_shape[:-1]) + (pad_zeros + input_shape[-1],), ()

    def apply_fun(params, inputs, **kwargs):
        if pad_zeros == 0:
            return inputs
        return torch.cat([inputs, inputs.new_zeros(inputs.shape[:-1] + (pad_zeros,))], dim=-1)

    return Layer(init_fun, apply_fun)


def _squeeze(x, r=2):
    """arch.py:131-153: NHWC -> NCHW -> space-to-depth -> NHWC (leading dims are batch)."""
    lead = x.shape[:-3]
    h, w, c = x.shape[-3:]
    v = x.reshape(-1, h, w, c).permute(0, 3, 1, 2)
    b = v.shape[0]
    v = v.reshape(b, c, h // r, r, w // r, r).permute(0, 1, 3, 5, 2, 4).reshape(b, c * r * r, h // r, w // r)
    return v.permute(0, 2, 3, 1).reshape(*lead, h // r, w // r, c * r * r).contiguous()


def SqueezeDownsample(pad_zeros):
    def init_fun(rng, input_shape):
        n, h, w, c = input_shape
        assert h % 2 == 0 and w % 2 == 0
        return (n, h // 2, w // 2, c * 4), ()

    def apply_fun(params, inputs, **kwargs):
        return _squeeze(inputs)

    return Layer(init_fun, apply_fun)


# ------------------------------------------------------------------------------------------------
# the stax pieces of the classifier head (examples/jax/sdebnn_classification.py:211) [3P semantics]
# ------------------------------------------------------------------------------------------------
def Dense(out_dim):
    def init_fun(rng, input_shape):
        g = _gen(rng)
        din = int(input_shape[-1])
        W = torch.randn(din, out_dim, generator=g) * math.sqrt(2.0 / (din + out_dim))  # glorot_normal
        b = torch.randn(out_dim, generator=g) * 1e-2
        return tuple(input_shape[:-1]) + (out_dim,), (W, b)

    def apply_fun(params, inputs, **kwargs):
        W, b = params
        if W.dim() == 3:  # one weight draw per MC sample: inputs (S,B,F)
            return torch.baddbmm(b.unsqueeze(1), inputs, W)
        return inputs @ W + b

    return init_fun, apply_fun


def _flatten_init(rng, input_shape):
    return (input_shape[0], int(math.prod(input_shape[1:]))), ()


Flatten = (_flatten_init, lambda params, inputs, **kw: inputs.reshape(*inputs.shape[:-3], -1))
LogSoftmax = (lambda rng, s: (s, ()), lambda params, inputs, **kw: F.log_softmax(inputs, dim=-1))


def serial(*layers):
    inits, applies = zip(*layers)

    def init_fun(rng, input_shape):
        params = []
        for r, init in zip(split(rng, len(inits)), inits):
            input_shape, p = init(r, input_shape)
            params.append(p)
        return input_shape, params

    def apply_fun(params, inputs, **kwargs):
        for f, p in zip(applies, params):
            inputs = f(p, inputs, **kwargs)
        return inputs

    return init_fun, apply_fun

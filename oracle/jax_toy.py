"""TEST INFRASTRUCTURE (CPU oracle, fp64 torch) -- config 1, the JAX toy: `SDELayer` and its augmented dynamics.

Line-by-line restatement of
    brax/_impl/layers.py:19-42    ConcatSquashLineartx      y = sigmoid(w_t t + w_tb) (x W + b) + b_t t, inputs (t, x)
    brax/_impl/layers.py:54-118   augmented_post_drift      state [w (P), x (B, D), kl ()]  (ravel order: w first)
    brax/_impl/layers.py:120-143  augmented_prior_drift
    brax/_impl/layers.py:145-157  augmented_diffusion
    brax/_impl/layers.py:164-248  SDELayer.apply_fun        (posterior and prior solves on linspace(0, 1, 20), :242-245)
    brax/utils/sdeint.py:8-60     _sdeint / sdeint          state += h f(t0, state, bm[:bm_dim]) + sqrt(h) bm g(t0, state)
    examples/jax/sdebnn_toy1d.py:32-95  mlp(): driftx = CSL(H) act CSL(H) act CSL(D); driftw = Additive(prior, CSL(H) act CSL(H)
                                        act CSL(P)) when diff_drift; prior = Affine(-1, 0) ("ou") or Const(0); diffusion = Const(sigma)
    brax/_impl/layers.py:310-326  trainable Swish_: x sigmoid(x softplus(beta)), beta = 0.5 per unit, one (beta,) leaf per use

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  PARITY UNPINNED for this file: the reference
is JAX 0.1.77 code that cannot run in this image and ships no golden vectors; the restatement is anchored on the reference's own
call sites and formulas (cited per function).

Parameter containers:
    driftx weights live in the STATE as the flat vector w = ravel_pytree(W0): for each ConcatSquashLineartx (W [din, dout], b, w_t,
    w_tb, b_t) then, for swish, the (beta,) leaf of the activation that follows;
    driftw (posterior drift of w) parameters: a list like oracle.jax_path.init_fw's -- 5-tuples (W, b, w_t, w_tb, b_t), with a
    sixth entry beta on every layer but the last when the activation is the trainable swish.
Random numbers: `bm[k]` (D_total,) unit normals per step are an INPUT (the reference draws random.normal(rng_k, state.shape),
sdeint.py:16; eps = bm[:bm_dim] goes to the drift, quirk Q5).
"""
import math

import torch

from .jax_path import act, concat_squash_linear


def driftx_layout(D, H, actfn):
    """ravel_pytree order of the driftx parameter tree for x of width D (sdebnn_toy1d.py:48-56).  Returns (entries, P)."""
    dims = [(D, H), (H, H), (H, D)]
    entries, off = [], 0
    for i, (din, dout) in enumerate(dims):
        e = dict(din=din, dout=dout, W=off, b=off + din * dout, wt=off + din * dout + dout, wtb=off + din * dout + 2 * dout,
                 bt=off + din * dout + 3 * dout, beta=-1)
        off += din * dout + 4 * dout
        if i < 2 and actfn == "swish":
            e["beta"] = off
            off += dout
        entries.append(e)
    return entries, off


def driftx_apply(flat_w, t, x, layout, actfn):
    """driftx_apply_fn(param_unravel_fn(flat_w), (t, x)) (layers.py:62): x (B, D) -> (B, D)."""
    h = x
    for i, e in enumerate(layout):
        din, dout = e["din"], e["dout"]
        p = (flat_w[e["W"]:e["W"] + din * dout].reshape(din, dout), flat_w[e["b"]:e["b"] + dout], flat_w[e["wt"]:e["wt"] + dout],
             flat_w[e["wtb"]:e["wtb"] + dout], flat_w[e["bt"]:e["bt"] + dout])
        h = concat_squash_linear(p, h, t)
        if i < len(layout) - 1:
            h = act(actfn, h, flat_w[e["beta"]:e["beta"] + dout] if e["beta"] >= 0 else None)
    return h


def driftw_mlp(layers, w, t, actfn):
    h = w
    for i, p in enumerate(layers):
        h = concat_squash_linear(p, h, t)
        if i < len(layers) - 1:
            h = act(actfn, h, p[5] if len(p) > 5 else None)
    return h


class ToySetting(dict):
    """`settings` of sdebnn_toy1d.py:186-198 (defaults of the argparse block :140-176 for config 1 given by the constructor)."""

    def __init__(self, driftw=True, diffw=True, diff_const=0.1, prior_driftw="ou", stop_grad=True, diff_drift=True, aug_dim=2):
        super().__init__(driftw=driftw, diffw=diffw, diff_const=diff_const, prior_driftw=prior_driftw, stop_grad=stop_grad,
                         diff_drift=diff_drift, aug_dim=aug_dim)


class SDELayerOracle:
    """SDELayer(driftx, prior_driftw, post_driftw, diffusion, x_dim, w_dim, setting) for x of shape (B, D) (layers.py:164-248);
    point estimate of W0 (priorw0_sigma < 0, the toy's default)."""

    def __init__(self, B, D, H, actfn="swish", setting=None):
        self.B, self.D, self.H, self.actfn = B, D, H, actfn
        self.setting = setting or ToySetting()
        self.layout, self.P = driftx_layout(D, H, actfn)

    # ---- pieces of mlp() -------------------------------------------------------------------------------------------------
    def prior_drift(self, w):  # DiffEqWrappertx(Affine(-1, 0.)) / Const(0.)  (sdebnn_toy1d.py:45-47)
        return -w if self.setting["prior_driftw"] == "ou" else torch.zeros_like(w)

    def post_driftw(self, W_driftwt, t, w):  # sdebnn_toy1d.py:66-83
        if not self.setting["driftw"]:
            return torch.zeros_like(w)  # DiffEqWrappertx(Const(0.))
        m = driftw_mlp(W_driftwt, w, t, self.actfn)
        return m + self.prior_drift(w) if self.setting["diff_drift"] else m  # Additive(prior, mlp), layers.py:272-292

    def diffusion(self, w):  # Const(diff_const) or Const(0.)  (sdebnn_toy1d.py:85-88)
        return torch.full_like(w, self.setting["diff_const"] if self.setting["diffw"] else 0.0)

    # ---- layers.py:54-118 ------------------------------------------------------------------------------------------------
    def unravel(self, aug):
        P, n = self.P, self.B * self.D
        return aug[:P], aug[P:P + n].reshape(self.B, self.D), aug[P + n]

    def augmented_post_drift(self, W_driftwt, t, aug, eps):
        st = self.setting
        flat_w, x, _ = self.unravel(aug)
        dxt = driftx_apply(flat_w, t, x, self.layout, self.actfn)              # :62
        dwtkl = self.post_driftw(W_driftwt, t, flat_w)                         # :64
        pw_drift = self.prior_drift(flat_w)                                    # :67
        diffwt = self.diffusion(flat_w)                                        # :68
        u = (dwtkl - pw_drift) * 0
        if st["diffw"]:
            u = (dwtkl - pw_drift) / diffwt                                    # :73-75
        dkl = (u ** 2 / 2).sum()                                               # :76
        dkl_stl = (u * eps).sum()                                              # :77-78
        if st["stop_grad"]:                                                    # :81-112: second term with stop_gradient(params)
            sg = [tuple(q.detach() for q in p) for p in W_driftwt]
            _dwtkl = self.post_driftw(sg, t, flat_w)
            u2 = (_dwtkl - pw_drift) * 0
            if st["diffw"]:
                u2 = (_dwtkl - pw_drift) / diffwt
            dkl_stl = (u2 * eps).sum()
        dkl = dkl + dkl_stl                                                    # :113
        return torch.cat([dwtkl, dxt.reshape(-1), dkl.reshape(1)])             # :116-118

    def augmented_prior_drift(self, t, aug, eps):  # layers.py:120-143
        flat_w, x, _ = self.unravel(aug)
        dxt = driftx_apply(flat_w, t, x, self.layout, self.actfn)
        dwtkl = self.prior_drift(flat_w)
        diffwt = self.diffusion(flat_w)
        u = dwtkl * 0
        if self.setting["diffw"]:
            u = dwtkl / diffwt
        dkl = (u ** 2 / 2).sum() + (u * eps).sum()
        return torch.cat([dwtkl, dxt.reshape(-1), dkl.reshape(1)])

    def augmented_diffusion(self, t, aug):  # layers.py:145-157
        flat_w, x, kl = self.unravel(aug)
        return torch.cat([self.diffusion(flat_w), torch.zeros_like(x).reshape(-1), torch.zeros_like(kl).reshape(1)])

    # ---- brax/utils/sdeint.py:30-60 --------------------------------------------------------------------------------------
    def sdeint(self, f, g, s_init, bm_dim, ts, bm):
        st, out = s_init, []
        for k in range(len(ts) - 1):
            t0, h = ts[k], ts[k + 1] - ts[k]
            f_out = f(t0, st, bm[k, :bm_dim])
            g_out = g(t0, st)
            st = st + h * f_out + torch.sqrt(h) * bm[k] * g_out
            out.append(st)
        return torch.stack(out)

    # ---- layers.py:205-246 -----------------------------------------------------------------------------------------------
    def apply(self, W0_flat, W_driftwt, inputs, bm, nt=20, entire=False):
        """apply_fun(params, inputs, rng): returns (x, w, kl, prior_x, prior_w, prior_kl) at t = 1 -- or, with entire=True, the
        whole trajectories (T-1 rows each).  The reference integrates the posterior and the prior with the SAME rng (:242-245), so
        both solves see the same normals `bm` (nt-1, P + B D + 1)."""
        dt_ = inputs.dtype
        if bm.shape[-1] == self.P:   # normals of the w rows only: the x / kl rows have g = 0, their draws never enter (sdeint.py:19-20)
            bm = torch.cat([bm, bm.new_zeros(bm.shape[0], self.B * self.D + 1)], dim=1)
        ts = torch.linspace(0.0, 1.0, nt, dtype=torch.float32).to(dt_)
        aug0 = torch.cat([W0_flat, inputs.reshape(-1), torch.zeros(1, dtype=dt_)])  # aug_init: [W0, inputs, 0.] (layers.py:159-162)
        post = self.sdeint(lambda t, a, e: self.augmented_post_drift(W_driftwt, t, a, e), self.augmented_diffusion, aug0, self.P, ts, bm)
        prior = self.sdeint(self.augmented_prior_drift, self.augmented_diffusion, aug0, self.P, ts, bm)
        P, n = self.P, self.B * self.D
        split = lambda y: (y[..., P:P + n].reshape(y.shape[:-1] + (self.B, self.D)), y[..., :P], y[..., P + n])
        if entire:
            return split(post) + split(prior)
        return split(post[-1]) + split(prior[-1])


def init_toy(B, D, H, actfn, gen, dtype=torch.float64, driftw_scale=0.1, last_std=0.0, beta_jitter=0.0):
    """SDELayer.init_fun (layers.py:180-203): W0 = driftx init (he_normal W, normal(1e-2) vectors: layers.py:19-29), W_driftwt =
    driftw init scaled by `driftw_scale` (:189-190; the last layer is zero-initialised in mlp() unless stop_grad and diff_drift,
    sdebnn_toy1d.py:69-71 -- `last_std` > 0 gives it a non-trivial value for the tests).  Returns (W0_flat, W_driftwt)."""
    layout, P = driftx_layout(D, H, actfn)
    mk = lambda *s, sd: torch.randn(*s, generator=gen, dtype=torch.float64) * sd
    w0 = torch.zeros(P, dtype=torch.float64)
    for e in layout:
        din, dout = e["din"], e["dout"]
        w0[e["W"]:e["W"] + din * dout] = mk(din * dout, sd=math.sqrt(2.0 / din))
        for key in ("b", "wt", "wtb", "bt"):
            w0[e[key]:e[key] + dout] = mk(dout, sd=1e-2)
        if e["beta"] >= 0:
            w0[e["beta"]:e["beta"] + dout] = 0.5 + mk(dout, sd=beta_jitter)
    dims = [P, H, H, P]
    layers = []
    for i in range(3):
        din, dout = dims[i], dims[i + 1]
        last = i == 2
        sw, so = (last_std, last_std) if last else (math.sqrt(2.0 / din), 1e-2)
        lay = tuple((mk(*s, sd=sd) * driftw_scale).to(dtype) for s, sd in (((din, dout), sw), ((dout,), so), ((dout,), so), ((dout,), so), ((dout,), so)))
        if actfn == "swish" and not last:
            lay = lay + (((0.5 + mk(dout, sd=beta_jitter)) * driftw_scale).to(dtype),)   # tree_map scales every leaf, beta too (:190)
        layers.append(lay)
    return w0.to(dtype), layers

"""CPU oracle for the stochastic-adjoint route.  TEST INFRASTRUCTURE ONLY.

Restates with torch CPU ops (fp64 by default, autograd for the vjps / jvps):
  * `sdeint_ito` (jaxsde/jaxsde/sdeint.py:11-13,24-27): `ito_integrate(..., fixed_grid=False)` = the `while _t < target_t` loop
    of `_stochastic_integrate` (sdeint.py:114-128; `next_t = min(_t + dt, ts[-1])`, so intermediate targets may be overshot),
    steps `ito_euler_step` / `ito_milstein_step` (:36-50) in the product form `g_prod` / `gdg_prod` (sde_utils.py:55-61);
  * its custom reverse pass `_sdeint_ito_rev` (jaxsde/jaxsde/sde_vjp.py:180-192): `vjp_ito_integrate` (:166-170) ->
    `vjp_integrate` (:137-163) with `time_reflect_ito` (sde_utils.py:30-37), `make_ito_adjoint_dynamics` (sde_vjp.py:13-47) and
    `make_milstein_prod` (:91-121); only the cotangent of the LAST output row is used (:186);
  * the virtual Brownian tree (jaxsde/jaxsde/brownian.py:12-95) on the Philox4x32-10 normals of
    bayesde_b200/csrc/common.cuh (`philox_normal`: counter (element, node, ., stream), Box-Muller), restated in numpy.

PARITY STATUS: **parity unpinned** by the reference (its tests for this path, jaxsde/tests/test_vjp.py, need JAX 0.1.77 and hold no
stored vectors).  Substitute pins (tests/test_adjoint.py): the adjoint gradient against fp64 autograd through the same solve as
dt -> 0 (the property test_vjp.py:88-114 checks), closed-form geometric Brownian motion, Brownian-tree laws.
"""
import math

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller, as in bayesde_b200/csrc/common.cuh
# ---------------------------------------------------------------------------------------------------------------
def philox4x32_10(c, k0, k1):
    c = [np.asarray(v, dtype=np.uint64) for v in c]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c


def philox_normal(seed, p, k, s, stream_id):
    """One standard normal per entry of `p` (element indices); fp32 Box-Muller like the device code."""
    p = np.asarray(p, dtype=np.uint64)
    full = lambda v: np.full(p.shape, v, dtype=np.uint64)
    c = philox4x32_10([p, full(k), full(s), full(stream_id)], seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = (c[0].astype(np.float32) + np.float32(1.0)) * np.float32(2.3283064365386963e-10)
    u2 = c[1].astype(np.float32) * np.float32(2.3283064365386963e-10)
    r = np.sqrt(np.float32(-2.0) * np.log(np.maximum(u1, np.float32(1e-37))))
    return (r * np.cos(np.float32(2.0 * math.pi) * u2.astype(np.float64)).astype(np.float32)).astype(np.float32)


class BrownianTree:
    """make_brownian_motion(t0, zeros(n), t1, rng, depth, rep) (brownian.py:92-95) as a callable t -> W(t) [n]."""

    def __init__(self, seed, stream_id, n, t0, t1, depth=10, rep=0, dtype=torch.float64):
        self.seed, self.sid, self.n, self.t0, self.t1, self.depth, self.rep, self.dtype = seed, stream_id, n, t0, t1, depth, rep, dtype
        e = np.arange(n)
        if rep:
            e[n - rep:] = e[n - 2 * rep:n - rep]      # sample_with_rep, brownian.py:69-74
        self.e = e

    def _z(self, node, kind):
        return philox_normal(self.seed, self.e, node, kind, self.sid).astype(np.float32)

    def __call__(self, t):
        f32 = np.float32
        t0, t1, t = f32(self.t0), f32(self.t1), f32(t)
        span = f32(t1 - t0)
        s = f32((t - t0) / span)                       # scaled_brownian_motion, brownian.py:12-17
        tl, tr = f32(0), f32(1)
        xl = np.zeros(self.n, dtype=f32)
        xr = self._z(0, 0)                             # w_end, brownian.py:80-82
        node = 1
        for _ in range(self.depth):                    # virtual_brownian_tree, brownian.py:26-51
            tm = f32((tl + tr) * f32(0.5))
            z = self._z(node, 1)
            xm = (xl * f32(tr - tm) + xr * f32(tm - tl)) / f32(tr - tl) + np.sqrt(f32(tr - tm) * f32(tm - tl) / f32(tr - tl)) * z
            if s < tm:
                tr, xr, node = tm, xm.astype(f32), 2 * node
            else:
                tl, xl, node = tm, xm.astype(f32), 2 * node + 1
        z = self._z(node, 1)
        mean = (xl * f32(tr - s) + xr * f32(s - tl)) / f32(tr - tl)
        sd = np.sqrt(max(f32(tr - s) * f32(s - tl), f32(0)) / f32(tr - tl))
        return torch.from_numpy((np.sqrt(span) * (mean + sd * z)).astype(f32)).to(self.dtype)


# ---------------------------------------------------------------------------------------------------------------
# autograd helpers (jax.vjp / jax.jvp for functions of tensors)
# ---------------------------------------------------------------------------------------------------------------
def _vjp(fun, primals, cot, create_graph=False):
    """fun(*primals) -> tensor; returns (value, tuple of cotangents w.r.t. primals) (zeros where unused)."""
    with torch.enable_grad():
        ps = [p.detach().requires_grad_(True) if not (create_graph and p.requires_grad) else p for p in primals]
        out = fun(*ps)
        if not out.requires_grad:
            return out, tuple(torch.zeros_like(p) for p in ps)
        gs = torch.autograd.grad(out, ps, cot, allow_unused=True, create_graph=create_graph)
    return out, tuple(torch.zeros_like(p) if g is None else g for g, p in zip(gs, ps))


def diag_jac(g_y, y, create_graph=False):
    """sde_utils.py:50-53: dg_i/dy_i of an elementwise g (jvp with ones)."""
    with torch.enable_grad():
        y_ = y if (create_graph and y.requires_grad) else y.detach().requires_grad_(True)
        gy = g_y(y_)
        if not gy.requires_grad:
            return torch.zeros_like(y)
        (d,) = torch.autograd.grad(gy.sum(), y_, create_graph=create_graph)
    return d


def make_gdg(g):
    """sde_utils.py:43-48: J_g(y) g(y) for an elementwise g."""
    def gdg(y, t, args):
        d = diag_jac(lambda yy: g(yy, t, args), y, create_graph=torch.is_grad_enabled())
        return d * g(y, t, args)
    return gdg


def make_gdg_prod(g_prod):
    """sde_utils.py:55-61."""
    def gdg_prod(y, t, args, v):
        g_y = lambda yy: g_prod(yy, t, args, torch.ones_like(yy))
        d = diag_jac(g_y, y, create_graph=torch.is_grad_enabled())
        return d * v * g_y(y)
    return gdg_prod


# ---------------------------------------------------------------------------------------------------------------
# forward solve: sdeint.py:53-64,81-138
# ---------------------------------------------------------------------------------------------------------------
def ito_integrate(f, g, y0, ts, bm, dt, args=(), gdg=None, g_prod=None, method="milstein", fixed_grid=False):
    if g_prod is None:
        g_prod = lambda y, t, a, noise: g(y, t, a) * noise
    if method not in ("milstein", "euler_maruyama"):
        raise ValueError("Unknown method: {}".format(method))
    if gdg is None:
        gdg = make_gdg_prod(g_prod)
    f32 = np.float32
    ts = [f32(v) for v in ts]
    dt = f32(dt)

    def step(y, t, noise, t_delta):
        out = y + float(t_delta) * f(y, float(t), args) + g_prod(y, float(t), args, noise)
        if method == "milstein":
            out = out + 0.5 * gdg(y, float(t), args, noise ** 2 - float(t_delta))
        return out

    ys, _t, _y = [y0], ts[0], y0
    for target in ts[1:]:
        if fixed_grid:          # sdeint.py:103-110 (quirk Q1)
            next_t = f32(target - _t)
            t_delta = f32(next_t - _t)
            _y = step(_y, _t, bm(next_t) - bm(_t), t_delta)
            _t = next_t
        else:                   # sdeint.py:114-128
            while _t < target:
                next_t = min(f32(_t + dt), ts[-1])
                t_delta = f32(next_t - _t)
                _y = step(_y, _t, bm(next_t) - bm(_t), t_delta)
                _t = next_t
        ys.append(_y)
    return torch.stack(ys)


# ---------------------------------------------------------------------------------------------------------------
# reverse pass: sde_vjp.py:13-47,91-121,137-192
# ---------------------------------------------------------------------------------------------------------------
def stochastic_adjoint(f, g, yT, v_yT, ts, bm, dt, flat_args, unravel, method="milstein"):
    """_sdeint_ito_rev: returns (y0 reconstructed, dL/dy0, dL/dflat_args)."""
    D, A = yT.numel(), flat_args.numel()
    flat_f = lambda y, t, fa: f(y, t, unravel(fa))
    flat_g = lambda y, t, fa: g(y, t, unravel(fa))
    gdg0 = make_gdg(flat_g)
    # time_reflect_ito, sde_utils.py:30-37
    rf = lambda y, t, fa: -flat_f(y, -t, fa) + gdg0(y, -t, fa)
    rg = lambda y, t, fa: flat_g(y, -t, fa)
    rb = lambda t: bm(-t)
    rts = [-np.float32(v) for v in list(ts)[::-1]]

    def unpack(aug):
        return aug[:D], aug[D:2 * D], aug[2 * D:]

    def aug_f(aug, t, fa):  # sde_vjp.py:17-31
        y, y_adj, _ = unpack(aug)
        fval, (f_a, f_args) = _vjp(lambda yy, aa: rf(yy, t, aa), (y, fa), -y_adj)
        _, (adj_dgdx, _) = _vjp(lambda yy, aa: rg(yy, t, aa), (y, fa), y_adj)
        _, (g_a, g_args) = _vjp(lambda yy, aa: rg(yy, t, aa), (y, fa), adj_dgdx)
        return torch.cat([fval.detach(), f_a + g_a, f_args + g_args])

    def aug_g_prod(aug, t, fa, v):  # sde_vjp.py:33-43
        y, y_adj, _ = unpack(aug)
        gval, (v_a, v_args) = _vjp(lambda yy, aa: rg(yy, t, aa), (y, fa), -y_adj * v)
        return torch.cat([gval.detach() * v, v_a, v_args])

    def aug_gdg(aug, t, fa, v):  # make_milstein_prod, sde_vjp.py:91-121
        y, adj, _ = unpack(aug)
        gfun = lambda yy, aa: rg(yy, t, aa)
        gval, (gdg_v, _) = _vjp(gfun, (y, fa), rg(y, t, fa).detach() * v)
        gval = gval.detach()
        dgdx = diag_jac(lambda yy: rg(yy, t, fa), y)
        _, (pp_adj, pp_args) = _vjp(gfun, (y, fa), adj * v * dgdx)
        c = adj * v * gval
        with torch.enable_grad():
            y_ = y.detach().requires_grad_(True)
            a_ = fa.detach().requires_grad_(True)
            gy = rg(y_, t, a_)
            if gy.requires_grad and y_.grad_fn is None:
                d = torch.autograd.grad(gy.sum(), y_, create_graph=True, allow_unused=True)[0]
            else:
                d = None
            if d is not None and d.requires_grad:
                mx = torch.autograd.grad((d * c).sum(), (y_, a_), allow_unused=True)
                m_adj = torch.zeros_like(y) if mx[0] is None else mx[0]
                m_args = torch.zeros_like(fa) if mx[1] is None else mx[1]
            else:
                m_adj, m_args = torch.zeros_like(y), torch.zeros_like(fa)
        return torch.cat([gdg_v, pp_adj - m_adj, pp_args - m_args])

    aug0 = torch.cat([yT.detach(), v_yT.detach(), torch.zeros(A, dtype=yT.dtype)])
    with torch.no_grad():
        out = ito_integrate(aug_f, None, aug0, rts, rb, dt, flat_args.detach(), gdg=aug_gdg, g_prod=aug_g_prod, method=method)
    y0_rec, y_adj, arg_adj = unpack(out[-1])
    return y0_rec, y_adj, arg_adj

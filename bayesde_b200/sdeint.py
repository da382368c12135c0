"""`sdeint_ito_fixed_grid` -- drop-in for jaxsde/jaxsde/sdeint.py:16-18 on torch CUDA tensors.

    ys = sdeint_ito_fixed_grid(f, g, y0, ts, rng, args=(), dt=1e-6, method='milstein', rep=0)

Two routes, both on libbsde.so kernels (no CPU fallback):
  * structured: `f`/`g` are the `f_aug`/`g_aug` of a `bayesde_b200.brax` SDEBNN block -> fused
    K1 (weight-SDE rows) + K2/K3 (x rows) kernels, reverse pass included;
  * generic: arbitrary torch callables f(y,t,args), g(y,t,args) on a flat state -> the callables run
    as torch ops on the GPU, the step itself (ito_euler_step / ito_milstein_step, sdeint.py:36-50,
    plus euler_heun) is one fused elementwise kernel, increments come from the Philox kernel.

`rng` is an int seed, a `(seed, stream_id)` pair (Philox), a tensor of explicit increments (len(ts)-1, D) -- the "supplied
increments" mode used for exact comparison with the reference -- or a `PRNGKey` (uint32 pair): the increments then come from the
reference's own virtual Brownian tree on JAX's threefry stream (jaxsde/jaxsde/brownian.py).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._fused import reference_time_grid

_METHODS = ("milstein", "euler_maruyama", "euler_heun")


def _parse_rng(rng):
    if isinstance(rng, torch.Tensor):
        return None, None, rng
    if isinstance(rng, (tuple, list)) and len(rng) == 2:
        return int(rng[0]), int(rng[1]), None
    return int(rng), 0, None


def philox_increments(seed, stream_id, S, dtk, P, device):
    """[S, nit, P] increments N(0, dt_k) from the in-kernel Philox stream (bsde_philox_increments)."""
    L = _lib.lib()
    dtk_d = torch.as_tensor(np.asarray(dtk, dtype=np.float32)).to(device)
    out = torch.empty(S, len(dtk), P, device=device, dtype=torch.float32)
    st = L.bsde_philox_increments(C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_uint32(int(stream_id) & 0xFFFFFFFF),
                                  S, len(dtk), P, _lib.ptr(dtk_d), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(st, "bsde_philox_increments")
    return out


class _StepUpdate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, method, dt, y, f, g, dW, gdg, g2):
        L = _lib.lib()
        y1 = torch.empty_like(y)
        st = L.bsde_step_update(method, y.numel(), float(dt), _lib.ptr(y), _lib.ptr(f), _lib.ptr(g), _lib.ptr(dW),
                                _lib.ptr(gdg), _lib.ptr(g2), _lib.ptr(y1), _lib.stream_ptr())
        _lib.check(st, "bsde_step_update")
        ctx.method, ctx.dt = method, float(dt)
        ctx.save_for_backward(dW)
        ctx.has = (gdg is not None, g2 is not None)
        return y1

    @staticmethod
    def backward(ctx, gy1):
        (dW,) = ctx.saved_tensors
        dt = ctx.dt
        heun = ctx.method == 2 and ctx.has[1]
        gg = gy1 * dW * (0.5 if heun else 1.0)
        ggdg = 0.5 * gy1 * (dW * dW - dt) if (ctx.method == 1 and ctx.has[0]) else None
        return None, None, gy1, gy1 * dt, gg, None, ggdg, (gg if heun else None)


def _gdg(g, y, t, args):
    """J_g(y) applied to g(y) for an elementwise (diagonal) g: jaxsde/jaxsde/sde_utils.py:55-61.

    The reference differentiates THROUGH this jvp (reverse mode over `make_gdg_prod`), so when gradients are being
    recorded the product stays attached to the graph of `y` and of g's parameters (second-order terms included);
    otherwise it is evaluated on a detached copy."""
    track = torch.is_grad_enabled()
    with torch.enable_grad():
        y_ = y if (track and y.requires_grad) else y.detach().requires_grad_(True)
        gy = g(y_, t, args)
        if not gy.requires_grad or gy.grad_fn is None:
            return None  # state-independent diffusion: the Milstein correction vanishes
        (diag,) = torch.autograd.grad(gy.sum(), y_, create_graph=track, allow_unused=True)
        if diag is None:
            return None
        out = diag * gy
    return out if track else out.detach()


def sdeint_ito_fixed_grid(f, g, y0, ts, rng, args=(), dt=1e-6, method="milstein", rep=0, time_grid="reference", *,
                          _layer_view=False):
    """Fixed-grid Ito solve; returns ys of shape (len(ts), D) including y0 (sdeint.py:138) -- also on the structured route
    (`f`, `g` = the `f_aug` / `g_aug` of an SDEBNN block: D = x_dim + 2P per sample, a leading sample axis is kept when y0 has
    one), so `ys[-1][:x_dim]`, `ys[:, x_dim:x_dim+P]` and `ys[-1][x_dim+P:].sum()` read as in brax.py:117-127.  `_layer_view`
    is the SDEBNN layer's private short cut to (x_T, kl, w trajectory) without materialising `ys`.

    `dt` is ignored on the fixed grid, as in the reference.  `rep` ties the noise of the last `rep`
    state entries to the `rep` entries before them (jaxsde/jaxsde/brownian.py:69-74).
    `time_grid="reference"` reproduces the scan's time bookkeeping (sdeint.py:105-106),
    "uniform" integrates over ts as written."""
    if method not in _METHODS:
        raise ValueError("Unknown method: {}".format(method))  # sdeint.py:62
    owner = getattr(f, "__self__", None)
    if owner is not None and getattr(owner, "_bsde_structured", False) and getattr(g, "__self__", None) is owner:
        return owner._solve(y0, ts, rng, args, method, rep, time_grid, want_ys=not _layer_view)

    if not y0.is_cuda:
        raise _lib.BsdeError("sdeint_ito_fixed_grid needs CUDA tensors (there is no CPU fallback)")
    tk, dtk = reference_time_grid(ts.detach().cpu().numpy() if isinstance(ts, torch.Tensor) else ts, time_grid)
    D = y0.numel()
    if is_jax_key(rng):  # the reference's own path: make_brownian_motion(ts[0], zeros(D), ts[-1], rng, rep=rep), sdeint.py:17
        tsl = [float(v) for v in (ts.detach().cpu().tolist() if isinstance(ts, torch.Tensor) else ts)]
        rng = tree_increments(BrownianTree(tsl[0], D, tsl[-1], rng, depth=10, rep=rep, device=y0.device), tsl, time_grid)
    seed, sid, inc = _parse_rng(rng)
    if inc is None:
        inc = philox_increments(seed, sid, 1, dtk, D, y0.device)[0]
        if rep:
            inc[:, D - rep:] = inc[:, D - 2 * rep:D - rep]
    inc = inc.to(torch.float32).contiguous()
    if tuple(inc.shape) != (len(tk), D):
        raise ValueError(f"increments must have shape {(len(tk), D)}, got {tuple(inc.shape)}")
    m = _lib.METHOD[method]
    y = y0.to(torch.float32).contiguous()
    ys = [y]
    for k in range(len(tk)):
        t = torch.tensor(float(tk[k]), device=y.device)
        h = float(dtk[k])
        fy = f(y, t, args).contiguous()
        gy = g(y, t, args).contiguous()
        gdg = g2 = None
        if method == "milstein":
            gdg = _gdg(g, y, t, args)
            gdg = gdg.contiguous() if gdg is not None else None
        elif method == "euler_heun":
            g2 = g(y + gy * inc[k], t + h, args).contiguous()
        y = _StepUpdate.apply(m, h, y, fy, gy, inc[k], gdg, g2)
        ys.append(y)
    return torch.stack(ys)


# ------------------------------------------------------------------------------------------------------------------
# sdeint_ito: variable-step solve with the stochastic-adjoint reverse pass
# ------------------------------------------------------------------------------------------------------------------
def PRNGKey(seed):
    """jax.random.PRNGKey(seed): the uint32 pair [seed >> 32, seed & 0xFFFFFFFF] (numpy, host)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def is_jax_key(rng):
    return isinstance(rng, np.ndarray) and rng.dtype == np.uint32 and rng.shape == (2,)


def split(key, num=2):
    """jax.random.split(key, num) on the host (bsde_threefry_split): (num, 2) uint32."""
    if not is_jax_key(key):
        raise ValueError("split needs a PRNGKey (numpy uint32 array of shape (2,))")
    out = (C.c_uint32 * (2 * int(num)))()
    _lib.check(_lib.lib().bsde_threefry_split(int(key[0]), int(key[1]), int(num), out), "bsde_threefry_split")
    return np.frombuffer(out, dtype=np.uint32).reshape(int(num), 2).copy()


class BrownianTree:
    """`make_brownian_motion(t0, zeros(n), t1, rng, depth, rep)` (jaxsde/jaxsde/brownian.py:92-95) as a callable
    t -> W(t) [n] on the device: every query is one launch (O(1) memory, any order of queries).
      * rng = int seed or (seed, stream_id): Philox4x32-10 draws keyed by (seed, stream_id, element, tree node) (bsde_brownian_tree);
      * rng = PRNGKey(...) (uint32 pair): the reference's own stream -- jax.random.split / jax.random.normal on threefry keys
        (bsde_brownian_tree_threefry), i.e. the path a JAX run of the reference saw.
    `query(t, first, count)` returns the slice W(t)[first:first+count] without drawing the rest (the x rows of an SDEBNN state
    have g = 0: quirk Q13)."""

    def __init__(self, t0, n, t1, rng, depth=10, rep=0, device="cuda"):
        self.key = None
        if is_jax_key(rng):
            self.key, self.seed, self.sid = (int(rng[0]), int(rng[1])), None, None
        else:
            seed, sid, inc = _parse_rng(rng)
            if inc is not None:
                raise ValueError("a Brownian tree needs an integer seed, a (seed, stream_id) pair or a PRNGKey")
            self.seed, self.sid = seed, sid
        self.n, self.depth, self.rep = int(n), int(depth), int(rep)
        self.t0, self.t1, self.device = float(np.float32(t0)), float(np.float32(t1)), torch.device(device)

    def query(self, t, first=0, count=None):
        count = self.n - first if count is None else int(count)
        out = torch.empty(count, device=self.device, dtype=torch.float32)
        L = _lib.lib()
        if self.key is not None:
            st = L.bsde_brownian_tree_threefry(self.key[0], self.key[1], self.n, self.rep, self.depth, self.t0, self.t1,
                                               float(np.float32(t)), int(first), count, _lib.ptr(out), _lib.stream_ptr())
            _lib.check(st, "bsde_brownian_tree_threefry")
            return out
        if first != 0 or count != self.n:
            full = self.query(t)
            return full[first:first + count].contiguous()
        st = L.bsde_brownian_tree(C.c_uint64(self.seed & 0xFFFFFFFFFFFFFFFF), C.c_uint32(self.sid & 0xFFFFFFFF), self.n,
                                  self.rep, self.depth, self.t0, self.t1, float(np.float32(t)), _lib.ptr(out), _lib.stream_ptr())
        _lib.check(st, "bsde_brownian_tree")
        return out

    def __call__(self, t):
        return self.query(t)


def tree_increments(bm, ts, time_grid="reference", first=0, count=None):
    """The increments `bm(next_t) - bm(curr_t)` the fixed-grid scan draws (jaxsde/jaxsde/sdeint.py:103-110), one row per scan
    iteration, with the scan's own time bookkeeping (quirk Q1) or on the uniform grid: (len(ts)-1, count)."""
    f32 = np.float32
    ts = [f32(v) for v in (ts.detach().cpu().tolist() if isinstance(ts, torch.Tensor) else ts)]
    rows, curr = [], ts[0]
    for k in range(1, len(ts)):
        nxt = f32(ts[k] - curr) if time_grid == "reference" else ts[k]
        rows.append(bm.query(nxt, first, count) - bm.query(curr, first, count))
        curr = nxt
    return torch.stack(rows)


def make_brownian_motion(t0, x0, t1, rng, depth=10, rep=0):
    """jaxsde/jaxsde/brownian.py:92-95; `x0` only gives the shape and the device."""
    return BrownianTree(t0, x0.numel(), t1, rng, depth, rep, x0.device)


_state_dtype = torch.float32  # the device path computes in fp32 ("all arithmetic is fp32")


def _require_cuda(y0):
    if not y0.is_cuda:
        raise _lib.BsdeError("sdeint_ito needs CUDA tensors (there is no CPU fallback)")


def _leaves(tree):
    if isinstance(tree, torch.Tensor):
        return [tree]
    if isinstance(tree, (tuple, list)):
        return [x for t in tree for x in _leaves(t)]
    if isinstance(tree, dict):
        return [x for k in sorted(tree) for x in _leaves(tree[k])]
    return []


def _rebuild(tree, it):
    if isinstance(tree, torch.Tensor):
        return next(it)
    if isinstance(tree, tuple):
        return tuple(_rebuild(t, it) for t in tree)
    if isinstance(tree, list):
        return [_rebuild(t, it) for t in tree]
    if isinstance(tree, dict):
        return {k: _rebuild(tree[k], it) for k in sorted(tree)}
    return tree


def _prod_step(dt, y, fy, gp, mp):
    """y + dt*f + g_prod [+ 0.5*gdg_prod]: ito_euler_step / ito_milstein_step in product form (sdeint.py:36-50)."""
    y1 = torch.empty_like(y)
    st = _lib.lib().bsde_step_update_prod(y.numel(), float(dt), _lib.ptr(y), _lib.ptr(fy.contiguous()), _lib.ptr(gp.contiguous()),
                                          _lib.ptr(mp.contiguous()) if mp is not None else None, _lib.ptr(y1), _lib.stream_ptr())
    _lib.check(st, "bsde_step_update_prod")
    return y1


def _integrate_prod(f, g_prod, gdg_prod, y0, ts, bm, dt, method):
    """The `fixed_grid=False` branch of `_stochastic_integrate` (sdeint.py:114-135): between output times step by `dt`, clipped
    at ts[-1] only (so interior output times may be overshot, as in the reference); fp32 time arithmetic on the host.
    Returns the list of states at the output times."""
    f32 = np.float32
    ts = [f32(v) for v in ts]
    dt = f32(dt)
    ys, _t, _y = [y0], ts[0], y0
    w_t = None
    for target in ts[1:]:
        while _t < target:
            next_t = min(f32(_t + dt), ts[-1])
            t_delta = f32(next_t - _t)
            if w_t is None:
                w_t = bm(_t)
            w_next = bm(next_t)
            dw = w_next - w_t
            fy = f(_y, float(_t))
            gp = g_prod(_y, float(_t), dw)
            mp = gdg_prod(_y, float(_t), dw * dw - float(t_delta)) if method == "milstein" else None
            _y = _prod_step(t_delta, _y, fy, gp, mp)
            _t, w_t = next_t, w_next
        ys.append(_y)
    return ys


def _grads(out, inputs, cot, create_graph=False):
    """vjp of `out` w.r.t. `inputs` with cotangent `cot`; zeros for inputs that `out` does not depend on."""
    if not out.requires_grad:
        return [torch.zeros_like(p) for p in inputs]
    gs = torch.autograd.grad(out, inputs, cot, allow_unused=True, create_graph=create_graph, retain_graph=True)
    return [torch.zeros_like(p) if g is None else g for g, p in zip(gs, inputs)]


class _SdeintIto(torch.autograd.Function):
    """`_sdeint_ito` with its custom vjp (sdeint.py:21-24, sde_vjp.py:180-192).  Forward: variable-step Ito solve on the Philox
    Brownian tree.  Backward: the augmented system [y, dL/dy, dL/dargs] of `make_ito_adjoint_dynamics` (sde_vjp.py:13-47) with the
    Milstein product of `make_milstein_prod` (:91-121) is integrated over reflected time (`time_reflect_ito`, sde_utils.py:30-37)
    on the SAME tree (queried at -t), starting from the last output row only (sde_vjp.py:186); memory is O(1) in the step count."""

    @staticmethod
    def forward(ctx, f, g, dt, method, rng, rep, ts, args, y0, *leaves):
        bm = make_brownian_motion(ts[0], y0, ts[-1], rng, depth=10, rep=rep)
        with torch.no_grad():
            fn = lambda y, t: f(y, torch.tensor(t, device=y.device), args).contiguous()
            gp = lambda y, t, v: g(y, torch.tensor(t, device=y.device), args) * v
            def gdgp(y, t, v):
                tt = torch.tensor(t, device=y.device)
                d = _gdg(g, y, tt, args)
                return None if d is None else d * v
            ys = _integrate_prod(fn, gp, gdgp, y0.detach().to(_state_dtype).contiguous(), ts, bm, dt, method)
        out = torch.stack(ys)
        ctx.f, ctx.g, ctx.dt, ctx.method, ctx.ts, ctx.args, ctx.bm = f, g, dt, method, ts, args, bm
        ctx.save_for_backward(out[-1], *leaves)
        return out

    @staticmethod
    def backward(ctx, g_ys):
        yT, *leaves = ctx.saved_tensors
        f, g, args = ctx.f, ctx.g, ctx.args
        D = yT.numel()
        shapes = [l.shape for l in leaves]
        sizes = [l.numel() for l in leaves]
        A = sum(sizes)
        fa0 = torch.cat([l.detach().reshape(-1).to(_state_dtype) for l in leaves]) if leaves else yT.new_zeros(0)

        def unravel(fa):
            parts = [p.reshape(s) for p, s in zip(torch.split(fa, sizes), shapes)] if sizes else []
            return _rebuild(args, iter(parts))

        def rg(y, t, fa):  # reflected diffusion, sde_utils.py:35
            return g(y, torch.tensor(-t, device=y.device), unravel(fa))

        def dgdy(y, t, fa, create_graph):  # diag_jac, sde_utils.py:50-53 (elementwise g)
            gy = rg(y, t, fa)
            if not (gy.requires_grad and y.requires_grad):
                return None, gy
            (d,) = torch.autograd.grad(gy.sum(), y, create_graph=create_graph, retain_graph=True, allow_unused=True)
            return d, gy

        def rf(y, t, fa):  # reflected drift incl. the Ito correction g dg/dy, sde_utils.py:34
            d, gy = dgdy(y, t, fa, True)
            out = -f(y, torch.tensor(-t, device=y.device), unravel(fa))
            return out if d is None else out + d * gy

        def split(aug):
            return aug[:D], aug[D:2 * D], aug[2 * D:]

        def prim(aug):
            y, adj, _ = split(aug)
            return y.detach().requires_grad_(True), fa0.detach().requires_grad_(True), adj

        def aug_f(aug, t):  # sde_vjp.py:17-31
            with torch.enable_grad():
                y, fa, adj = prim(aug)
                fval = rf(y, t, fa)
                f_a, f_args = _grads(fval, (y, fa), -adj)
                gval = rg(y, t, fa)
                (adj_dgdx,) = _grads(gval, (y,), adj)
                g_a, g_args = _grads(gval, (y, fa), adj_dgdx)
            return torch.cat([fval.detach(), f_a + g_a, f_args + g_args])

        def aug_g_prod(aug, t, v):  # sde_vjp.py:33-43
            with torch.enable_grad():
                y, fa, adj = prim(aug)
                gval = rg(y, t, fa)
                v_a, v_args = _grads(gval, (y, fa), -adj * v)
            return torch.cat([gval.detach() * v, v_a, v_args])

        def aug_gdg(aug, t, v):  # make_milstein_prod, sde_vjp.py:91-121
            with torch.enable_grad():
                y, fa, adj = prim(aug)
                d, gval = dgdy(y, t, fa, True)
                gv = gval.detach()
                (gdg_v,) = _grads(gval, (y,), gv * v)
                if d is None:
                    zero = torch.zeros_like(y)
                    return torch.cat([gdg_v, zero, torch.zeros_like(fa)])
                pp_adj, pp_args = _grads(gval, (y, fa), adj * v * d.detach())
                m_adj, m_args = _grads((d * (adj * v * gv)).sum(), (y, fa), None)
            return torch.cat([gdg_v, pp_adj - m_adj, pp_args - m_args])

        rts = [-np.float32(v) for v in list(ctx.ts)[::-1]]  # time_reflect_ito, sde_utils.py:37
        rb = lambda t: ctx.bm(-t)
        aug0 = torch.cat([yT.detach(), g_ys[-1].detach().to(_state_dtype), yT.new_zeros(A)]).contiguous()
        out = _integrate_prod(aug_f, aug_g_prod, aug_gdg, aug0, rts, rb, ctx.dt, ctx.method)[-1]
        _, y_adj, arg_adj = split(out)
        g_leaves = [p.reshape(s).to(l.dtype) for p, s, l in zip(torch.split(arg_adj, sizes), shapes, leaves)] if sizes else []
        return (None,) * 8 + (y_adj, *g_leaves)


def sdeint_ito(f, g, y0, ts, rng, args=(), dt=1e-6, method="milstein", rep=0):
    """Drop-in for jaxsde/jaxsde/sdeint.py:11-13 on CUDA tensors: Ito solve with steps of `dt` between the output times `ts`,
    differentiable w.r.t. `y0` and the tensors in `args` through the stochastic adjoint (O(1) memory); returns (len(ts), D).
    `rng` is an int seed or a (seed, stream_id) pair; gradients w.r.t. `ts` are zero as in the reference (sde_vjp.py:192)."""
    if method not in ("milstein", "euler_maruyama"):
        raise ValueError("Unknown method: {}".format(method))  # sdeint.py:62
    owner = getattr(f, "__self__", None)
    if owner is not None and getattr(owner, "_bsde_structured", False) and getattr(g, "__self__", None) is owner:
        return owner._solve_adjoint(y0, ts, rng, args, dt, method, rep)  # SDEBNN dynamics: fused K1/K2/K3 steps
    _require_cuda(y0)
    tsl = [float(v) for v in (ts.detach().cpu().tolist() if isinstance(ts, torch.Tensor) else ts)]
    return _SdeintIto.apply(f, g, float(dt), method, rng, int(rep), tsl, args, y0, *_leaves(args))

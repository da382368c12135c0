"""Dev probe: reconstruction error and gradient agreement of the stochastic-adjoint route of one SDEBNN block vs its step size."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesde_b200 import _fused, arch, brax
from bayesde_b200.sdeint import sdeint_ito, sdeint_ito_fixed_grid, BrownianTree
dev = "cuda:0"
S, B, H, Cc, nsteps, seed = 8, int(os.environ.get("PB", 128)), 32, 5, 20, 5
math = os.environ.get("PMATH", "tc")
SIG = float(os.environ.get("PSIG", 0.2))
KLW = float(os.environ.get("PKLW", 1e-6))
fw = arch.MLP((2, 64, 2), actfn="softplus", xt=True, ou_dw=True)
fx = arch.FxSpec(0, (1, H, H, Cc), 64, "softplus")
P = fx.P
g = torch.Generator().manual_seed(0)
_, fw_params = fw.init(3, (P,))
fw_params = brax._to_device(fw_params, dev)
lays = arch.fw_layers(fw_params)
with torch.no_grad():
    for t in lays[-1]:
        t.normal_(0.0, 1e-2)
leaves = [t.requires_grad_(True) for lay in lays for t in lay]
w0 = (fx.init(1).to(dev) * float(os.environ.get("PW0", 1.0))).requires_grad_(True)
x = torch.randn(S, B, H, H, Cc, generator=g).to(dev).requires_grad_(True)
cot = torch.randn(S, B, H, H, Cc, generator=g).to(dev)
ts_lin = torch.linspace(0.0, 1.0, nsteps)

def run(route, dyn):
    y0 = torch.cat([x.reshape(S, -1), w0.expand(S, P), x.new_zeros(S, P)], dim=1)
    xo, kl, _ = route(y0)
    loss = (xo * cot).mean() + KLW * kl.sum()
    return xo.detach(), kl.detach(), torch.autograd.grad(loss, [x, w0] + leaves)

for sub in (1, 4):
    dt = 1.0 / (19 * sub)
    dynA = brax._AugDynamics(fx, fw.spec, x.shape, SIG, False, True, math, fx.init(0).to(dev), 0)
    xo_a, kl_a, g_a = run(lambda y0: sdeint_ito(dynA.f_aug, dynA.g_aug, y0, ts_lin, seed, fw_params, dt=dt, method="euler_maruyama", rep=0), dynA)
    x0r, w0r = _fused.LAST_RECONSTRUCTION
    steps, _ = _fused.ito_step_schedule(ts_lin.tolist(), dt)
    tree = BrownianTree(0.0, S * P, 1.0, (seed, 0), device=dev)
    inc = torch.stack([tree(tn) - tree(t) for (t, h, tn) in steps]).reshape(len(steps), S, P).permute(1, 0, 2).contiguous()
    ts_sched = torch.tensor([float(steps[0][0])] + [float(s[2]) for s in steps])
    dynB = brax._AugDynamics(fx, fw.spec, x.shape, SIG, False, True, math, fx.init(0).to(dev), 0, remat=True)
    xo_b, kl_b, g_b = run(lambda y0: sdeint_ito_fixed_grid(dynB.f_aug, dynB.g_aug, y0, ts_sched, inc, fw_params, method="euler_maruyama", rep=0, time_grid="uniform", _layer_view=True), dynB)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    cos = lambda a, b: float(torch.dot(a.flatten().double(), b.flatten().double()) / (a.double().norm() * b.double().norm()))
    print(f"sub={sub} steps={len(steps)} fwd rel {rel(xo_a, xo_b):.2e} | x_T norm {float(xo_a.norm()):.3e} x0 norm {float(x.norm()):.3e} | rec x rel {rel(x0r, x.detach()):.3e} rec w rel {rel(w0r, w0.detach().expand(S, P)):.3e} | "
          f"cos gx {cos(g_a[0], g_b[0]):.4f} gw0 {cos(g_a[1], g_b[1]):.4f} gW4 {cos(g_a[-5], g_b[-5]):.4f} | norm ratio gx {float(g_a[0].norm() / g_b[0].norm()):.3f}", flush=True)

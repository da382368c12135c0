"""Measurements for the SURVEY 8.6 "next" rows (N1-N4) at the BASELINE sizes, one JSON line per row.

    python tools/bench_next_rows.py [--reps 5]

  N1  stochastic-adjoint route of one config-4 group-0 SDEBNN block (S = 8 x B = 128, 32x32x5, fx 64, TF32 convs, dt = 1/19) against the
      checkpointing fixed-grid route on the same tree increments: ms per forward + reverse pass, peak memory.
  N2  Brownian-tree queries (Philox and threefry) for the S*P = 651 664 weight rows of that block and for a full D = 818 276 entry
      state: us per query, normals per second (11 + 1 draws per entry per query), output GB/s.
  N3  evaluation path: predict() of the config-4 classifier, S = 8 samples x 128 images: images/s.
  N4  config-5 SDENet training step (B = 128, midpoint dt 0.05): backprop-through-solver vs adjoint=True vs adaptive=True:
      ms per step, peak memory, accepted steps.
  C1  config 1 (the JAX toy, examples/jax/sdebnn_toy1d.py --activn swish --aug_dim 2 --stop_grad --prior_dw ou): one ELBO
      value_and_grad (10 samples x 100 points, hidden 64, P = 5132, posterior + prior solve of 19 steps) through the fused
      SDELayer launches, and the oracle restatement (torch CPU fp64, all host threads) beside it.
All times are CUDA events on the launching stream after warm-up; memory is torch.cuda.max_memory_allocated over the timed region."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesde_b200 import _fused, arch, brax, torch_models as tm  # noqa: E402
from bayesde_b200.sdeint import BrownianTree, PRNGKey, sdeint_ito, sdeint_ito_fixed_grid  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--reps", type=int, default=5)
p.add_argument("--rows", default="N1,N2,N3,N4,C1")
a = p.parse_args()
dev = "cuda:0"
rows = set(a.rows.split(","))


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (torch.cuda.max_memory_allocated() - base) / 2 ** 20


if "N2" in rows:
    for n, label in ((8 * 81458, "S*P weight rows of a group-0 block"), (655360 + 2 * 81458, "full state D of one sample")):
        for kind, rng in (("philox", (5, 0)), ("threefry", PRNGKey(5))):
            tree = BrownianTree(0.0, n, 1.0, rng, device=dev)
            ts = np.linspace(0.01, 0.99, 64)
            k = [0]

            def q():
                k[0] += 1
                return tree(float(ts[k[0] % 64]))
            ms, _ = timed(q, 64, warm=4)
            print(json.dumps({"row": "N2", "what": f"Brownian tree query, {kind}, n = {n} ({label}), depth 10", "us_per_query": ms * 1e3,
                              "normals_per_s": 12 * n / (ms * 1e-3), "output_GBps": 4 * n / (ms * 1e-3) / 1e9}), flush=True)

if "N1" in rows:
    S, B, H, Cc, nsteps, seed = 8, 128, 32, 5, 20, 5
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
    w0 = (0.5 * fx.init(1).to(dev)).requires_grad_(True)
    x = torch.randn(S, B, H, H, Cc, generator=g).to(dev).requires_grad_(True)
    cot = torch.randn(S, B, H, H, Cc, generator=g).to(dev)
    ts_lin = torch.linspace(0.0, 1.0, nsteps)
    dt = 1.0 / (nsteps - 1)
    steps, _ = _fused.ito_step_schedule(ts_lin.tolist(), dt)
    tree = BrownianTree(0.0, S * P, 1.0, (seed, 0), device=dev)
    inc = torch.stack([tree(tn) - tree(t) for (t, h, tn) in steps]).reshape(len(steps), S, P).permute(1, 0, 2).contiguous()
    ts_sched = torch.tensor([float(steps[0][0])] + [float(s[2]) for s in steps])

    def run(route):
        y0 = torch.cat([x.reshape(S, -1), w0.expand(S, P), x.new_zeros(S, P)], dim=1)
        xo, kl, _ = route(y0)
        torch.autograd.grad((xo * cot).mean() + 1e-12 * kl.sum(), [x, w0] + leaves)

    for name, remat in (("checkpointing (activations kept)", False), ("checkpointing (remat: x_k kept, activations recomputed)", True)):
        dyn = brax._AugDynamics(fx, fw.spec, x.shape, 1e-3, False, True, "tc", fx.init(0).to(dev), 0, remat=remat)
        ms, mem = timed(lambda: run(lambda y0: sdeint_ito_fixed_grid(dyn.f_aug, dyn.g_aug, y0, ts_sched, inc, fw_params,
                                                                     method="euler_maruyama", rep=0, time_grid="uniform", _layer_view=True)), a.reps)
        print(json.dumps({"row": "N1", "what": f"config-4 group-0 block, {len(steps)} steps, fwd + reverse, {name}", "ms": ms,
                          "images_per_s": B / (ms * 1e-3), "peak_MiB": mem}), flush=True)
    dyn = brax._AugDynamics(fx, fw.spec, x.shape, 1e-3, False, True, "tc", fx.init(0).to(dev), 0)
    ms, mem = timed(lambda: run(lambda y0: sdeint_ito(dyn.f_aug, dyn.g_aug, y0, ts_lin, seed, fw_params, dt=dt,
                                                     method="euler_maruyama", rep=0)), a.reps)
    print(json.dumps({"row": "N1", "what": f"config-4 group-0 block, {len(steps)} steps, fwd + reverse, stochastic adjoint (sdeint_ito)",
                      "ms": ms, "images_per_s": B / (ms * 1e-3), "peak_MiB": mem}), flush=True)

if "N3" in rows:
    from bayesde_b200 import evaluation
    from bayesde_b200.classification import SDEBNNClassifier
    model = SDEBNNClassifier(nsamples=8, diff_coef=0.2, math="tc", w_init=1e-2, b_init=1e-2, seed=0, device=dev,
                             input_size=(32, 32, 3), aug=2, nblocks=(2, 2, 2), fx_dim=64, fw_dims=(2, 64, 2), nsteps=20)
    xb = torch.randn(128, 32, 32, 3, device=dev)
    with torch.no_grad():
        ms, mem = timed(lambda: evaluation.predict(model, xb, 7, nsamples=8), a.reps)
    print(json.dumps({"row": "N3", "what": "predict(): config-4 classifier, 8 samples x 128 images, forward only", "ms": ms,
                      "images_per_s": 128 / (ms * 1e-3), "peak_MiB": mem}), flush=True)

if "N4" in rows:
    torch.manual_seed(0)
    m = tm.SDENet(input_size=(3, 32, 32), blocks=(2, 2, 2), weight_network_sizes=(1, 128, 1), hidden_width=32, sigma=0.1,
                  inhomogeneous=True, math="tc").to(dev)
    with torch.no_grad():
        last = m.w_net.linears()[-1]
        last.weight.normal_(0, 1e-2), last.bias.normal_(0, 1e-2)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    xb = torch.randn(128, 3, 32, 32, device=dev)
    yb = torch.randint(0, 10, (128,), device=dev)

    def step(**kw):
        opt.zero_grad(set_to_none=True)
        logits, logqp = m(xb, dt=0.05, method="midpoint", **kw)
        (F.cross_entropy(logits, yb) + 1e-4 * logqp).backward()
        opt.step()

    for name, kw in (("backprop through the solver", {}), ("adjoint=True", dict(adjoint=True)),
                     ("adaptive=True (rtol 1e-3, atol 1e-3)", dict(adaptive=True, rtol=1e-3, atol=1e-3))):
        ms, mem = timed(lambda: step(**kw), a.reps)
        extra = {"accepted_steps": len(m.adaptive_grid[0]), "nfe": m.nfe} if "adaptive" in kw else {"nfe": m.nfe}
        print(json.dumps({"row": "N4", "what": f"config-5 SDENet train step, B = 128, midpoint dt 0.05, {name}", "ms": ms,
                          "images_per_s": 128 / (ms * 1e-3), "peak_MiB": mem, **extra}), flush=True)


if "C1" in rows:
    import time
    from bayesde_b200 import _lib, brax_layers as BL
    st = dict(driftw=True, diffw=True, diff_const=0.1, prior_driftw="ou", driftw_scale=0.1, stop_grad=True, aug_dim=2,
              priorw0_sigma=-1.0, diff_drift=True)
    (init_fn, apply_fn), P = BL.mlp(0, 1, 64, "swish", st)
    _, params = init_fn(1, (-1, 1))
    to_dev = lambda tr: BL._tree_map(lambda t: t.to(dev).requires_grad_(True), tr)
    params = [(), tuple(to_dev(q) if q is not None else None for q in params[1])]
    xin = torch.linspace(-1, 1, 100, device=dev).reshape(100, 1)
    tgt = torch.cos(3 * xin)
    leaves = [t for q in params[1][0] for t in q] + [t for q in params[1][1][1] for t in q]
    k = [0]

    def step():
        k[0] += 1
        loss = BL.elbo(params, (xin, tgt), apply_fn, k[0], 1.0, num_samples=10)
        torch.autograd.grad(loss, leaves)
    L = _lib.lib()
    n0 = L.bsde_launch_count()
    step()
    launches = L.bsde_launch_count() - n0
    ms, mem = timed(step, a.reps * 4)
    row = {"row": "C1", "what": "config 1 JAX toy: elbo value_and_grad, 10 samples x 100 points, hidden 64, P = 5132, swish, STL, OU prior "
                                "(posterior + prior solve, 19 steps each)", "ms": ms, "steps_per_s": 1e3 / ms, "bsde_launches_per_step": int(launches),
           "peak_MiB": mem}
    try:   # the oracle restatement on the host cores (test infrastructure; timed here only as the CPU figure beside the GPU one)
        from oracle import jax_toy as jt
        g_ = torch.Generator().manual_seed(0)
        orc = jt.SDELayerOracle(100, 3, 64, "swish", jt.ToySetting())
        w0, lays = jt.init_toy(100, 3, 64, "swish", g_, driftw_scale=0.1, last_std=0.05)
        x = torch.randn(100, 3, generator=g_, dtype=torch.float64)
        eps = torch.randn(10, 19, orc.P, generator=g_, dtype=torch.float64)
        w0 = w0.requires_grad_(True)
        lays = [tuple(t.requires_grad_(True) for t in lay) for lay in lays]

        def cpu_step():
            outs = [orc.apply(w0, lays, x, eps[s_]) for s_ in range(10)]
            loss = sum(o[0][..., -1].abs().mean() + o[2] for o in outs) / 10
            torch.autograd.grad(loss, [w0] + [t for lay in lays for t in lay])
        cpu_step()
        t0 = time.perf_counter()
        for _ in range(3):
            cpu_step()
        row["cpu_oracle_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        row["cpu_threads"] = torch.get_num_threads()
    except Exception as e:  # noqa: BLE001
        row["cpu_oracle_ms"] = repr(e)
    print(json.dumps(row), flush=True)

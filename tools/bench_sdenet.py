"""Config 5 (torch front end) training throughput: SDENet 3x32x32, blocks 2-2-2, hidden 32, w-net 1-128-1, sigma 0.1,
midpoint, dt 0.05 (20 steps, 40 drift evaluations of the 26-conv y_net), inhomogeneous=True.

    python tools/bench_sdenet.py [--batch 128] [--steps 3] [--warmup 2] [--math tc|fp32] [--remat] [--cpu_batch 4]

One JSON line: images/s (fwd + bwd + Adam, CUDA events), K2 algorithmic FLOPs (49.6 MFLOP per image per drift evaluation,
x 40 evaluations x 3 for fwd + dgrad + wgrad) and, beside it, the oracle's torch-CPU restatement of the same
step on the host cores (bounded sample)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesde_b200 import _lib, torch_models as tm  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--batch", type=int, default=128)
p.add_argument("--steps", type=int, default=3)
p.add_argument("--warmup", type=int, default=2)
p.add_argument("--math", default="tc")
p.add_argument("--remat", action="store_true")
p.add_argument("--hidden", type=int, default=32)
p.add_argument("--cpu_batch", type=int, default=4)
p.add_argument("--no_cpu", action="store_true")
a = p.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = tm.SDENet(input_size=(3, 32, 32), blocks=(2, 2, 2), weight_network_sizes=(1, 128, 1), hidden_width=a.hidden, sigma=0.1,
                  inhomogeneous=True, math=a.math, remat=a.remat).to(dev)
with torch.no_grad():   # N(0, 1e-2) on the zero-initialised last w-net layer, else the w drift is trivially -w
    last = model.w_net.linears()[-1]
    last.weight.normal_(0, 1e-2), last.bias.normal_(0, 1e-2)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
g = torch.Generator().manual_seed(1)
x = torch.randn(a.batch, 3, 32, 32, generator=g).to(dev)
y = torch.randint(0, 10, (a.batch,), generator=g).to(dev)
L = _lib.lib()


def step():
    opt.zero_grad(set_to_none=True)
    logits, logqp = model(x, dt=0.05, method="midpoint")
    loss = F.cross_entropy(logits, y) + 1e-4 * logqp     # examples/torch/sdebnn_classification.py:34-46
    loss.backward()
    opt.step()
    return loss


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
n0 = L.bsde_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
launches = (L.bsde_launch_count() - n0) // a.steps
flops_eval = sum(2 * (l["hout"] * l["wout"] if not l["transpose"] else l["hin"] * l["win"]) * 9 * (l["cin"] + 1) * l["cout"]
                 for l in model.y_net.layers)
alg = 3.0 * flops_eval * 40 * a.batch
line = {"workload": f"config5: SDENet 3x32x32, blocks 2-2-2, hidden {a.hidden}, w-net 1-128-1, midpoint dt 0.05, P={model.params_size}",
        "batch": a.batch, "math": a.math, "remat": a.remat, "ms_per_step": ms, "images_per_s": a.batch / (ms * 1e-3),
        "gpu_launches_per_step": int(launches), "loss": float(loss), "flops_per_image_eval": flops_eval,
        "conv_tflops": alg / (ms * 1e-3) / 1e12}
if not a.no_cpu:
    from oracle import torch_path as tp
    torch.set_num_threads(os.cpu_count() or 1)
    net = tp.SDENetOracle((3, 32, 32), (2, 2, 2), (1, 128, 1), sigma=0.1, hidden_width=a.hidden)
    gg = torch.Generator().manual_seed(0)
    w0 = tp.init_y_params(net.ops, net.P, gg).requires_grad_(True)
    wl = [(W.requires_grad_(True), b.requires_grad_(True)) for W, b in tp.init_w_net(net.P, net.hidden_sizes, gg, last_std=1e-2)]
    feat = 3 * 32 * 32
    proj = [((torch.randn(1024, feat, generator=gg) * 0.01).requires_grad_(True), torch.zeros(1024, requires_grad=True)),
            ((torch.randn(10, 1024, generator=gg) * 0.01).requires_grad_(True), torch.zeros(10, requires_grad=True))]
    leaves = [w0] + [t for pr in wl for t in pr] + [t for pr in proj for t in pr]
    copt = torch.optim.Adam(leaves, lr=1e-3)
    xc = torch.randn(a.cpu_batch, 3, 32, 32, generator=gg)
    yc = torch.randint(0, 10, (a.cpu_batch,), generator=gg)
    ts = []
    for it in range(2):
        t0 = time.perf_counter()
        I = torch.randn(20, net.P, generator=gg) * 0.05 ** 0.5
        copt.zero_grad()
        lo, lq = net.forward(xc, w0, wl, proj, I, dt=0.05)
        (F.cross_entropy(lo, yc) + 1e-4 * lq).backward()
        copt.step()
        ts.append(time.perf_counter() - t0)
    line["cpu_baseline"] = {"value": a.cpu_batch / ts[-1], "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": f"1 warm-up + 1 timed train step of oracle/torch_path.py at {a.cpu_batch} images, {ts[-1]:.2f} s"}
print(json.dumps(line))

"""Where is the GPU idle inside a training step?  Profiles a few steps of the bench workload with the CUPTI-backed torch profiler
(kernel start / end on every stream), merges the kernel intervals and reports: span, busy time (union over streams), idle
time, and the largest idle gaps with the kernels on either side.  Usage: python tools/step_gaps.py [--steps 2]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bayesde_b200.classification import SDEBNNClassifier, Trainer  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--steps", type=int, default=2)
p.add_argument("--batch", type=int, default=128)
p.add_argument("--nsamples", type=int, default=8)
a = p.parse_args()
dev = torch.device("cuda", 0)
mk = dict(nsamples=a.nsamples, sample_offset=0, total_samples=a.nsamples, diff_coef=0.2, time_grid="reference", remat=False,
          w_init=1e-2, b_init=1e-2, seed=0, device=dev, **bench.MODEL_KW)
model = SDEBNNClassifier(math="tc", **mk)
tr = Trainer(model, lr=1e-3, kl_coef=1e-3, train_size=45000)
g = torch.Generator().manual_seed(0)
x = torch.randn(a.batch, 32, 32, 3, generator=g).to(dev)
y = torch.randint(0, 10, (a.batch,), generator=g).to(dev)
for i in range(3):
    tr.step(x, y, seed=i)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(a.steps):
        tr.step(x, y, seed=10 + i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
span = iv[-1][1] - iv[0][0]
busy, gaps, cur_s, cur_e, last_name = 0.0, [], iv[0][0], iv[0][1], iv[0][2]
for s, e, n in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, last_name, n))
        cur_s, cur_e = s, e
        last_name = n
    elif e > cur_e:
        cur_e, last_name = e, n
busy += cur_e - cur_s
tot = sum(e - s for s, e, _ in iv)
print(f"steps {a.steps}: span {span / 1e3 / a.steps:.2f} ms/step, busy (union) {busy / 1e3 / a.steps:.2f}, idle {(span - busy) / 1e3 / a.steps:.2f}, "
      f"sum of kernel times {tot / 1e3 / a.steps:.2f} ({len(iv) // a.steps} kernels/step, {len(gaps) // a.steps} gaps/step)")
t0 = iv[0][0]
print("largest single gaps (ms since the first kernel, us):")
big = []
cur_e, last_name = iv[0][1], iv[0][2]
for s_, e_, n_ in iv[1:]:
    if s_ > cur_e:
        big.append((s_ - cur_e, (cur_e - t0) / 1e3, last_name, n_))
        cur_e, last_name = e_, n_
    elif e_ > cur_e:
        cur_e, last_name = e_, n_
for d, at, pn, nn in sorted(big, reverse=True)[:24]:
    print(f"  {d:7.1f} us at {at:8.2f} ms   {pn[:46]} -> {nn[:46]}")
import collections  # noqa: E402

agg = collections.defaultdict(lambda: [0, 0.0])
for d, pn, nn in gaps:
    k = (pn[:40], nn[:40])
    agg[k][0] += 1
    agg[k][1] += d
print("idle by (kernel before -> kernel after), per step:")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"  {d / 1e3 / a.steps:7.3f} ms  {c // a.steps:5d} x {d / c:7.1f} us   {k[0]} -> {k[1]}")

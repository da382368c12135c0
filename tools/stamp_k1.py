"""Per-phase timing of the K1 kernels (BSDE_K1_STAMP=1): %globaltimer stamps of CTA 0 at the phase boundaries of every scan
iteration of one forward and one reverse launch of a config-4 block.  Tools only."""
import ctypes as C
import os
import sys

os.environ["BSDE_K1_STAMP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bayesde_b200 import _lib, arch, brax

dev = "cuda:0"
S, B = 8, 4
for (H, Cc) in ((32, 5), (8, 80)):
    layer = brax.SDEBNN(0, 64, "softplus", arch.MLP((2, 64, 2), xt=True, ou_dw=True, nonzero_w=1e-2, nonzero_b=1e-2), diff_coef=0.2, xt=True,
                        nsteps=20, math="tc")
    _, params = layer.init(0, (B, H, H, Cc))
    params = brax.tree_map(lambda t: t.to(dev).requires_grad_(True), params)
    x = torch.randn(S, B, H, H, Cc, device=dev)
    for rep in range(2):
        xo, kl = layer.apply(params, x, 5)
        (xo.sum() + kl.sum()).backward()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (2 * 64 * 8))()
    L = _lib.lib()
    L.bsde_debug_k1_stamps(buf)
    st = np.frombuffer(buf, dtype=np.uint64).reshape(2, 64, 8).astype(np.int64)
    f, b = st[0, :19], st[1, :19]
    names_f = ["gather", "tiny_fwd", "main loop", "block_reduce", "grid.sync"]
    print(f"--- P = {params[0].numel()}  forward, ns per phase (median over iterations 1..17)")
    for i, n in enumerate(names_f):
        print(f"  {n:14s} {np.median((f[1:18, i + 1] - f[1:18, i])):8.0f}")
    print(f"  iteration      {np.median(f[2:18, 0] - f[1:17, 0]):8.0f}")
    names_b = ["tiny_fwd", "main loop", "block_reduce", "grid.sync", "gather", "tiny_bwd"]
    print("--- reverse")
    for i, n in enumerate(names_b):
        print(f"  {n:14s} {np.median((b[1:18, i + 1] - b[1:18, i])):8.0f}")
    print(f"  iteration      {np.median(b[1:17, 0] - b[2:18, 0]):8.0f}")

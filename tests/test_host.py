"""CPU tests: the C-ABI library loads and exports every symbol include/bsde.h declares; host logic
(layouts, time grid, pytrees, mean-field head, error behaviour) against the oracle; world_size-2 gloo
test of the gradient all-reduce."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import jax_path as jp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "bsde.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bsde_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = C.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), f"libbsde.so does not export {name}"
    nm = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bsde_[a-z0-9_]+)", nm))
    assert declared <= exported
    lib.bsde_version.restype = C.c_int
    assert lib.bsde_version() == 200


def test_struct_sizes_match_c(built_lib, tmp_path):
    """ctypes mirrors of the POD structs have the C compiler's sizes."""
    from bayesde_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "bsde.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(bsde_fw_params),'
                   'sizeof(bsde_wsde_desc), sizeof(bsde_conv_layer), sizeof(bsde_xpath_desc), sizeof(bsde_sdelayer_desc));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.FwParams), C.sizeof(_lib.WsdeDesc), C.sizeof(_lib.ConvLayer), C.sizeof(_lib.XpathDesc),
                     C.sizeof(_lib.SDELayerDesc)]


def test_no_cpu_fallback_and_error_mapping(built_lib):
    from bayesde_b200 import _lib, sdeint_ito_fixed_grid
    with pytest.raises(_lib.BsdeError):
        _lib.ptr(torch.zeros(3))
    with pytest.raises(ValueError, match="Unknown method"):
        sdeint_ito_fixed_grid(None, None, torch.zeros(3), torch.linspace(0, 1, 3), 0, method="heun")
    with pytest.raises(_lib.BsdeError):
        sdeint_ito_fixed_grid(lambda y, t, a: y, lambda y, t, a: y, torch.zeros(3), torch.linspace(0, 1, 3), 0)
    L = _lib.lib()
    st = L.bsde_step_update(7, 4, 0.1, None, None, None, None, None, None, None, None)
    assert st == -1 and b"Unknown method" in L.bsde_last_error()
    with pytest.raises(ValueError):
        _lib.check(st, "bsde_step_update")


def test_fx_spec_matches_oracle_layout():
    from bayesde_b200.arch import FxSpec
    for bt in (0, 1, 2):
        for C_in, H in ((5, 32), (20, 16), (80, 8)):
            fx = FxSpec(bt, (1, H, H, C_in), 64, "softplus")
            lay, P = jp.fx_layout(bt, C_in, 64)
            assert fx.P == P
            for a, b in zip(fx.layers, [e for e in lay if e["kind"] == "conv"]):
                assert (a["w_off"], a["b_off"], a["cin"], a["cout"], a["k"]) == (b["w_off"], b["b_off"], b["cin"], b["cout"], b["k"])
    with pytest.raises(AssertionError):  # quirk Q10: 7x7 does not survive the stride-2 / transpose pair
        FxSpec(0, (1, 7, 7, 4), 8, "softplus")
    with pytest.raises(ValueError):
        FxSpec(3, (1, 8, 8, 4), 8, "softplus")


def test_time_grid_matches_oracle():
    from bayesde_b200._fused import reference_time_grid
    for n in (2, 5, 20):
        ts = torch.linspace(0, 1, n)
        for mode in ("reference", "uniform"):
            tk, dtk = reference_time_grid(ts.numpy(), mode)
            tk_o, dtk_o = jp.time_grid(ts, mode)
            assert np.array_equal(tk, tk_o.numpy()) and np.array_equal(dtk, dtk_o.numpy())
    with pytest.raises(ValueError):
        reference_time_grid(np.linspace(0, 1, 3), "adaptive")


def test_parameter_free_layers_and_head_match_oracle():
    from bayesde_b200 import arch, brax
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 4, 4, 6, generator=g)
    assert torch.equal(arch.Augment(2).apply((), x), jp.augment(x, 2))
    assert torch.equal(arch.SqueezeDownsample(2).apply((), x), jp.squeeze_downsample(x))
    xs = torch.stack([x, 2 * x])
    assert torch.equal(arch.SqueezeDownsample(2).apply((), xs)[1], jp.squeeze_downsample(2 * x))
    # mean-field head, two MC samples with explicit noise (brax.py:138-168)
    head = brax.MeanField(arch.serial(arch.Flatten, arch.Dense(10), arch.LogSoftmax))
    _, params = head.init(1, (3, 4, 4, 6))
    (mean, logstd) = params
    eps = torch.randn(2, 96 * 10 + 10, generator=g)
    logp, kl = head.apply(params, xs, (0, eps), mc_samples=2)
    for s in range(2):
        lp_o, kl_o = jp.meanfield_dense_head((mean[1], logstd[1]), xs[s], eps[s])
        assert torch.allclose(logp[s], lp_o, atol=1e-5) and torch.allclose(kl[s], kl_o, rtol=1e-5)


def test_mlp_apply_quirk_q3_and_param_tree():
    from bayesde_b200 import arch
    fw = arch.MLP((2, 8, 2), xt=True, ou_dw=True, nonzero_w=0.05, nonzero_b=0.05)
    _, params = fw.init(3, (50,))
    assert params[0] == () and len(params[1]) == 7 and params[1][1] == ()
    w = torch.randn(50)
    t = torch.tensor(0.3)
    out, t2 = fw.apply(params, (w, t))
    ref = jp.fw_apply(arch.fw_layers(params), w, t)
    assert torch.allclose(out, ref) and torch.allclose(t2, 2 * t)  # Additive returns (mlp_out, 2t): OU term NOT added
    with pytest.raises(NotImplementedError):
        arch.MLP((2, 8, 2), xt=False)
    with pytest.raises(ValueError):
        arch.MLP((8, 8, 2), xt=True)


def test_swish_param_tree_and_layout_match_oracle():
    """Trainable swish (arch.py:14-29): f_w carries a (beta,) = 0.5 leaf after every hidden layer (arch.py:41-42), f_x's beta
    vectors live in the flat weight vector right after the conv they follow (ravel_pytree order, arch.py:172-188)."""
    from bayesde_b200 import arch
    fw = arch.MLP((2, 8, 2), actfn="swish", xt=True, ou_dw=True, nonzero_w=0.05, nonzero_b=0.05)
    _, params = fw.init(3, (50,))
    tree = params[1]
    assert len(tree) == 7 and [len(q) for q in tree] == [5, 1, 5, 1, 5, 1, 5]
    assert all(torch.equal(q[0], torch.full_like(q[0], 0.5)) for q in tree if len(q) == 1)
    assert [b.numel() for b in arch.fw_betas(params)] == [2, 8, 2] and len(arch.fw_layers(params)) == 4
    with torch.no_grad():
        for b in arch.fw_betas(params):
            b += 0.3 * torch.randn(b.shape)
    w, t = torch.randn(50), torch.tensor(0.3)
    out, _ = fw.apply(params, (w, t))
    lays, betas = arch.fw_layers(params), arch.fw_betas(params)
    ref = jp.fw_apply([lay + ((betas[i],) if i < len(betas) else ()) for i, lay in enumerate(lays)], w, t, "swish")
    assert torch.allclose(out, ref)
    for bt, cin in ((0, 5), (1, 20), (2, 80)):
        spec = arch.FxSpec(bt, (1, 8, 8, cin), 64, "swish")
        lay, P = jp.fx_layout(bt, cin, 64, "swish")
        assert spec.P == P
        convs, betas_o = [e for e in lay if e["kind"] == "conv"], [e for e in lay if e["kind"] == "beta"]
        assert [L["w_off"] for L in spec.layers] == [e["w_off"] for e in convs]
        assert [L["beta_off"] for L in spec.layers[:-1]] == [e["off"] for e in betas_o] and spec.layers[-1]["beta_off"] == -1
        w0 = spec.init(0)
        assert all(torch.equal(w0[e["off"]:e["off"] + 64], torch.full((64,), 0.5)) for e in betas_o)
    with pytest.raises(KeyError):
        arch.FxSpec(0, (1, 8, 8, 5), 64, "gelu")


def test_classifier_parameter_count_and_split_determinism():
    from bayesde_b200 import arch
    from bayesde_b200.classification import SDEBNNClassifier
    m = SDEBNNClassifier(device="cpu")
    assert sum(p.numel() for p in m.parameters()) == 6386760  # ~6.39 M
    assert arch.split(5, 3) == arch.split(5, 3) and len(set(arch.split(5, 3))) == 3


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from bayesde_b200.classification import Trainer
rank = int(os.environ["RANK"])
dist.init_process_group("gloo")
lin = torch.nn.Linear(4, 3)
with torch.no_grad():
    for p in lin.parameters():
        p.fill_(0.5)
tr = Trainer(lin.__class__(4, 3), lr=0.0) if False else None
class T(Trainer):
    def __init__(self, model):
        self.model = model
t = T(lin)
for i, p in enumerate(lin.parameters()):
    p.grad = torch.full_like(p, float(rank + 1 + i))
t._allreduce()
for i, p in enumerate(lin.parameters()):
    assert torch.allclose(p.grad, torch.full_like(p, 1.5 + i)), (rank, p.grad)
dist.destroy_process_group()
print("ok", rank)
"""


def test_gradient_allreduce_gloo_world2(tmp_path):
    """The only collective on the path: mean of parameter gradients over ranks."""
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), ROOT],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


_GLOO_SAMPLE_SHARD_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from bayesde_b200.classification import SDEBNNClassifier, Trainer, sample_shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
S = 4
kw = dict(input_size=(4, 4, 3), aug=1, nblocks=(), device="cpu", seed=3)   # Augment -> mean-field Dense head: runs on CPU
g = torch.Generator().manual_seed(0)
x, y = torch.randn(6, 4, 4, 3, generator=g), torch.randint(0, 10, (6,), generator=g)
# one process integrating all S samples
full = SDEBNNClassifier(nsamples=S, **kw)
tf = Trainer(full, lr=0.0, kl_coef=1.0, train_size=10)
lf, nf, kf = tf.step(x, y, seed=11)
ref = [p.grad.clone() for p in full.parameters()]
# this rank's shard of the samples, same batch; gradients meet in ONE all-reduce (mean): overlap on and off
off, per, bgrp, grp = sample_shard(S, world, rank)
assert (per, grp) == (S // world, 1)
for overlap in (False, True):
    m = SDEBNNClassifier(nsamples=per, sample_offset=off, total_samples=S, **kw)
    t = Trainer(m, lr=0.0, kl_coef=1.0, train_size=10, overlap=overlap)
    lo, nl, kl = t.step(x, y, seed=11)
    for p, r in zip(m.parameters(), ref):
        assert torch.allclose(p.grad, r, rtol=1e-5, atol=1e-7), (rank, overlap, (p.grad - r).abs().max())
    vals = torch.stack([lo, nl, kl])
    dist.all_reduce(vals)
    vals /= world                                   # every sample's KL is counted once: mean over ranks of the local means
    assert torch.allclose(vals, torch.stack([lf, nf, kf]), rtol=1e-5), (vals, lf, nf, kf)
# --acc_grad (script :220-240): chunk 0 twice, last chunk never, equal weights -- and the intended walk
m = SDEBNNClassifier(nsamples=S, **kw)
t = Trainer(m, lr=0.0, kl_coef=1.0, train_size=10, acc_grad=3)
assert [(c.start, c.stop) for c in t._chunks(6)] == [(0, 2), (0, 2), (2, 4)]
t.step(x, y, seed=11)
g_ref = [p.grad.clone() for p in m.parameters()]
acc = None
for sl in (slice(0, 2), slice(0, 2), slice(2, 4)):
    m.zero_grad()
    m.loss(x[sl], y[sl], 0.1, rng=11)[0].backward()
    acc = [p.grad / 3 for p in m.parameters()] if acc is None else [a + p.grad / 3 for a, p in zip(acc, m.parameters())]
for a, r in zip(acc, g_ref):
    assert torch.allclose(a, r, rtol=1e-5, atol=1e-6 * float(r.abs().max()))
assert [(c.start, c.stop) for c in Trainer(m, acc_grad=3, acc_grad_mode="intended")._chunks(6)] == [(0, 2), (2, 4), (4, 6)]
dist.destroy_process_group()
print("ok", rank)
"""


def test_sample_sharding_and_acc_grad_gloo_world2(tmp_path):
    """SURVEY 8.5 samples x batch grid: two ranks with S/2 samples each (global Philox sample counters, global head-noise rows)
    reproduce one rank with S samples after the single gradient all-reduce (serial and overlapped per-bucket launches); the
    KL of every sample is counted once.  Also the `--acc_grad` loop of the script (:220-240) with its chunk quirk."""
    script = tmp_path / "w2.py"
    script.write_text(_GLOO_SAMPLE_SHARD_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29618", str(script), ROOT],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_sass_carries_the_blackwell_instructions(built_lib):
    """Static evidence that the hot kernels are the sm_100a tensor-core / TMA code and not a recompiled legacy path: the library's SASS
    holds tcgen05.mma (UTCHMMA), tcgen05.ld (LDTM), tcgen05.commit (UTCBAR), TMA tensor loads (UTMALDG) and mbarrier waits (SYNCS), and
    the conv / wgrad kernels are compiled for sm_100a (profiles/r01_summary.md section 3)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True, check=True).stdout
    assert "arch = sm_100a" in sass
    for mnemonic, least in (("UTCHMMA", 50), ("LDTM", 10), ("UTCBAR", 20), ("UTMALDG", 20), ("SYNCS", 100)):
        assert sass.count(mnemonic) >= least, (mnemonic, sass.count(mnemonic))
    for kernel in ("conv_win_kernel", "wgrad_win_kernel", "wsde_fwd_kernel", "wsde_bwd_kernel", "brownian_tree_threefry_kernel"):
        assert kernel in sass, kernel
    assert "HMMA.16816" not in sass and "HGMMA" not in sass        # no mma.sync / wgmma tensor-core path in the library


def test_host_path_makes_no_per_call_device_copies():
    """The per-call host -> device copies of the step path are cached (a pageable copy stalls the host until the stream has
    drained: profiles/r02_summary.md section 8): the scalar branch of normal_logprob (brax.py:171-178) equals the tensor branch and
    the oracle, whole-solve time grids share one device copy per (grid, device), one-step grids (adjoint routes) are not cached."""
    from bayesde_b200 import _fused, arch, brax
    g = torch.Generator().manual_seed(0)
    z, mean = torch.randn(4, 7, generator=g), torch.randn(7, generator=g)
    ls = float(np.log(0.1))
    a = brax.normal_logprob(z, mean, ls)
    b = brax.normal_logprob(z, mean, torch.tensor(ls))
    r = jp.normal_logprob(z.double(), mean.double(), torch.tensor(ls, dtype=torch.float64))
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    assert torch.allclose(a.double(), r, rtol=1e-5, atol=1e-5)
    assert torch.allclose(brax.normal_logprob(z, 0.0, ls).double(), jp.normal_logprob(z.double(), 0.0, torch.tensor(ls, dtype=torch.float64)),
                          rtol=1e-5, atol=1e-5)

    fx = arch.FxSpec(0, (1, 8, 8, 5), 8, "softplus")
    fw = dict(hidden_dims=(2,), actfn="softplus", ou_dw=True)
    c1 = _fused.BlockConfig(fx, fw, 2, 2, 6, 0.2, False, "reference")
    c2 = _fused.BlockConfig(fx, fw, 2, 2, 6, 0.2, False, "reference")
    t1, t2 = c1.grid_on("cpu"), c2.grid_on("cpu")
    assert t1[0].data_ptr() == t2[0].data_ptr() and t1[1].data_ptr() == t2[1].data_ptr()
    assert np.array_equal(t1[0].numpy(), c1.tk) and np.array_equal(t1[1].numpy(), c1.dtk)
    c3 = _fused.BlockConfig(fx, fw, 2, 2, 6, 0.2, False, "uniform")
    if not (np.array_equal(c3.tk, c1.tk) and np.array_equal(c3.dtk, c1.dtk)):
        assert c3.grid_on("cpu")[1].data_ptr() != t1[1].data_ptr()
    n = len(_fused._GRID_CACHE)
    c1.set_step(0.25, 0.05)
    s1 = c1.grid_on("cpu")
    assert len(_fused._GRID_CACHE) == n and float(s1[0][0]) == 0.25 and abs(float(s1[1][0]) - 0.05) < 1e-7

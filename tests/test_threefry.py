"""Row N2: the reference's virtual Brownian tree on JAX's threefry PRNG (jaxsde/jaxsde/brownian.py:12-95 with
jax.random.split / jax.random.normal, jax==0.1.77).

JAX cannot be installed here, so the oracle (oracle/jax_threefry.py) restates the published algorithm and is pinned on known-answer
vectors: the Threefry-2x32-20 vectors of the Random123 distribution (the ones JAX's own `testThreefry2x32` uses) and the values the JAX
documentation prints for PRNGKey(0).  The CUDA path (bsde_brownian_tree_threefry through the C ABI) is then compared with the oracle:
the integer stream is exact, the float32 transform agrees to a few ulp (tolerance 1e-5 absolute on W(t), |W| = O(1))."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import jax_threefry as jt


# --------------------------------------------------------------------------------------------------------------
# CPU: known-answer pins of the oracle
# --------------------------------------------------------------------------------------------------------------
def test_threefry2x32_known_answers():
    """Random123 kat_vectors (threefry2x32, 20 rounds); also jax/tests/random_test.py::testThreefry2x32."""
    cases = [((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
             ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
             ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, want in cases:
        a, b = jt.threefry2x32(key[0], key[1], np.uint32([ctr[0]]), np.uint32([ctr[1]]))
        assert (int(a[0]), int(b[0])) == want


def test_jax_documented_values_for_key_zero():
    key = jt.PRNGKey(0)
    assert key.tolist() == [0, 0] and jt.PRNGKey((7 << 32) + 5).tolist() == [7, 5]
    assert jt.split(key).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]      # random.split(PRNGKey(0))
    assert abs(float(jt.uniform(key, 1)[0]) - 0.41845703) < 5e-9                                # random.uniform(key, (1,))
    assert abs(float(jt.normal(key, 1)[0]) - (-0.20584226)) < 5e-9                              # random.normal(key, (1,))
    want = np.float32([1.8160863, -0.48262316, 0.33988908])                                     # random.normal(key, (3,))
    assert np.allclose(jt.normal(key, 3), want, rtol=0, atol=2e-7)


def test_random_bits_lane_layout():
    """counts [0, half) go to lane 0, [half, 2 half) to lane 1; an odd count is padded with one zero and the last word dropped."""
    key = jt.PRNGKey(42)
    even, odd = jt.random_bits(key, 6), jt.random_bits(key, 5)
    y0, y1 = jt.threefry2x32(key[0], key[1], np.uint32([0, 1, 2]), np.uint32([3, 4, 5]))
    assert even.tolist() == y0.tolist() + y1.tolist()
    y0, y1 = jt.threefry2x32(key[0], key[1], np.uint32([0, 1, 2]), np.uint32([3, 4, 0]))
    assert odd.tolist() == (y0.tolist() + y1.tolist())[:5]


def test_erfinv_against_scipy():
    from scipy.special import erfinv
    x = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
    got, want = jt.erfinv_f32(x).astype(np.float64), erfinv(x.astype(np.float64))
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) < 5e-6      # Giles' single-precision fit: 8e-7 for |x| < 0.999, 3e-6 in the tails
    assert np.isinf(jt.erfinv_f32(np.float32([1.0, -1.0]))).all()


def test_normal_law():
    z = jt.normal(jt.PRNGKey(3), 200001)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1.0) < 0.01 and abs(np.mean(z ** 4) - 3.0) < 0.1
    assert np.isfinite(z).all()


def test_tree_structure_and_laws():
    n, rep = 4001, 1000
    bm = jt.ThreefryBrownianTree(jt.PRNGKey(9), n, 0.5, 2.5, rep=rep)
    assert np.abs(bm(0.5)).max() == 0.0
    assert np.allclose(bm(2.5), np.sqrt(np.float32(2.0)) * bm.w_end, rtol=1e-6, atol=1e-6)
    w = bm(1.7)
    assert np.array_equal(w[n - rep:], w[n - 2 * rep:n - rep]) and np.array_equal(w, bm(1.7))
    big = jt.ThreefryBrownianTree(jt.PRNGKey(10), 40000, 0.0, 1.0)
    a, b, c = big(0.2), big(0.55), big(0.9)
    d1, d2 = b - a, c - b
    assert abs(d1.var() - 0.35) < 0.01 and abs(d2.var() - 0.35) < 0.01 and abs(np.mean(d1 * d2)) < 0.01 and abs(np.mean(a * d1)) < 0.01


def test_product_host_split_matches_oracle(built_lib):
    import bayesde_b200 as bs
    for seed in (0, 1, 2 ** 40 + 17):
        key = bs.PRNGKey(seed)
        assert key.tolist() == jt.PRNGKey(seed).tolist()
        for num in (1, 2, 3, 8):
            assert np.array_equal(bs.split(key, num), jt.split(key, num))
    with pytest.raises(ValueError):
        bs.split(np.zeros(3, dtype=np.uint32))


# --------------------------------------------------------------------------------------------------------------
# GPU
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,rep,depth", [(1, 0, 10), (2, 1, 10), (1000, 0, 10), (4099, 1024, 10), (301, 150, 0), (257, 0, 17)])
def test_gpu_threefry_tree_matches_oracle(built_lib, n, rep, depth):
    from bayesde_b200.sdeint import BrownianTree, PRNGKey
    key = PRNGKey(20201025)
    t0, t1 = -0.5, 1.25
    dev = BrownianTree(t0, n, t1, key, depth=depth, rep=rep, device="cuda:0")
    ref = jt.ThreefryBrownianTree(key, n, t0, t1, depth=depth, rep=rep)
    for t in (t0, t1, 0.0, 0.3333, 1.2499, -0.4999, 0.375):
        a, b = dev(t).cpu().numpy(), ref(t)
        assert np.abs(a - b).max() <= 1e-5, (t, np.abs(a - b).max())
    lo, cnt = n // 3, max(n // 2, 1)
    cnt = min(cnt, n - lo)
    assert torch.equal(dev.query(0.77, lo, cnt), dev(0.77)[lo:lo + cnt])          # slices draw the same entries


@pytest.mark.gpu
def test_gpu_threefry_normals_exact_bits(built_lib):
    """depth 0 at t = t1 returns sqrt(span) * normal(end_key, (n,)): the device normals against the oracle's, element by element,
    over both lanes and the odd-length padding (<= 4 ulp: log1p / sqrt rounding)."""
    from bayesde_b200.sdeint import BrownianTree, PRNGKey
    for n in (7, 8, 100001):
        key = PRNGKey(5)
        dev = BrownianTree(0.0, n, 1.0, key, depth=0, device="cuda:0")(1.0).cpu().numpy()
        ref = jt.normal(jt.split(key)[0], n)
        assert np.max(np.abs(dev - ref) / np.maximum(np.abs(ref), 1e-2)) < 5e-6


@pytest.mark.gpu
def test_gpu_fixed_grid_with_jax_key(built_lib):
    """sdeint_ito_fixed_grid(..., rng=PRNGKey) draws bm(next_t) - bm(curr_t) from the reference's tree with the scan's own time
    bookkeeping (quirk Q1): against the oracle's generic solver fed with the oracle tree's increments."""
    import bayesde_b200 as bs
    from oracle import jax_path as jp
    D, rep = 12, 4
    key = bs.PRNGKey(3)
    ts = torch.linspace(0.0, 1.0, 7)
    f = lambda y, t, a: -a[0] * y + torch.sin(t)
    g = lambda y, t, a: a[1] * torch.cos(y)
    for tg in ("reference", "uniform"):
        for method in ("euler_maruyama", "milstein"):
            a_d = torch.tensor([0.7, 0.3], device="cuda:0")
            y0 = torch.linspace(-1.0, 1.0, D)
            ys = bs.sdeint_ito_fixed_grid(f, g, y0.to("cuda:0"), ts, key, a_d, method=method, rep=rep, time_grid=tg)
            tree = jt.ThreefryBrownianTree(key, D, 0.0, 1.0, rep=rep)
            tsn = ts.numpy().astype(np.float32)
            rows, curr = [], tsn[0]
            for k in range(1, len(tsn)):
                nxt = np.float32(tsn[k] - curr) if tg == "reference" else tsn[k]
                rows.append(tree(nxt) - tree(curr))
                curr = nxt
            inc = torch.from_numpy(np.stack(rows)).double()
            ys_r = jp.sdeint_ito_fixed_grid(f, g, y0.double(), ts.double(), inc, a_d.cpu().double(), method=method, time_grid_mode=tg)
            assert torch.allclose(ys.cpu().double(), ys_r, rtol=1e-4, atol=1e-5), (tg, method)


@pytest.mark.gpu
def test_gpu_sdebnn_block_with_jax_keys(built_lib):
    """One key per MC sample (vmap over rngs): the block's w rows use entries [x_dim, x_dim + P) of each sample's D-entry tree; the
    result equals the run with those increments supplied explicitly."""
    import bayesde_b200 as bs
    from bayesde_b200 import arch, brax
    S, B, H, Cc, fx_dim, nsteps = 2, 2, 8, 5, 8, 5
    fw_layer = arch.MLP((2, 8, 2), actfn="softplus", xt=True, ou_dw=True)
    layer = brax.SDEBNN(0, fx_dim, "softplus", fw_layer, diff_coef=0.2, stl=True, xt=True, nsteps=nsteps)
    _, params = layer.init(1, (B, H, H, Cc), device="cuda:0")
    x = torch.randn(S, B, H, H, Cc, generator=torch.Generator().manual_seed(0)).to("cuda:0")
    keys = bs.split(bs.PRNGKey(7), S)
    xo, kl, info = layer.apply(params, x, keys, full_output=True)
    P = params[0].numel()
    x_dim = B * H * H * Cc
    incs = []
    for s in range(S):
        tree = jt.ThreefryBrownianTree(keys[s], x_dim + 2 * P, 0.0, 1.0, rep=P)
        tsn = np.linspace(0.0, 1.0, nsteps, dtype=np.float32)
        rows, curr = [], tsn[0]
        for k in range(1, nsteps):
            nxt = np.float32(tsn[k] - curr)
            rows.append((tree(nxt) - tree(curr))[x_dim:x_dim + P])
            curr = nxt
        incs.append(np.stack(rows))
    dW = torch.from_numpy(np.stack(incs)).to("cuda:0")
    xo2, kl2, info2 = layer.apply(params, x, dW, full_output=True)
    assert torch.allclose(xo, xo2, rtol=1e-4, atol=1e-5) and torch.allclose(kl, kl2, rtol=1e-4, atol=1e-4)
    assert torch.allclose(info["sdebnn_w"], info2["sdebnn_w"], rtol=1e-5, atol=2e-6)


def test_tree_increments_follow_the_scan_bookkeeping():
    """Host logic of `tree_increments`: with time_grid="reference" the queries are bm(next_t) - bm(curr_t) with next_t = target - curr
    (quirk Q1: the scan's own bookkeeping, jaxsde/jaxsde/sdeint.py:103-110), with "uniform" the intended ts[k] - ts[k-1] pairs."""
    from bayesde_b200.sdeint import tree_increments

    class FakeTree:
        def __init__(self):
            self.calls = []

        def query(self, t, first=0, count=None):
            self.calls.append((float(np.float32(t)), first, count))
            return torch.full((3,), float(t))

    ts = np.linspace(0.0, 1.0, 5, dtype=np.float32)
    bm = FakeTree()
    inc = tree_increments(bm, ts, "reference", 2, 3)
    f32 = np.float32
    want, curr = [], ts[0]
    for k in range(1, 5):
        nxt = f32(ts[k] - curr)
        want.append((float(nxt), float(curr)))
        curr = nxt
    got = [(bm.calls[2 * i][0], bm.calls[2 * i + 1][0]) for i in range(4)]
    assert got == want and all(c[1:] == (2, 3) for c in bm.calls)
    assert inc.shape == (4, 3) and torch.allclose(inc[:, 0], torch.tensor([a - b for a, b in want]))
    assert [w[0] for w in want] == [0.25, 0.25, 0.5, 0.5]            # dt = 0.25, 0, 0.25, 0: the zero steps of quirk Q1
    bm2 = FakeTree()
    inc2 = tree_increments(bm2, torch.from_numpy(ts), "uniform")
    assert [(bm2.calls[2 * i][0], bm2.calls[2 * i + 1][0]) for i in range(4)] == [(0.25, 0.0), (0.5, 0.25), (0.75, 0.5), (1.0, 0.75)]
    assert torch.allclose(inc2, torch.full((4, 3), 0.25))

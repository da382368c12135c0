"""CPU oracle for SURVEY 8.6 row N2: the reference's virtual Brownian tree driven by JAX's threefry PRNG.  TEST INFRASTRUCTURE ONLY.

`jaxsde/jaxsde/brownian.py:12-95` draws every bridge point with `jax.random.normal(key, shape)` and walks the tree with
`jax.random.split(key)`.  JAX is a third-party dependency that is absent here (`requirements.txt:1-2`: jax==0.1.77, jaxlib>=0.1.60;
no wheel, no network), so its published algorithm is restated in numpy:
  * Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; jax/_src/prng.py / jax/random.py
    `threefry_2x32`): rotations (13, 15, 26, 6) / (17, 29, 16, 24), key schedule with 0x1BD11BDA;
  * `PRNGKey(seed)` = [seed >> 32, seed & 0xFFFFFFFF]; `split(key, n)` = threefry of iota(2n) reshaped (n, 2); `_random_bits`:
    counts = iota(size) padded to an even length, first half -> lane 0, second half -> lane 1, outputs concatenated;
  * `uniform`: mantissa bits `bits >> 9 | 0x3F800000` - 1, scaled to [minval, maxval); `normal` = sqrt(2) * erf_inv(uniform(-1+ulp, 1));
  * XLA's float32 `erf_inv` (Giles, "Approximating the erfinv function", 2010: two degree-8 polynomials in w = -log1p(-x*x)).

PARITY STATUS: pinned on known-answer vectors only (tests/test_threefry.py): the three Threefry-2x32 vectors of the Random123
distribution (also JAX's own `testThreefry2x32`), `split(PRNGKey(0))` = [[4146024105, 967050713], [2718843009, 1272950319]],
`uniform(PRNGKey(0), (1,))` = 0.41845703 and `normal(PRNGKey(0), (1,))` = -0.20584226 as printed in the JAX documentation.  No JAX
run could be made in this container, so the tree built on top of them is otherwise **parity unpinned**.
"""
import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << _U32(r)) | (x >> _U32(32 - r))).astype(_U32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 of the counter lanes (x0, x1) under key (k0, k1); all uint32 arrays / scalars."""
    with np.errstate(over="ignore"):
        k0, k1 = _U32(k0), _U32(k1)
        x0 = np.asarray(x0, dtype=_U32).copy()
        x1 = np.asarray(x1, dtype=_U32).copy()
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 = (x0 + ks[0]).astype(_U32)
        x1 = (x1 + ks[1]).astype(_U32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r) ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(i + 2) % 3] + _U32(i + 1)).astype(_U32)
    return x0, x1


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def _threefry_counts(key, counts):
    """jax.random.threefry_2x32(keypair, count): count is split in two halves (padded with one zero when odd)."""
    counts = np.asarray(counts, dtype=_U32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, dtype=_U32)])
    half = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num=2):
    return _threefry_counts(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits(key, n):
    return _threefry_counts(key, np.arange(n, dtype=_U32))


def erfinv_f32(x):
    """XLA's float32 erf_inv (Giles 2010, single precision)."""
    f = np.float32
    x = np.asarray(x, dtype=f)
    lt5 = (f(2.81022636e-08), f(3.43273939e-07), f(-3.5233877e-06), f(-4.39150654e-06), f(0.00021858087), f(-0.00125372503),
           f(-0.00417768164), f(0.246640727), f(1.50140941))
    ge5 = (f(-0.000200214257), f(0.000100950558), f(0.00134934322), f(-0.00367342844), f(0.00573950773), f(-0.0076224613),
           f(0.00943887047), f(1.00167406), f(2.83297682))
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-np.log1p((-x * x).astype(f))).astype(f)
        lt = w < f(5.0)
        ww = np.where(lt, w - f(2.5), np.sqrt(w) - f(3.0)).astype(f)
        p = np.where(lt, lt5[0], ge5[0]).astype(f)
        for a, b in zip(lt5[1:], ge5[1:]):
            p = (np.where(lt, a, b).astype(f) + p * ww).astype(f)
        out = (p * x).astype(f)
        return np.where(np.abs(x) == f(1.0), x * f(np.inf), out).astype(f)


def uniform(key, n, minval=0.0, maxval=1.0):
    f = np.float32
    bits = random_bits(key, n)
    floats = ((bits >> _U32(9)) | _U32(0x3F800000)).view(f) - f(1.0)
    minval, maxval = f(minval), f(maxval)
    return np.maximum(minval, (floats * (maxval - minval) + minval).astype(f)).astype(f)


def normal(key, n):
    f = np.float32
    lo = np.nextafter(f(-1.0), f(0.0), dtype=f)
    return (f(np.sqrt(2.0)) * erfinv_f32(uniform(key, n, lo, 1.0))).astype(f)


def sample_with_rep(key, n, rep):
    """jaxsde/jaxsde/brownian.py:69-74."""
    z = normal(key, n)
    if rep:
        z[n - rep:] = z[n - 2 * rep:n - rep]
    return z


class ThreefryBrownianTree:
    """make_brownian_motion(t0, zeros(n), t1, rng, depth, rep) (jaxsde/jaxsde/brownian.py:77-95) in float32, rng = a JAX key."""

    def __init__(self, key, n, t0, t1, depth=10, rep=0):
        self.n, self.t0, self.t1, self.depth, self.rep = int(n), np.float32(t0), np.float32(t1), int(depth), int(rep)
        end_key, self.path_key = split(np.asarray(key, dtype=_U32))           # brownian.py:81
        self.w_end = sample_with_rep(end_key, self.n, self.rep)             # :82

    def _bridge(self, t0, x0, t1, x1, t, key):                              # brownian_bridge_point, :19-23
        f = np.float32
        mean = ((x0 * f(t1 - t) + x1 * f(t - t0)) / f(t1 - t0)).astype(f)
        std = np.sqrt(f(f(t1 - t) * f(t - t0)) / f(t1 - t0)).astype(f)
        return (mean + std * sample_with_rep(key, self.n, self.rep)).astype(f)

    def standard(self, s):                                                  # virtual_brownian_tree, :26-51
        f = np.float32
        s = f(s)
        t0, x0, t1, x1, key = f(0.0), np.zeros(self.n, dtype=f), f(1.0), self.w_end, self.path_key
        for _ in range(self.depth):
            tm = f(f(t0 + t1) / f(2.0))
            xm = self._bridge(t0, x0, t1, x1, tm, key)
            left, right = split(key)
            if s < tm:
                t1, x1, key = tm, xm, left
            else:
                t0, x0, key = tm, xm, right
        return self._bridge(t0, x0, t1, x1, s, key)

    def __call__(self, t):                                                  # scaled_brownian_motion, :12-17
        f = np.float32
        span = f(self.t1 - self.t0)
        return (np.sqrt(span) * self.standard(f(f(f(t) - self.t0) / span))).astype(f)

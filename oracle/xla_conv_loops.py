"""Explicit-loop numpy restatement of XLA's ConvGeneralDilated (NHWC/HWIO), used only to
pin ``oracle/jax_path.conv_general_dilated_nhwc`` on tiny shapes (TEST INFRASTRUCTURE).

Definition followed (XLA operation semantics, "ConvWithGeneralPadding / dilation" [3P]):
  1. lhs is dilated: d-1 zeros between neighbouring spatial entries;
  2. then padded with (lo, hi) zeros per spatial dim;
  3. out[n,oy,ox,co] = sum_{kh,kw,ci} lhs'[n, oy*s+kh, ox*s+kw, ci] * rhs[kh,kw,ci,co]
     (cross-correlation: the kernel is not flipped).
Call sites in the reference: brax/_impl/diffeq_layers.py:127-135 via stax.GeneralConv /
stax.GeneralConvTranspose -> lax.conv_general_dilated / lax.conv_transpose.
"""
import numpy as np


def conv_general_dilated_loops(x, w, stride, pads, lhs_dilation=1):
    n, h, wd, ci = x.shape
    kh, kw, ci2, co = w.shape
    assert ci == ci2
    d = lhs_dilation
    hd, wdd = (h - 1) * d + 1, (wd - 1) * d + 1
    (pt, pb), (pl, pr) = pads
    xp = np.zeros((n, hd + pt + pb, wdd + pl + pr, ci), dtype=np.float64)
    for i in range(h):
        for j in range(wd):
            xp[:, pt + i * d, pl + j * d, :] = x[:, i, j, :]
    ho = (xp.shape[1] - kh) // stride + 1
    wo = (xp.shape[2] - kw) // stride + 1
    out = np.zeros((n, ho, wo, co), dtype=np.float64)
    for oy in range(ho):
        for ox in range(wo):
            for a in range(kh):
                for b in range(kw):
                    out[:, oy, ox, :] += xp[:, oy * stride + a, ox * stride + b, :] @ w[a, b]
    return out

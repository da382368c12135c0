"""Index-math prototype (numpy, CPU) for DESIGN.md section 8 item 2: "taps on the N side" for the small-cout layers (64 -> 5, 64 -> 20).

Today `conv_win_kernel` issues nine MMAs per k-step for a 3x3 layer -- one per tap, each reading the same 128-row A operand through a
row-shifted descriptor -- so with cout = 5 (N padded to 16) the tensor-core shared-memory port carries the A tile nine times (ncu:
22.6 M wavefronts, 81 of the launch's 168-173 us).  The alternative evaluated here computes, per tile of 128 INPUT positions p of the
flattened padded grid, ONE product per k-step with all nine taps on the N side,

        T[p][tap][co] = sum_ci in[p][ci] * w[tap][ci][co]                       (M = 128, K = cin + 1, N = 9 * cout = 45 -> 48)

and moves the tap shift into the epilogue:   out[q][co] = sum_tap T[q + s_tap][tap][co],   s_tap = dy * pitch + dx.

A persistent CTA walks consecutive input tiles and scatter-adds each tile's T into a rolling accumulation window over output positions;
after the tile that starts at p0 every output q < p0 + 128 - (pitch + 1) is complete and can be flushed (bias / activation / store), so the
window needs 128 + 2 (pitch + 1) rows of cout floats (198 x 5 floats at 32 x 32) and every CTA processes ceil((pitch + 1) / 128) halo tiles
on each side of the output range it owns.  This script checks that schedule against the direct convolution on the same padded grid
(pad column at x = W, pad row at y = H, zero filled, as the TMA window of conv_win_kernel produces them) and prints the operand traffic of
both forms.  Run: python tools/proto_taps_on_n.py"""
import numpy as np


def padded_grid(x):
    """x [N, H, W, C] -> flattened padded positions [N * (H + 1) * (W + 1), C] with a zero pad column / row (and its geometry)."""
    n, h, w, c = x.shape
    g = np.zeros((n, h + 1, w + 1, c), dtype=x.dtype)
    g[:, :h, :w] = x
    return g.reshape(-1, c), w + 1, (h + 1) * (w + 1)


def direct(xp, wgt, pitch):
    """out[q][co] = sum_tap sum_ci xp[q + s_tap][ci] wgt[tap][ci][co] on the padded grid (out-of-range rows are zero)."""
    q, cin = xp.shape
    out = np.zeros((q, wgt.shape[2]), dtype=np.float64)
    for tap in range(9):
        s = (tap // 3 - 1) * pitch + (tap % 3 - 1)
        lo, hi = max(0, -s), min(q, q - s)
        out[lo:hi] += xp[lo + s:hi + s].astype(np.float64) @ wgt[tap].astype(np.float64)
    return out


def taps_on_n(xp, wgt, pitch, q_lo, q_hi, tile=128):
    """The rolling-window schedule for one CTA that owns the outputs [q_lo, q_hi): returns (outputs, tiles processed)."""
    qn, cin = xp.shape
    cout = wgt.shape[2]
    halo = pitch + 1
    wn = np.concatenate([wgt[t] for t in range(9)], axis=1).astype(np.float64)          # [cin, 9 * cout]: taps on N
    shifts = [(t // 3 - 1) * pitch + (t % 3 - 1) for t in range(9)]
    out = np.zeros((q_hi - q_lo, cout))
    p_first = (max(q_lo - halo, 0) // tile) * tile
    p_last = min(q_hi + halo, qn)
    win_lo = p_first - halo                                                             # rolling window [win_lo, win_lo + tile + 2 halo)
    win = np.zeros((tile + 2 * halo, cout))
    ntiles = 0
    for p0 in range(p_first, p_last, tile):
        rows = min(tile, qn - p0)
        t_acc = xp[p0:p0 + rows].astype(np.float64) @ wn                                # ONE product per k-step: the MMA
        ntiles += 1
        for t, s in enumerate(shifts):                                                  # epilogue: scatter-add with the tap shift
            for r in range(rows):
                q = p0 + r - s
                assert win_lo <= q < win_lo + win.shape[0]
                win[q - win_lo] += t_acc[r, t * cout:(t + 1) * cout]
        # every output below p0 + tile - halo is complete: flush `tile` rows and slide the window
        for q in range(win_lo, win_lo + tile):
            if q_lo <= q < q_hi:
                out[q - q_lo] = win[q - win_lo]
        win = np.concatenate([win[tile:], np.zeros((tile, cout))])
        win_lo += tile
    for q in range(win_lo, win_lo + win.shape[0]):                                      # drain
        if q_lo <= q < q_hi:
            out[q - q_lo] = win[q - win_lo]
    return out, ntiles


def main():
    rng = np.random.default_rng(0)
    for (n, h, w, cin, cout) in ((3, 8, 8, 9, 5), (2, 16, 16, 17, 20), (1, 32, 32, 65, 5)):
        x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
        x[..., -1] = 0.37                                                               # the time channel (constant inside the image)
        wgt = rng.standard_normal((9, cin, cout)).astype(np.float32)
        xp, pitch, _ = padded_grid(x)
        ref = direct(xp, wgt, pitch)
        qn = xp.shape[0]
        cuts = [0, qn // 3 + 5, 2 * qn // 3 - 7, qn]                                     # three CTAs with ragged output ranges
        got = np.concatenate([taps_on_n(xp, wgt, pitch, a, b)[0] for a, b in zip(cuts, cuts[1:])])
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12), (n, h, w)
        # cross-check the padded-grid convolution itself against a plain SAME convolution on the image
        xi = np.pad(x.astype(np.float64), ((0, 0), (1, 1), (1, 1), (0, 0)))
        same = sum(xi[:, dy:dy + h, dx:dx + w] @ wgt[dy * 3 + dx].astype(np.float64) for dy in range(3) for dx in range(3))
        assert np.allclose(ref.reshape(n, h + 1, w + 1, cout)[:, :h, :w], same, rtol=1e-12, atol=1e-12)
        k = cin
        npad = 16 if cout <= 16 else 32
        a_now, b_now = 9 * 128 * k * 4, 9 * npad * k * 4                                 # bytes through the smem port per tile today
        a_new, b_new = 128 * k * 4, ((9 * cout + 15) // 16 * 16) * k * 4
        print(f"{n}x{h}x{w}x{cin}->{cout}: ok | operand bytes per 128-position tile: {a_now + b_now} (9 taps x N={npad}) -> "
              f"{a_new + b_new} (N={(9 * cout + 15) // 16 * 16}), {(a_now + b_now) / (a_new + b_new):.1f}x fewer; "
              f"window {128 + 2 * (pitch + 1)} rows x {cout} floats, {(pitch + 1 + 127) // 128} halo tile(s) per side")


if __name__ == "__main__":
    main()

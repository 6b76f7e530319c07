"""CPU restatement of scikit-image 0.18.3 `transform.warp(image, tf.inverse, order=3)` for a pure
translation — TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline leg and the reference import shim; never by the product path).

PARITY UNPINNED at this boundary: scikit-image is a third-party dependency of the reference
(pinned `scikit_image == 0.18.3`, /root/reference/drl4ao/AO_OOPAO/requirements.txt:19-20) that is not
vendored under /root/reference and not installable offline, and the reference ships no golden vector
for it. What follows restates the published 0.18.3 algorithm; call sites it serves:
OOPAO/tools/tools.py:210-217 (translationImageMatrix / globalTransformation), used by
OOPAO/Atmosphere.py:303-305 (integer shift in add_row) and :406-407 (sub-pixel shift every step).

Published algorithm (skimage/transform/_warps.py `warp`, `_clip_warp_output`;
skimage/transform/_warps_cy.pyx `_warp_fast`; skimage/_shared/interpolation.pxd, all at tag v0.18.3):

* `SimilarityTransform(translation=(tx, ty)).inverse` is the homography [[1,0,-tx],[0,1,-ty],[0,0,1]];
  warp() sees a homography with order in (0,1,3) and takes the Cython fast path, whose "metric" branch maps
  output pixel (row, col) to input coordinates r = row - ty, c = col - tx (x is the column axis).
* order=3 -> `bicubic_interpolation`: r0 = (long)r - 1, c0 = (long)c - 1, each decremented once more when
  the coordinate is negative (C truncation -> floor); the fractional position is rescaled to the span of the
  four taps, xr = (r - r0)/3, xc = (c - c0)/3; taps outside the image read `cval` (=0, mode='constant').
  For each of the 4 tap rows the 4 tap columns are combined by `cubic_interpolation(xc, .)`, then the 4 row
  results by `cubic_interpolation(xr, .)`.
* `cubic_interpolation(x, f)` is the cubic through f[0..3] placed at x = 0, 1/3, 2/3, 1, in Horner form
  f0 + x(-5.5f0 + 9f1 - 4.5f2 + f3 + x(9f0 - 22.5f1 + 18f2 - 4.5f3 + x(-4.5f0 + 13.5f1 - 13.5f2 + 4.5f3))).
  (scikit-image >= 0.19 replaced this by a Catmull-Rom spline; `kernel="catmull_rom"` gives that variant.)
* clip=True (default) then clamps the output to [image.min(), image.max()]; when cval lies outside that
  range, output pixels exactly equal to cval are restored after clamping.
"""
import numpy as np


def cubic_weights(x, kernel="lagrange018"):
    """Weights w[0..3] such that cubic_interpolation(x, f) == sum_k w[k] f[k].

    lagrange018: x is the rescaled position in [1/3, 2/3); catmull_rom: x is the plain fraction in [0, 1).
    """
    x = np.asarray(x, dtype=np.float64)
    if kernel == "lagrange018":
        w0 = 1.0 + x * (-5.5 + x * (9.0 + x * -4.5))
        w1 = x * (9.0 + x * (-22.5 + x * 13.5))
        w2 = x * (-4.5 + x * (18.0 + x * -13.5))
        w3 = x * (1.0 + x * (-4.5 + x * 4.5))
    elif kernel == "catmull_rom":
        w0 = 0.5 * x * (-1.0 + x * (2.0 - x))
        w1 = 1.0 + 0.5 * x * x * (-5.0 + 3.0 * x)
        w2 = 0.5 * x * (1.0 + x * (4.0 - 3.0 * x))
        w3 = 0.5 * x * x * (x - 1.0)
    else:
        raise ValueError(kernel)
    return np.stack([w0, w1, w2, w3])


def _cubic(x, f0, f1, f2, f3, kernel):
    if kernel == "lagrange018":
        return f0 + x * (-5.5 * f0 + 9.0 * f1 - 4.5 * f2 + f3
                         + x * (9.0 * f0 - 22.5 * f1 + 18.0 * f2 - 4.5 * f3
                                + x * (-4.5 * f0 + 13.5 * f1 - 13.5 * f2 + 4.5 * f3)))
    return f1 + 0.5 * x * (f2 - f0 + x * (2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3
                                          + x * (3.0 * (f1 - f2) + f3 - f0)))


def tap_origin_and_frac(coord, kernel="lagrange018"):
    """First tap index and interpolation argument for input coordinate(s) `coord`."""
    coord = np.asarray(coord, dtype=np.float64)
    base = np.trunc(coord).astype(np.int64) - 1
    base = np.where(coord < 0, base - 1, base)
    if kernel == "lagrange018":
        x = (coord - base) / 3.0
    else:
        # >= 0.19: taps start at floor(coord) - 1, argument is the plain fraction
        base = np.floor(coord).astype(np.int64) - 1
        x = coord - (base + 1)
    return base, x


def warp_translate(image, tx, ty, kernel="lagrange018", clip=True, cval=0.0):
    """`warp(image, SimilarityTransform(translation=(tx, ty)).inverse, order=3)` for a 2-D float image."""
    img = np.asarray(image, dtype=np.float64)
    rows, cols = img.shape
    r0, xr = tap_origin_and_frac(np.arange(rows) - ty, kernel)      # per output row
    c0, xc = tap_origin_and_frac(np.arange(cols) - tx, kernel)      # per output column
    pad = np.full((rows + 8, cols + 8), cval)
    pad[4:-4, 4:-4] = img
    r0 = np.clip(r0, -4, rows) + 4       # indices into the padded image; taps beyond the pad are cval anyway
    c0 = np.clip(c0, -4, cols) + 4
    fr = []
    for pr in range(4):
        rowsel = pad[r0 + pr, :]                                    # [rows_out, cols+8]
        fc = [rowsel[:, c0 + pc] for pc in range(4)]                # each [rows_out, cols_out]
        fr.append(_cubic(xc[None, :], fc[0], fc[1], fc[2], fc[3], kernel))
    out = _cubic(xr[:, None], fr[0], fr[1], fr[2], fr[3], kernel)
    if clip:
        lo, hi = img.min(), img.max()
        keep = None
        if not (lo <= cval <= hi):
            keep = out == cval
        out = np.clip(out, lo, hi)
        if keep is not None:
            out[keep] = cval
    return out

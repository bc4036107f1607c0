#!/usr/bin/env python3
"""CPU simulation (numpy, no GPU) of the arithmetic of img_interpolate_logpolar against the
reference's typing (image_sampler_interpolate_kernel.cl:28-44), on EVERY pixel of a frame:

* i_f: host table {1/c, K ln c} per (exponent, top 7 mantissa bits) of d2 = dx^2 + dy^2 and a cubic
  in r = d2/c - 1, in double, rounded to float  vs  (float)(ow * (log(sqrt(d2)) / 10));
* j_f: octant-reduced degree-7 polynomial arctangent in float32  vs  the float/double chain of the
  reference (float division, correctly rounded atanf, double scaling, "+ 2 oh", fmod);
  reports how often the two land on the same float, the largest difference, and whether any
  round(j_f) decision outside the `zone` band differs (it must not).

    python tools/logpolar_numerics_sim.py 3840 1920 > profiles/r02_logpolar_numerics_sim.txt
"""
import math
import sys

import numpy as np

f32, f64 = np.float32, np.float64
ATAN = [0.9999993443489075, -0.33329862356185913, 0.19946566224098206, -0.1390863060951233,
        0.09642196446657181, -0.05591226741671562, 0.02186289243400097, -0.004054544493556023]


def ref_ij(W, H, ow, oh, cx, cy):
    cxp, cyp = int(f32(cx) * f32(W)), int(f32(cy) * f32(H))
    x = np.arange(W)[None, :].repeat(H, 0)
    y = np.arange(H)[:, None].repeat(W, 1)
    x = np.where(x - cxp > W // 2, x - W, np.where(x - cxp < -(W // 2), x + W, x))
    dx, dy = x - cxp, y - cyp
    d2 = dx.astype(f64) ** 2 + dy.astype(f64) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        i_f = np.where(d2 == 0, 0.0, ow * (np.log(np.sqrt(d2)) / f64(f32(10.0)))).astype(f32)
        q = dy.astype(f32) / dx.astype(f32)
        at = np.arctan(q.astype(f64)).astype(f32)
        j1 = ((at.astype(f64) + math.pi * (dx < 0)) * (f64(f32(oh)) / (2.0 * math.pi))).astype(f32)
        j1 = np.fmod((j1 + f32(2 * oh)).astype(f64), float(oh)).astype(f32)
    j0 = ((math.pi / 2 + math.pi * (dy < 0)) * (oh / (2.0 * math.pi))).astype(f32)
    return dx, dy, i_f, np.where(dx != 0, j1, j0)


def lntab(ow):
    K, tab = ow / 20.0, np.zeros((31, 128, 2))
    for e in range(31):
        for k in range(128):
            lo, hi = math.ldexp(1 + k / 128, e), math.ldexp(1 + (k + 1) / 128, e)
            c = 0.5 * (lo + hi)
            if math.ceil(hi) - math.ceil(lo) <= 1:
                c = max(float(math.ceil(lo)), 1.0)
            tab[e, k] = (1.0 / c, K * math.log(c))
    return tab.reshape(-1, 2)


def fast_i(d2, ow, tab):
    K = ow / 20.0
    bits = np.maximum(d2.astype(f32), f32(1)).view(np.uint32)
    t = tab[(bits >> 16).astype(np.int64) - (127 << 7)]
    r = d2.astype(f64) * t[..., 0] - 1.0
    p = ((r * (-K / 4) + K / 3) * r - K / 2) * r + K
    return np.where(d2 == 0, f32(0), (r * p + t[..., 1]).astype(f32))


def fma(a, b, c):
    return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)


def fast_j(dx, dy, oh):
    ax, ay = np.abs(dx).astype(f32), np.abs(dy).astype(f32)
    with np.errstate(divide="ignore"):
        rdx = np.where(dx == 0, f32(0), f32(1) / ax).astype(f32)
        rdy = np.where(dy == 0, f32(0), f32(1) / ay).astype(f32)
    swap = ay > ax
    a = (np.minimum(ax, ay) * np.where(swap, rdy, rdx)).astype(f32)
    s = (a * a).astype(f32)
    p = np.full(a.shape, f32(ATAN[-1]))
    for k in ATAN[-2::-1]:
        p = fma(p, s, np.full(a.shape, f32(k)))
    t = (p * a).astype(f32)
    turns = f32(oh / (2 * math.pi))
    neg = (dy < 0) ^ (dx < 0)
    sp = np.where(neg ^ swap, -turns, turns).astype(f32)
    base = (2 * oh + np.where(dx < 0, oh / 2, 0) + np.where(swap, np.where(neg, -oh / 4, oh / 4), 0)).astype(f32)
    val = fma(sp, t, base)
    return np.where(val >= f32(2 * oh), val - f32(2 * oh), val - f32(oh)).astype(f32)


W, H = int(sys.argv[1]), int(sys.argv[2])
ow, oh = 16 * math.ceil(W / 1.8 / 16), 16 * math.ceil(H / 1.8 / 16)
top = f32(2.75 * oh)
zone = f32(6e-7 * oh / (2 * math.pi)) + f32(1.5) * (np.nextafter(top, f32(np.inf)) - top)
tab = lntab(ow)
print("%dx%d -> %dx%d, zone %.3e" % (W, H, ow, oh, zone))
for cx, cy in [(0.5, 0.5), (0.65, 0.75), (0.02, 0.3), (1.0, 1.0)]:
    dx, dy, i_f, j_f = ref_ij(W, H, ow, oh, cx, cy)
    d2 = dx.astype(np.int64) ** 2 + dy.astype(np.int64) ** 2
    fi = fast_i(d2, ow, tab)
    fj = fast_j(dx, dy, oh)
    centre = d2 == 0  # the kernel sends the gaze pixel itself to the exact path
    d = fj.astype(f64) - j_f.astype(f64)
    d = np.where(d > oh / 2, d - oh, np.where(d < -oh / 2, d + oh, d))
    d[centre] = 0
    wrong_round = (np.round(j_f) != np.round(fj)) & ~centre
    amb = np.abs((fj - np.floor(fj)) - f32(0.5)) < zone
    print("gaze (%g, %g): i_f identical on %d of %d pixels; j_f identical on %.2f %%, max |diff| %.2e; "
          "round(j_f) differs on %d pixels, %d of them outside the zone band (%.3f %% of the pixels are "
          "inside it)" % (cx, cy, int((fi == i_f).sum()), i_f.size, 100 * (d == 0).mean(),
                          np.abs(d).max(), int(wrong_round.sum()), int((wrong_round & ~amb).sum()),
                          100 * amb.mean()))

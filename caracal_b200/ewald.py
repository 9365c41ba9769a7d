"""Host-side set-up of the SPME reciprocal sum, the mirror of set_periodic.f90:114-231 for a run without
the Fortran drivers: PME grid size from the `multi` table (from the x box length, as the reference
does), Ewald coefficient by bisection on erfc(a r_c)/r_c = 1e-8 with r_c = 7 Angstrom, bsorder = 5 and
the B-spline moduli (bspline.f90, dftmod.f90).  The Fortran side hands the same quantities over from
module pbc_mod (crcl_set_ewald)."""
import math

import numpy as np

BOHR = 0.52917721092
MULTI = [2, 4, 6, 8, 10, 12, 16, 18, 20, 24, 30, 32, 36, 40, 48, 50, 54, 60, 64, 72, 80, 90, 96, 100, 108, 120, 128,
         144, 150, 160, 162, 180, 192, 200, 216, 240, 250, 256, 270, 288, 300, 320, 324, 360, 384, 400, 432, 450, 480,
         486, 500, 512, 540, 576, 600, 640, 648, 720, 750, 768, 800, 810, 864]


def _bspline(x, n):
    """bspline.f90:30-60, returns c(1..n)"""
    c = np.zeros(n + 1)
    c[1], c[2] = 1.0 - x, x
    for k in range(3, n + 1):
        denom = 1.0 / (k - 1)
        c[k] = x * c[k - 1] * denom
        for i in range(1, k - 1):
            c[k - i] = ((x + i) * c[k - i - 1] + ((k - i) - x) * c[k - i]) * denom
        c[1] = (1.0 - x) * c[1] * denom
    return c[1:]


def _dftmod(bsarray, nfft, order):
    """dftmod.f90:30-101"""
    j = np.arange(nfft)
    bsmod = np.zeros(nfft)
    for i in range(nfft):
        arg = 2.0 * math.pi / nfft * (i * j)
        bsmod[i] = (bsarray * np.cos(arg)).sum() ** 2 + (bsarray * np.sin(arg)).sum() ** 2
    eps = 1.0e-7
    if bsmod[0] < eps:
        bsmod[0] = 0.5 * bsmod[1]
    for i in range(1, nfft - 1):
        if bsmod[i] < eps:
            bsmod[i] = 0.5 * (bsmod[i - 1] + bsmod[i + 1])
    if bsmod[nfft - 1] < eps:
        bsmod[nfft - 1] = 0.5 * bsmod[nfft - 2]
    jj = np.arange(1, 51)
    for i in range(1, nfft + 1):
        k = i - 1
        if i > nfft // 2:
            k -= nfft
        if k == 0:
            zeta = 1.0
        else:
            f = math.pi * k / nfft
            a1, a2 = f / (f + math.pi * jj), f / (f - math.pi * jj)
            zeta = (1.0 + (a1 ** (2 * order)).sum() + (a2 ** (2 * order)).sum()) / \
                   (1.0 + (a1 ** order).sum() + (a2 ** order).sum())
        bsmod[i - 1] *= zeta * zeta
    return bsmod


def ewald_setup(box):
    """box[3] in bohr -> dict(box, a_ewald, nfft, bsorder, bsmod[3, nfft])"""
    box = np.asarray(box, dtype=np.float64)
    r_ew_cut = 7 / BOHR
    ifft = int(box[0] * BOHR * 1.2 - 1e-8) + 1
    nfft = 864
    for k in reversed(MULTI):
        if k >= ifft:
            nfft = k
    nfft = max(nfft, 16)
    eps = 1.0e-8
    ratio, x, i = eps + 1.0, 0.5, 0
    while ratio >= eps:
        i += 1
        x *= 2.0
        ratio = math.erfc(x * r_ew_cut) / r_ew_cut
    lo, hi = 0.0, x
    for _ in range(i + 60):
        x = (lo + hi) / 2.0
        if math.erfc(x * r_ew_cut) / r_ew_cut >= eps:
            lo = x
        else:
            hi = x
    bsorder = 5
    bsarray = np.zeros(nfft)
    bsarray[1:1 + bsorder] = _bspline(0.0, bsorder)
    bsmod = _dftmod(bsarray, nfft, bsorder)
    return dict(box=box, a_ewald=x, nfft=nfft, bsorder=bsorder, bsmod=np.stack([bsmod, bsmod, bsmod]))

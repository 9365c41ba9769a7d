"""CPU model of the cell sweep of the inter-molecular QMDFF part (caracal_b200/csrc/qmdff_kernels.cu:
qm_cellsort_kernel, qm_inter_cell_kernel): the same cell assignment, half-shell rows, wrap segments and shifts in
numpy.  Property: the candidate pairs the sweep produces are exactly the pairs whose minimum-image distance lies inside
the (slackened) cut-off, every unordered pair once -- for M = 2 and 3, cubic and non-cubic boxes, atoms on faces."""
import numpy as np
import pytest

SLACK = 1.002


def grid(box, rc, m):
    rcs = rc * np.sqrt(SLACK) * 1.0005
    nc = [int(np.floor(box[d] * m / rcs)) for d in range(3)]
    return nc if all(n >= 2 * m + 1 for n in nc) else None


def sweep_pairs(x, box, rc, m):
    n = len(x)
    nc = grid(box, rc, m)
    assert nc is not None
    u = x - box * np.floor(x / box)
    ci = np.minimum((u * (np.array(nc) / box)).astype(int), np.array(nc) - 1)
    cell = (ci[:, 2] * nc[1] + ci[:, 1]) * nc[0] + ci[:, 0]
    order = np.argsort(cell, kind="stable")
    w = u[order].astype(np.float32)
    cs = np.searchsorted(cell[order], np.arange(nc[0] * nc[1] * nc[2] + 1))
    rc2f = np.float32(rc * rc) * np.float32(SLACK)
    pairs = []
    ncx, ncy, ncz = nc
    for home in range(ncx * ncy * ncz):
        hb, he = cs[home], cs[home + 1]
        if hb == he:
            continue
        cx, cy, cz = home % ncx, (home // ncx) % ncy, home // (ncx * ncy)
        runs = [(hb, he - hb, (0.0, 0.0, 0.0), True)]
        rows = [(0, oy) for oy in range(0, m + 1)] + [(oz, oy) for oz in range(1, m + 1) for oy in range(-m, m + 1)]
        for oz, oy in rows:
            z2, y2, shz, shy = cz + oz, cy + oy, 0.0, 0.0
            if z2 >= ncz:
                z2, shz = z2 - ncz, box[2]
            if y2 < 0:
                y2, shy = y2 + ncy, -box[1]
            elif y2 >= ncy:
                y2, shy = y2 - ncy, box[1]
            row = (z2 * ncy + y2) * ncx
            xa, xb = cx + (1 if (oz == 0 and oy == 0) else -m), cx + m
            lo = [xa, max(xa, 0), ncx]
            hi = [min(-1, xb), min(xb, ncx - 1), xb]
            off = [ncx, 0, -ncx]
            shx = [-box[0], 0.0, box[0]]
            for s in range(3):
                if lo[s] <= hi[s]:
                    b = cs[row + lo[s] + off[s]]
                    ln = cs[row + hi[s] + off[s] + 1] - b
                    runs.append((b, ln, (shx[s], shy, shz), False))
        for ipos in range(hb, he):
            for b, ln, sh, own in runs:
                for jpos in range(b, b + ln):
                    if own and not jpos > ipos:
                        continue
                    c = w[jpos] + np.array(sh, dtype=np.float32)
                    d = w[ipos] - c
                    if np.float32(d @ d) <= rc2f:
                        a, bb = order[ipos], order[jpos]
                        pairs.append((min(a, bb), max(a, bb)))
    return pairs


@pytest.mark.parametrize("m,box", [(2, [10.5, 10.5, 10.5]), (3, [15.0, 15.0, 15.0]), (2, [11.0, 13.0, 16.5]),
                                   (3, [14.5, 21.0, 17.0])])
def test_cell_sweep_finds_every_pair_inside_the_cutoff_once(m, box):
    rng = np.random.default_rng(m * 100 + int(box[0]))
    box = np.array(box)
    rc, n = 4.0, 260
    x = rng.uniform(-0.5, 1.5, (n, 3)) * box                     # unwrapped coordinates, some outside the box
    x[:6] = [[0, 0, 0], box, [0, box[1], 0], [1e-9, -1e-9, box[2] - 1e-9], box * 0.5, [box[0], 0, box[2]]]
    got = sweep_pairs(x, box, rc, m)
    assert len(got) == len(set(got)), "a pair was produced twice"
    got = set(got)
    d = x[:, None, :] - x[None, :, :]
    d -= box * np.round(d / box)
    r2 = (d ** 2).sum(-1)
    iu = np.triu_indices(n, 1)
    inside = {(int(a), int(b)) for a, b, v in zip(iu[0], iu[1], r2[iu]) if v <= rc * rc}
    loose = {(int(a), int(b)) for a, b, v in zip(iu[0], iu[1], r2[iu]) if v <= rc * rc * SLACK * 1.0001}
    assert inside <= got, "pairs inside the cut-off are missing: %s" % sorted(inside - got)[:5]
    assert got <= loose, "pairs outside the slackened cut-off were produced"
    assert len(inside) > 500


def test_small_boxes_keep_the_n2_sweep():
    assert grid([9.9, 20.0, 20.0], 4.0, 2) is None and grid([10.1, 20.0, 20.0], 4.0, 2) is not None
    assert grid([9.3, 14.1, 14.1], 4.0, 3) is None and grid([9.4, 14.1, 14.1], 4.0, 3) is not None

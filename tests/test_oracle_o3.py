"""CPU checks of the O3 1 1A" restatement (oracle/pes_o3.c <- egrad_o3.f).  The reference ships
no vectors for this surface (parity unpinned); pinned by the zero of energy its Eref encodes (O + O2 at the fitted
diatomic minimum), finite differences, permutation and rigid-motion invariance, a stationary point, and -- where the
reference tree is present -- by re-reading the recurrence table, the basis selection and the coefficients out of the
source text."""
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from tests import common as C

REF = "/root/reference/src/egrad_o3.f"


def test_zero_of_energy_is_o_plus_o2():
    # Eref = -0.19172848 Eh shifts the fit so that O + O2(r_e) is zero (pot_o3 :93-94,112)
    x = np.array([[0, 0, 0], [1.2075 / C.BOHR, 0, 0], [0, 60.0, 0]])
    V, g, _ = O.egrad("o3", x[None])
    assert abs(V[0]) < 1e-6 and np.abs(g).max() < 5e-4
    for r in (1.15, 1.30):                              # the diatomic well: both sides are higher
        y = x.copy()
        y[1, 0] = r / C.BOHR
        assert O.egrad("o3", y[None])[0][0] > V[0] + 1e-3


def test_gradient_is_the_derivative_of_the_energy():
    rng = np.random.default_rng(1)
    q = C.ts_cloud("o3", 20, 0.2, rng, min_dist=1.6)
    V, g, info = O.egrad("o3", q)
    assert info == 0
    h = 1e-5
    for c in range(9):
        dq = np.zeros(9)
        dq[c] = h
        Vp, Vm = O.egrad("o3", q + dq.reshape(3, 3))[0], O.egrad("o3", q - dq.reshape(3, 3))[0]
        assert np.abs((Vp - Vm) / (2 * h) - g.reshape(-1, 9)[:, c]).max() < 5e-9


def test_invariances_and_the_c2v_minimum():
    rng = np.random.default_rng(2)
    q = C.ts_cloud("o3", 100, 0.25, rng, min_dist=1.6)
    V, g, _ = O.egrad("o3", q)
    for perm in ([1, 0, 2], [2, 1, 0], [1, 2, 0]):      # three identical atoms
        V2, g2, _ = O.egrad("o3", q[:, perm])
        assert np.abs(V2 - V).max() < 1e-12 and np.abs(g2 - g[:, perm]).max() < 1e-11
    A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    V3, g3, _ = O.egrad("o3", q @ A.T - 3.0)
    assert np.abs(V3 - V).max() < 1e-12 and np.abs(g3 - g @ A.T).max() < 1e-11
    assert np.abs(g.sum(axis=1)).max() < 1e-13
    # the shallow C2v minimum of this excited singlet state: r = 1.35709 A, 105.2475 deg, 10.17 kcal/mol above O + O2
    r, th = 1.3570860596 / C.BOHR, np.deg2rad(105.24749176)
    x = np.array([[0, 0, 0], [r, 0, 0], [r * np.cos(th), r * np.sin(th), 0]])
    Vm, gm, _ = O.egrad("o3", x[None])
    assert np.abs(gm).max() < 1e-7 and abs(Vm[0] * 627.5095 - 10.1652) < 1e-3


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_tables_against_the_source_text():
    src = open(REF).read()
    mine = open(os.path.join(os.path.dirname(__file__), "..", "oracle", "pes_o3.c")).read()
    coef = {int(m.group(1)): float(m.group(2).replace("D", "e"))
            for m in re.finditer(r"data C\(\s*(\d+)\)\s*/\s*([-0-9.D+]+)\s*/", src)}
    got = [float(x) for x in re.findall(r"[-+]?\d\.\d+e[+-]\d+", mine[mine.index("O3_C[56]"):mine.index("/* ev2gm2_o3")])]
    assert len(coef) == 56 and got == [coef[i + 1] for i in range(56)]
    poly = src[src.index("subroutine EvPoly_o3"):src.index("end subroutine EvPoly_o3")]
    rec = {int(m.group(1)): m.group(2).replace(" ", "") for m in re.finditer(r"p\(\s*(\d+)\)\s*=\s*(.*)", poly)}
    tab = re.findall(r"\{(-?\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\}", mine[mine.index("O3_REC[67][6]"):mine.index("/* evbas_o3")])
    assert len(tab) == 67 and len(rec) == 67
    for k, t in enumerate(tab):
        a, b, n, c1, c2, c3 = map(int, t)
        if a < 0:
            assert "rm(" in rec[k]
            continue
        want = "p(%d)*p(%d)" % (a, b) + "".join("-p(%d)" % c for c in (c1, c2, c3)[:n])
        assert re.sub(r"p\(0*(\d+)\)", lambda m: "p(%d)" % int(m.group(1)), rec[k]) == want, k
    # the derivative routine is the product rule on the same recurrences, operands in this order (:771-846)
    dp = src[src.index("subroutine EvdPdr_o3"):src.index("subroutine evdbdr_o3")].replace("\n     $", "")
    drec = {int(m.group(1)): m.group(2).replace(" ", "") for m in re.finditer(r"dpdr\(i,\s*(\d+)\)\s*=\s*(.*)", dp)}
    for k, t in enumerate(tab):
        a, b, n, c1, c2, c3 = map(int, t)
        if a < 0:
            continue
        want = "dpdr(i,%d)*p(%d)+p(%d)*dpdr(i,%d)" % (a, b, a, b) + "".join("-dpdr(i,%d)" % c for c in (c1, c2, c3)[:n])
        norm = re.sub(r"\(i,0*(\d+)\)", lambda m: "(i,%d)" % int(m.group(1)), re.sub(r"p\(0*(\d+)\)", lambda m: "p(%d)" % int(m.group(1)), drec[k]))
        assert norm == want, (k, norm, want)

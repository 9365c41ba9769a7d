"""The reference's SHIPPED QMDFF inputs through the oracle (C5 parity on
examples/dynamic/ethanol_box/box.{qmdff,xyz}, C4 parity on examples/evbopt/DG-EVB/min1|min2.qmdff): the fixture
tests/golden/qmdff_examples.npz holds the example files and the D3 / setnonb element data, tests/qmdff_file.py
restates the host-side set-up (prepare.f90, setnonb.f90, ncoord_qmdff.f90, getc6.f90, set_periodic.f90).
Known answers the files themselves provide: a QMDFF is built around its own reference structure -- every bonded
term is at its minimum there, so ff_eg gives exactly zero gradient on the structure stored in the file -- and D3
gives the published C6 coefficients for the coordination numbers of that structure."""
import numpy as np
import pytest

from tests import qmdff_file as QF


def test_d3_coordination_numbers_and_c6_of_the_example_molecule():
    T = QF.tables("min1", e_zero=0.0)
    at, cn = T["at"], T["cn"]
    # water ... ammonia complex: O (2 H: one bonded + one H bond), N with three hydrogens
    assert list(at) == [8, 1, 7, 1, 1, 1]
    assert 0.9 < cn[1] < 1.1 and 2.9 < cn[2] < 3.3 and 0.9 < cn[0] < 1.3
    c6 = T["c6xy"]
    assert np.allclose(c6, c6.T)
    # D3 reference values bracket the interpolated ones: C6(H-H) in [3.03, 7.6], C6(N-N) in [15.6, 25.3] a.u.
    assert 3.0 < c6[1, 1] < 7.6 and 15.0 < c6[2, 2] < 25.3 and 10.0 < c6[0, 0] < 15.6


@pytest.mark.parametrize("tag", ["min1", "min2"])
def test_bonded_terms_vanish_on_the_force_fields_own_structure(oracle, tag):
    T = QF.tables(tag)
    Q = oracle.Qmdff(T)
    e, g = Q.ff_eg(T["xyz"])
    # every bond / angle term sits at its reference value on the structure stored in the .qmdff file (the box
    # file does not have this property: its stored lattice of molecules is off the minimum by a few 1e-3 Eh/bohr)
    assert np.abs(g).max() < 5e-6, np.abs(g).max()
    x = T["xyz"] + np.random.default_rng(0).normal(0, 0.05, T["xyz"].shape)
    e2, g2 = Q.ff_eg(x)
    assert e2 > e and np.abs(g2).max() > 1e-3


def test_ethanol_box_energy_gradient_consistency(oracle):
    """periodic box as equilibration.key sets it up (27 A, Zahn, 10 A cut-offs) at the shipped start structure"""
    T = QF.tables("box", periodic_angstrom=[27.0, 27.0, 27.0])
    Q = oracle.Qmdff(T)
    x = QF.box_start_bohr()
    assert x.shape == (1125, 3) and T["nmols"] == 125 and len(T["bond"]) == 2625 and len(T["nci"]) == 1875
    V, g = Q.egrad(x[None])
    assert np.isfinite(V).all() and np.abs(g[0].sum(axis=0)).max() < 1e-9
    # a liquid-like box of 125 ethanol: the non-covalent + Coulomb energy per molecule is a few kcal/mol, the
    # largest force is that of a thermally distorted bond (not a clash)
    assert -0.05 < V[0] / 125 < 0.05 and np.abs(g).max() < 0.2
    # a rigid shift leaves everything unchanged; so does the lattice translation of one molecule once the H-bond
    # part is left out -- eabhag.f90:60-62 images only one of its three vectors, which is reproduced
    V2, g2 = Q.egrad(x[None] + 0.37)
    assert abs(V2[0] - V[0]) < 1e-9 and np.abs(g2 - g).max() < 1e-9
    Tn = {k: v for k, v in T.items() if k not in ("hb", "vhb", "scalehb", "scalexb", "q_glob")}
    Tn["nhb"] = 0
    Qn = oracle.Qmdff(Tn)
    y = x.copy()
    y[T["molnum"] == 17] += np.array([T["box"][0], -T["box"][1], 0.0])
    Vn, gn = Qn.egrad(x[None])
    Vn2, gn2 = Qn.egrad(y[None])
    assert abs(Vn2[0] - Vn[0]) < 1e-9 and np.abs(gn2 - gn).max() < 1e-9
    assert Vn[0] != V[0]                                            # the donor/acceptor search did contribute
    # bonded part: finite differences on a few atoms
    e0, gb = Q.ff_eg(x)
    rng = np.random.default_rng(1)
    for _ in range(6):
        a, d = int(rng.integers(0, 1125)), int(rng.integers(0, 3))
        xp, xm = x.copy(), x.copy()
        xp[a, d] += 1e-5
        xm[a, d] -= 1e-5
        assert abs((Q.ff_eg(xp)[0] - Q.ff_eg(xm)[0]) / 2e-5 - gb[a, d]) < 2e-9


def test_diabats_of_the_dgevb_example_follow_the_reference_path_energies(oracle):
    """struc.xyz carries the reference (QM) energy of each of its 42 path structures; evbopt.key's eshift values are
    the energies of the two end points.  Each QMDFF, evaluated by the oracle with the tables of tests/qmdff_file.py,
    reproduces the reference energy at its own minimum up to its non-covalent part (+0.25 kcal/mol for min1, the
    -4 kcal/mol hydrogen bond of ff_hb for min2) and rises away from it, and the lower diabat tracks the reference
    path to within the barrier region's 0.035 Eh."""
    T1, T2, cd, _ = QF.dgevb_example()
    x, e_ref = QF.dgevb_path()
    assert abs(e_ref[0] - T1["e_zero"]) < 1e-6 and abs(e_ref[-1] - T2["e_zero"]) < 1e-6
    e1 = oracle.Qmdff(T1).egrad(x)[0]
    D = oracle.Dgevb(T1, T2, dict(mode=1, coord_def=cd, g_thres=1e-10, point_int=np.zeros((1, len(cd))),
                                  alph=np.ones(1), b_vec=np.zeros(64)))
    e2 = np.array([D.second_state(xx)[0] for xx in x])
    kcal = 627.5095
    assert abs((e1[0] - e_ref[0]) * kcal - 0.255) < 0.02
    assert abs((e2[-1] - e_ref[-1]) * kcal + 4.06) < 0.1
    assert np.all(np.diff(e1 - e_ref)[3:15] > 0)              # diabat 1 leaves the reference curve monotonically
    assert np.abs(np.minimum(e1, e2) - e_ref).max() < 0.035
    assert (e1 < e2)[:15].all() and (e2 < e1)[-15:].all()     # the diabats cross between the minima

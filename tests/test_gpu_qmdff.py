"""GPU parity of the QMDFF force-field kernels (rows a16/a17 of SURVEY.md section 8): energies and
gradients per image against the oracle within 1e-10 relative, through crcl_egrad, plus an RPMD
trajectory on the split path with the QMDFF as PES and invariances at the periodic-box size."""
import numpy as np
import pytest

from tests import common as C
from tests.qmdff_synth import make_system

pytestmark = pytest.mark.gpu


def handle(gpu, T, nbeads=1, dt_fs=0.5):
    mass = np.array([C.atomic_mass_au({1: "H", 6: "C", 8: "O", 17: "CL"}[int(z)]) for z in T["at"]])
    g = gpu.RPMD(gpu.PES_QMDFF, nbeads, mass, C.beta_calc_rate(300.0), C.dt_au(dt_fs))
    g.set_qmdff(T)
    return g, mass


def torsion_conditioning(T, x):
    """min |sin phi| over the proper torsions of each image.  dphidr.f90:74-84 divides by
    |na||nb| sin(phi) with phi = acos(na.nb) (valijkl.f90:80-98): near phi = 0 or pi a one-ulp libm
    difference in acos is amplified by 1/sin^2(phi) in that torsion's gradient, in the reference as
    much as here, so the comparison tolerance follows the conditioning of the reference formula."""
    out = np.ones(x.shape[0])
    tr = np.array([t[:4] - 1 for t in T["tors"] if t[5] != 2])
    if len(tr) == 0:
        return out
    L = T["box"] if T["periodic"] else None

    def img(v):
        return v - L * np.round(v / L) if L is not None else v
    ra = img(x[:, tr[:, 1]] - x[:, tr[:, 0]])
    rb = img(x[:, tr[:, 2]] - x[:, tr[:, 1]])
    rc = img(x[:, tr[:, 3]] - x[:, tr[:, 2]])
    na, nb = np.cross(ra, rb), np.cross(rb, rc)
    cs = (na * nb).sum(-1) / np.linalg.norm(na, axis=-1) / np.linalg.norm(nb, axis=-1)
    return np.sqrt(np.maximum(1 - cs ** 2, 1e-30)).min(axis=1)


@pytest.mark.parametrize("periodic,zahn,nmol,nimg", [(True, True, 8, 48), (True, False, 8, 48), (False, False, 8, 48),
                                                     (True, True, 125, 3), (False, False, 1, 16)])
def test_egrad_matches_oracle(gpu, oracle, periodic, zahn, nmol, nimg):
    T = make_system(nmol=nmol, seed=nmol + periodic + 2 * zahn, periodic=periodic, zahn=zahn)
    g, _ = handle(gpu, T)
    Q = oracle.Qmdff(T)
    rng = np.random.default_rng(1)
    x = T["xyz"][None] + rng.normal(0, 0.06, (nimg,) + T["xyz"].shape)
    Vo, go = Q.egrad(x)
    Vd, gd, _ = g.egrad(x)
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    tol = np.maximum(C.TOL_EG, 2e-17 / torsion_conditioning(T, x) ** 2)
    err = C.rel_err_G(gd.reshape(go.shape), go)
    assert (err < tol).all(), (err / tol).max()
    # how far the stated 1e-10 is widened: over the five parameter sets 15 of 163 images hold a torsion
    # within |sin phi| < 4.5e-4 of 0 or pi and are compared at a looser bound, the loosest 2.2e-8; every other image at 1e-10
    widened = int((tol > C.TOL_EG).sum())
    assert widened <= max(2, nimg // 6) and tol.max() < 5e-8, (widened, tol.max())


@pytest.mark.parametrize("periodic,nmol,nimg,halogen", [(True, 12, 32, 0.0), (True, 12, 32, 0.3), (False, 12, 32, 0.3),
                                                         (True, 125, 2, 0.1)])
def test_hbond_terms_match_oracle(gpu, oracle, periodic, nmol, nimg, halogen):
    """row a18: ff_hb list + donor/acceptor search, analytic H-bond (eabhag) and numeric X-bond
    (eabxag, including the broadcast of the last numeric derivative)"""
    T = make_system(nmol=nmol, seed=31 + nmol, periodic=periodic, zahn=periodic, hb=True, frac_halogen=halogen)
    g, _ = handle(gpu, T)
    Q = oracle.Qmdff(T)
    T0 = {k: v for k, v in T.items() if k not in ("hb", "vhb", "scalehb", "scalexb", "q_glob")}
    T0["nhb"] = 0
    Q0 = oracle.Qmdff(T0)
    rng = np.random.default_rng(1)
    x = T["xyz"][None] + rng.normal(0, 0.06, (nimg,) + T["xyz"].shape)
    Vo, go = Q.egrad(x)
    V0, g0 = Q0.egrad(x)
    assert np.abs(Vo - V0).max() > 1e-5                       # the H/X-bond terms are really active
    Vd, gd, _ = g.egrad(x)
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    tol = np.maximum(C.TOL_EG, 2e-17 / torsion_conditioning(T, x) ** 2)
    err = C.rel_err_G(gd.reshape(go.shape), go)
    assert (err < tol).all(), (err / tol).max()
    # the hb part alone, free of the torsion conditioning: (full - without) on both sides.  The X-bond
    # gradient is a central difference with step 1e-6 (eabxag.f90:108-140): libm-level differences in
    # eabx are amplified by 5e5, so the bound is relative to the whole gradient, not to the hb part
    g2, _ = handle(gpu, T0)
    Vd0, gd0, _ = g2.egrad(x)
    dh_o, dh_d = go - g0, gd.reshape(go.shape) - gd0.reshape(go.shape)
    assert np.abs(dh_d - dh_o).max() < C.TOL_EG * np.abs(go).max()


def test_rpmd_with_qmdff_on_split_path(gpu, oracle):
    T = make_system(nmol=4, seed=9, periodic=True, zahn=True)
    nb, nsteps = 4, 40
    g, mass = handle(gpu, T, nbeads=nb)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 9, 300.0)
    Q = oracle.Qmdff(T)
    rng = np.random.default_rng(2)
    q0 = T["xyz"][None, None] + rng.normal(0, 0.02, (1, nb) + T["xyz"].shape)
    q = q0.copy()
    tid = np.array([11], dtype=np.uint32)
    p, d, dxi, ev = g.mdinit(q, 0, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=-1, traj_id=tid, event=ev)
    o = oracle.System(0, nb, mass, C.beta_calc_rate(300.0), C.dt_au(0.5))
    o.set_custom_grad(lambda x: tuple(a[0] for a in Q.egrad(x)))
    o.q[:] = q0[0]
    o.set_rng(C.SEED, 11)
    o.set_thermostat(1, 9, 300.0)
    o.mdinit(0.0, 0)
    for i in range(1, nsteps + 1):
        epo, _, sto = o.verlet(i, 0.0, -1)
        assert sto == 0
    assert st[0] == 0
    assert np.abs(q[0] - o.q).max() < C.TOL_QP
    assert (np.abs(p[0] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
    assert abs(ep[0] - epo) < 1e-9 * max(1.0, abs(epo))


def test_invariances_at_box_size(gpu):
    """~3000 atoms (config 5 shape), periodic Zahn: net force vanishes, rigid and lattice translations
    leave E and g unchanged; 8 images = 8 beads in one call"""
    T = make_system(nmol=385, seed=12, periodic=True, zahn=True)
    assert 2800 < T["n"] < 3500
    g, _ = handle(gpu, T)
    rng = np.random.default_rng(3)
    x = T["xyz"][None] + rng.normal(0, 0.05, (8,) + T["xyz"].shape)
    V, grad, _ = g.egrad(x)
    grad = grad.reshape(x.shape)
    assert np.isfinite(V).all()
    assert np.abs(grad.sum(axis=1)).max() < 1e-9
    V2, g2, _ = g.egrad(x + np.array([1.0, -2.0, 0.5]))
    assert C.rel_err_E(V2, V).max() < 1e-9 and np.abs(g2.reshape(x.shape) - grad).max() < 1e-9
    y = x.copy()
    y[:, T["molnum"] == 7] += np.array([0, T["box"][1], 0])
    V3, g3, _ = g.egrad(y)
    assert C.rel_err_E(V3, V).max() < 1e-9 and np.abs(g3.reshape(x.shape) - grad).max() < 1e-8

"""GPU parity of the second QMDFF (*_two semantics, row a19) and the DG-EVB coupling / mixing (row a20
of SURVEY.md section 8): per-image energies and gradients against the oracle within 1e-10 relative,
through crcl_egrad, and an RPMD trajectory on the split path with the DG-EVB surface as PES."""
import numpy as np
import pytest

from tests import common as C
from tests.qmdff_synth import HEXANE, make_dgevb
from tests.test_gpu_qmdff import torsion_conditioning

pytestmark = pytest.mark.gpu


def wilson_conditioning(E, x):
    """min |sin phi| over the dihedral internal coordinates of each image.  The reference builds the
    Wilson B matrix by central differences with shift 1e-3 (calc_wilson.f90:114-178, num_wilson is
    always .true., init_int.f90:142) of dihed.f90's phi = acos(cv): a libm-level difference eps in cv
    becomes eps / sin(phi) in phi and eps / (2e-3 sin(phi)) in B -- in the reference as much as here."""
    out = np.ones(x.shape[0])
    for cd in E["coord_def"]:
        if cd[0] != 3:
            continue
        a1, a2, a3, a4 = (int(v) - 1 for v in cd[1:5])
        u, v, w = x[:, a1] - x[:, a2], x[:, a4] - x[:, a3], x[:, a3] - x[:, a2]
        uxw, vxw = np.cross(u, w), np.cross(v, w)
        cv = (uxw * vxw).sum(-1) / np.linalg.norm(uxw, axis=-1) / np.linalg.norm(vxw, axis=-1)
        out = np.minimum(out, np.sqrt(np.maximum(1 - cv ** 2, 1e-30)))
    return out


def handle(gpu, T1, T2, E, nbeads=1, dt_fs=0.5):
    mass = np.array([C.atomic_mass_au({1: "H", 6: "C", 8: "O", 17: "CL"}[int(z)]) for z in T1["at"]])
    g = gpu.RPMD(gpu.PES_DGEVB, nbeads, mass, C.beta_calc_rate(300.0), C.dt_au(dt_fs))
    g.set_qmdff(T1)
    g.set_qmdff(T2, second=True)
    g.set_dgevb(E)
    return g, mass


@pytest.mark.parametrize("mode,npoints,nimg", [(1, 5, 64), (2, 5, 64), (3, 5, 64), (3, 7, 1), (2, 1, 7)])
def test_egrad_matches_oracle(gpu, oracle, mode, npoints, nimg):
    T1, T2, E = make_dgevb(seed=mode + npoints, mode=mode, npoints=npoints)
    g, _ = handle(gpu, T1, T2, E)
    D = oracle.Dgevb(T1, T2, E)
    rng = np.random.default_rng(1)
    x = T1["xyz"][None] + rng.normal(0, 0.05, (nimg,) + T1["xyz"].shape)
    Vo, go = D.egrad(x)
    e1, _ = oracle.Qmdff(T1).egrad(x)
    assert np.abs(Vo - e1).max() > 1e-5        # the coupling / second state really contribute
    Vd, gd, _ = g.egrad(x)
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    tol = np.maximum(C.TOL_EG, 2e-17 / torsion_conditioning(T1, x) ** 2)
    tol = np.maximum(tol, 4e-12 / wilson_conditioning(E, x))
    err = C.rel_err_G(gd.reshape(go.shape), go)
    assert (err < tol).all(), (err / tol).max()
    assert (tol < 1e-9).mean() > 0.8          # the conditioning model must not swallow the test


def test_gaussian_threshold_and_negative_root(gpu, oracle):
    """expo < g_thres skips a Gaussian (sum_v12.f90 / sum_dv12.f90 `cycle`); a negative
    dE^2 + 4 V12 switches the coupling gradient off (gradient.f90:476-489)"""
    T1, T2, E = make_dgevb(seed=8, mode=1, npoints=3)
    rng = np.random.default_rng(4)
    x = T1["xyz"][None] + rng.normal(0, 0.03, (16,) + T1["xyz"].shape)
    for Ev in (dict(E, alph=np.array([1e6, 1.0, 1e6])), dict(E, b_vec=-np.abs(E["b_vec"]) - 0.05)):
        g, _ = handle(gpu, T1, T2, Ev)
        D = oracle.Dgevb(T1, T2, Ev)
        Vo, go = D.egrad(x)
        Vd, gd, _ = g.egrad(x)
        assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
        tol = np.maximum(1e-9, 4e-12 / wilson_conditioning(Ev, x))
        assert (C.rel_err_G(gd.reshape(go.shape), go) < tol).all()


def test_needs_all_three_tables(gpu):
    T1, T2, E = make_dgevb(seed=1, mode=1, npoints=2)
    mass = np.ones(T1["n"]) * 1837.0
    g = gpu.RPMD(gpu.PES_DGEVB, 1, mass, C.beta_calc_rate(300.0), C.dt_au(0.5))
    g.set_qmdff(T1)
    with pytest.raises(gpu.CaracalGpuError):
        g.egrad(T1["xyz"][None])


def test_rpmd_with_dgevb_on_split_path(gpu, oracle):
    T1, T2, E = make_dgevb(seed=5, mode=3, npoints=4)
    nb, nsteps = 8, 50
    g, mass = handle(gpu, T1, T2, E, nbeads=nb, dt_fs=0.2)
    g.set_seed(C.SEED)
    g.set_thermostat(1, 13, 300.0)
    D = oracle.Dgevb(T1, T2, E)
    rng = np.random.default_rng(2)
    q0 = T1["xyz"][None, None] + rng.normal(0, 0.01, (1, nb) + T1["xyz"].shape)
    q = q0.copy()
    tid = np.array([5], dtype=np.uint32)
    p, d, dxi, ev = g.mdinit(q, 0, traj_id=tid)
    ep, xr, st = g.verlet(q, p, d, nsteps=nsteps, constrain=-1, traj_id=tid, event=ev)
    o = oracle.System(0, nb, mass, C.beta_calc_rate(300.0), C.dt_au(0.2))
    o.set_custom_grad(lambda x: tuple(a[0] for a in D.egrad(x)))
    o.q[:] = q0[0]
    o.set_rng(C.SEED, 5)
    o.set_thermostat(1, 13, 300.0)
    o.mdinit(0.0, 0)
    for i in range(1, nsteps + 1):
        epo, _, sto = o.verlet(i, 0.0, -1)
        assert sto == 0
    assert st[0] == 0
    assert np.abs(q[0] - o.q).max() < C.TOL_QP
    assert (np.abs(p[0] - o.p) / np.abs(o.p).max()).max() < C.TOL_QP
    assert abs(ep[0] - epo) < 1e-9 * max(1.0, abs(epo))


def test_twenty_atom_two_state_system_matches_oracle(gpu, oracle):
    """BASELINE config 4's shape: ~20 atoms (n-hexane-like: 19 bonds, 36 angles, 45 torsions, 135
    nci pairs), 7 distributed Gaussians, mode 3, nat6 = 12 with three dihedral internal coordinates."""
    T1, T2, E = make_dgevb(seed=5, mode=3, npoints=7, template=HEXANE)
    assert T1["n"] == 20 and len(E["coord_def"]) == 12
    g, _ = handle(gpu, T1, T2, E)
    D = oracle.Dgevb(T1, T2, E)
    rng = np.random.default_rng(2)
    x = T1["xyz"][None] + rng.normal(0, 0.04, (48,) + T1["xyz"].shape)
    Vo, go = D.egrad(x)
    Vd, gd, _ = g.egrad(x)
    assert C.rel_err_E(Vd, Vo).max() < C.TOL_EG
    tol = np.maximum(C.TOL_EG, 2e-17 / torsion_conditioning(T1, x) ** 2)
    tol = np.maximum(tol, 4e-12 / wilson_conditioning(E, x))
    err = C.rel_err_G(gd.reshape(go.shape), go)
    assert (err < tol).all(), (err / tol).max()
    assert (tol < 1e-9).mean() > 0.8

"""Host-side set-up of the flexible SPC water box (pes WATER_SPC): what water_init.f90:53-107 and
set_periodic.f90:66-104 leave in modules evb_mod / qmdff / pbc_mod, as the dict RPMD.set_water takes.
The Fortran drivers keep doing this themselves; this mirror exists for the tests and benchmarks."""
import math

import numpy as np

BOHR = 0.52917721092       # general.f90:256
HARTREE = 627.5094743      # general.f90:257 (kcal/mol per hartree)


def water_pars():
    """water_pars(1:11) in atomic units (water_init.f90:75-101).  k_rr = 111.70765 and eps_OO = 0.1554 are written
    without a D exponent: REAL*4 literals in gfortran."""
    f32 = lambda v: float(np.float32(v))   # noqa: E731
    p = [1.0, 1.633, 101.9188, 2.567, 328.645606, -211.4672, f32(111.70765), 0.41, -0.82, 3.166, f32(0.1554)]
    p[0] = p[0] / BOHR
    p[1] = p[1] / BOHR
    p[2] = p[2] / HARTREE
    p[3] = p[3] * BOHR
    p[4] = p[4] / HARTREE * BOHR * BOHR
    p[5] = p[5] / HARTREE * BOHR * BOHR
    p[6] = p[6] / HARTREE * BOHR * BOHR
    p[9] = p[9] / BOHR
    p[10] = p[10] / HARTREE
    return np.array(p)


def water_box(nwater, periodic_angstrom=None, zahn=True, cut_coul_angstrom=10.0):
    """nwater molecules ordered O,H,H.  periodic_angstrom: the box lengths of the key file's `periodic` line (None:
    gas-phase cluster, no cut-off: egrad_water.f90:268)."""
    n = 3 * nwater
    pars = water_pars()
    q = np.tile([pars[8], pars[7], pars[7]], nwater)
    is_O = np.tile([1, 0, 0], nwater).astype(np.int32)
    periodic = periodic_angstrom is not None
    box = np.array(periodic_angstrom, dtype=np.float64) / BOHR if periodic else np.zeros(3)
    coul_cut, zahn_a, zahn_par = cut_coul_angstrom / BOHR, 0.0, 0.0
    if periodic:
        half = 0.5 * box.min()
        if coul_cut < 5.0:
            coul_cut = half - 0.1
        if zahn:
            zahn_a = 0.2 * BOHR
            coul_cut = 10.0 / BOHR
            zac = zahn_a * coul_cut
            zahn_par = math.erfc(zac) / coul_cut ** 2 + 2 * zahn_a / math.sqrt(math.pi) * math.exp(-zac ** 2) / coul_cut
        if coul_cut > half:
            coul_cut = half - 0.1
    return dict(n=n, periodic=int(periodic), zahn=int(zahn and periodic), box=box, coul_cut=coul_cut, zahn_a=zahn_a,
                zahn_par=zahn_par, pars=pars, q=q, is_O=is_O)


def water_lattice(nwater, box_angstrom, rng, jitter=0.05):
    """A liquid-like start structure: molecules at the equilibrium geometry the parameters encode (r_OH 1.0 A, r_HH
    1.633 A), random orientation, on a cubic lattice filling the box; bohr, [3 nwater, 3]."""
    pars = water_pars()
    r0, rhh = pars[0], pars[1]
    half = math.asin(rhh / 2 / r0)
    mol = np.array([[0, 0, 0], [r0 * math.sin(half), r0 * math.cos(half), 0], [-r0 * math.sin(half), r0 * math.cos(half), 0]])
    side = int(math.ceil(nwater ** (1 / 3)))
    a = box_angstrom / BOHR / side
    out = []
    for m in range(nwater):
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        cell = (np.array([m % side, (m // side) % side, m // (side * side)]) + 0.5) * a
        out.append(mol @ A.T + cell + rng.normal(0, jitter, (3, 3)))
    return np.concatenate(out)

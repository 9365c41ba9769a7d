"""Synthetic, well-formed QMDFF systems for the parity tests and benches of the force-field kernels.

The real tables come from qmdffgen / prepare.f90 / setnonb.f90 / copyc6.f90 (D3 reference data),
which are out of scope: the GPU library RECEIVES tables.  These generators produce
random but physically shaped tables (ethanol-like and formaldehyde-like molecules: bonds, angles,
proper torsions with 1-3 cosine terms, an inversion centre, intramolecular nci pairs of screening
classes 3..6, inter-molecular pairs through molnum) so that every branch of ff_eg / ff_nonb runs.
"""
import numpy as np

BOHR = 0.52917721092

# templates: (Z, xyz in Angstrom), bonds (1-based within molecule)
ETHANOL = dict(
    Z=[6, 6, 8, 1, 1, 1, 1, 1, 1],
    xyz=np.array([[0.00, 0.00, 0.00], [1.52, 0.00, 0.00], [2.00, 1.34, 0.00], [-0.40, 1.02, 0.00],
                  [-0.38, -0.52, 0.88], [-0.38, -0.52, -0.88], [1.90, -0.52, 0.88], [1.90, -0.52, -0.88],
                  [2.96, 1.30, 0.05]]),
    bonds=[(1, 2), (2, 3), (1, 4), (1, 5), (1, 6), (2, 7), (2, 8), (3, 9)], inversions=[])
FORMALDEHYDE = dict(
    Z=[6, 8, 1, 1],
    xyz=np.array([[0.0, 0.0, 0.0], [1.21, 0.0, 0.0], [-0.58, 0.94, 0.05], [-0.58, -0.94, 0.05]]),
    bonds=[(1, 2), (1, 3), (1, 4)], inversions=[(2, 1, 3, 4)])   # (i, centre j, k, l)


CHLOROMETHANE = dict(
    Z=[6, 17, 1, 1, 1],
    xyz=np.array([[0.0, 0.0, 0.0], [1.78, 0.0, 0.0], [-0.36, 1.03, 0.0], [-0.36, -0.51, 0.89], [-0.36, -0.51, -0.89]]),
    bonds=[(1, 2), (1, 3), (1, 4), (1, 5)], inversions=[])


def _hexane():
    """n-hexane, 20 atoms: zig-zag carbon chain (C1..C6 first), tetrahedral hydrogens; SURVEY.md 8(d) C4 asks for a
    ~20-atom two-state system with nat6 = 12"""
    cc, ch, th = 1.53, 1.09, np.deg2rad(111.5) / 2
    C = np.array([[i * cc * np.sin(th), (i % 2) * cc * np.cos(th), 0.0] for i in range(6)])
    xyz, bonds = list(C), [(i + 1, i + 2) for i in range(5)]
    for i in range(6):
        up = np.array([0.0, -1.0 if i % 2 == 0 else 1.0, 0.0])
        hs = [up * 0.55 + np.array([0, 0, 0.83]), up * 0.55 - np.array([0, 0, 0.83])]
        if i in (0, 5):
            hs.append(np.array([-0.9 if i == 0 else 0.9, 0.5 * (1.0 if i % 2 == 0 else -1.0), 0.0]))
        for h in hs:
            xyz.append(C[i] + ch * h / np.linalg.norm(h))
            bonds.append((i + 1, len(xyz)))
    xyz = np.array(xyz)
    # twist the all-trans chain about its three inner C-C bonds: dihedrals at exactly pi would put every torsion and
    # dihedral internal coordinate on the ill-conditioned point of the reference's acos formulas
    owner = list(range(6)) + [b[0] - 1 for b in bonds[5:]]          # carbon each atom hangs on
    for k, deg in ((1, 55.0), (2, -70.0), (3, 40.0)):               # bond C(k+1)-C(k+2), 0-based k
        a, b = xyz[k], xyz[k + 1]
        u = (b - a) / np.linalg.norm(b - a)
        t = np.deg2rad(deg)
        K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
        Rm = np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * (K @ K)
        for i in range(20):
            if owner[i] > k:
                xyz[i] = b + Rm @ (xyz[i] - b)
    return dict(Z=[6] * 6 + [1] * 14, xyz=xyz, bonds=bonds, inversions=[])


HEXANE = _hexane()


def _topology(tpl):
    n = len(tpl["Z"])
    nb = {i: set() for i in range(1, n + 1)}
    for a, b in tpl["bonds"]:
        nb[a].add(b)
        nb[b].add(a)
    angles = [(j, i, k) for j in nb for i in sorted(nb[j]) for k in sorted(nb[j]) if i < k]   # centre first
    tors = []
    for j, k in tpl["bonds"]:
        for i in sorted(nb[j] - {k}):
            for l in sorted(nb[k] - {j}):
                tors.append((i, j, k, l))
    dist = np.full((n + 1, n + 1), 99)
    for a in range(1, n + 1):
        dist[a, a] = 0
        frontier, d = {a}, 0
        seen = {a}
        while frontier:
            d += 1
            nxt = set()
            for u in frontier:
                for v in nb[u]:
                    if v not in seen:
                        seen.add(v)
                        dist[a, v] = d
                        nxt.add(v)
            frontier = nxt
    nci = [(a, b, min(6, int(dist[a, b]))) for a in range(1, n + 1) for b in range(a + 1, n + 1) if dist[a, b] >= 3]
    return angles, tors, nci


def make_system(nmol=8, seed=0, periodic=True, zahn=True, box_A=None, frac_formaldehyde=0.25, hb=False,
                frac_halogen=0.0, template=None):
    """hb=True adds the ff_hb tables (scalehb/scalexb, q_glob, an hb list); frac_halogen > 0 adds
    chloromethane molecules so that the X-bond branch (eabxag) runs."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(nmol ** (1 / 3)))
    spacing = 5.2  # Angstrom
    box_A = box_A or side * spacing
    Z, xyz, q, molnum = [], [], [], []
    bond, vbond, angl, vangl, tors, vtors, nci, hbl, vhb = [], [], [], [], [], [], [], [], []
    ldvt = 14
    off = 0
    for m in range(nmol):
        u = rng.random()
        tpl = CHLOROMETHANE if u < frac_halogen else (FORMALDEHYDE if u < frac_halogen + frac_formaldehyde else ETHANOL)
        if template is not None:
            tpl = template
        n = len(tpl["Z"])
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        cell = np.array([m % side, (m // side) % side, m // (side * side)]) * spacing + spacing / 2
        x = (tpl["xyz"] - tpl["xyz"].mean(axis=0)) @ A.T + cell + rng.normal(0, 0.03, (n, 3))
        xyz.append(x / BOHR)
        Z += tpl["Z"]
        ch = rng.normal(0, 0.25, n)
        q += list(ch - ch.mean())
        molnum += [m + 1] * n
        angles, torsl, ncil = _topology(tpl)
        for a, b in tpl["bonds"]:
            r0 = np.linalg.norm(tpl["xyz"][a - 1] - tpl["xyz"][b - 1]) / BOHR
            bond.append((a + off, b + off))
            vbond.append((r0 * rng.uniform(0.97, 1.03), rng.uniform(0.05, 0.25), rng.uniform(2.0, 6.0)))
        for t_i, (j, i, k) in enumerate(angles):
            v1 = tpl["xyz"][i - 1] - tpl["xyz"][j - 1]
            v2 = tpl["xyz"][k - 1] - tpl["xyz"][j - 1]
            th = np.arccos(v1 @ v2 / np.linalg.norm(v1) / np.linalg.norm(v2))
            # one (near-)linear reference angle to reach the pi - c0 < 1e-6 branch of ff_eg.f90:206
            th0 = np.pi if (m == 0 and t_i == 0) else th + rng.normal(0, 0.05)
            angl.append((j + off, i + off, k + off))
            vangl.append((th0, rng.uniform(0.02, 0.12)))
        for (i, j, k, l) in torsl:
            nt = int(rng.integers(1, 4))
            row = np.zeros(ldvt)
            row[0] = rng.uniform(0, np.pi)
            row[1] = rng.uniform(0.2, 1.0)
            for it in range(nt):
                row[2 + 3 * it:5 + 3 * it] = (float(rng.integers(1, 4)), np.pi * rng.integers(0, 2), rng.uniform(0.001, 0.01))
            tors.append((i + off, j + off, k + off, l + off, nt, 1))
            vtors.append(row)
        for t_i, (i, j, k, l) in enumerate(tpl["inversions"]):
            row = np.zeros(ldvt)
            row[0] = rng.uniform(0.0, 0.2)
            row[1] = rng.uniform(0.005, 0.03)
            row[2] = 0.0 if (m % 2 == 0) else 1.0          # both inversion forms (ff_eg.f90:560-570)
            tors.append((i + off, j + off, k + off, l + off, 1, 2))
            vtors.append(row)
        nci += [(a + off, b + off, c) for a, b, c in ncil]
        if hb and tpl is ETHANOL:          # intramolecular list entry (A, B, H): O, C1, H(O)
            hbl.append((3 + off, 1 + off, 9 + off))
            vhb.append((rng.uniform(0.1, 0.6), rng.uniform(0.1, 0.6)))
        if hb and tpl is CHLOROMETHANE:    # list entry whose third atom is a halogen -> eabxag
            hbl.append((1 + off, 3 + off, 2 + off))
            vhb.append((rng.uniform(0.1, 0.6), 0.0))
        off += n
    n = off
    xyz = np.concatenate(xyz)
    T = lambda: np.zeros((94, 94))   # noqa: E731
    r0ab, zab, r094, sr42 = T(), T(), T(), T()
    els = [1, 6, 8, 17]
    for a in els:
        for b in els:
            if a <= b:
                v = (rng.uniform(1.6, 2.4), rng.uniform(5.0, 60.0), rng.uniform(4.0, 6.0), rng.uniform(2.0, 9.0))
                for tab, val in zip((r0ab, zab, r094, sr42), v):
                    tab[a - 1, b - 1] = tab[b - 1, a - 1] = val
    rad = np.zeros(94)
    rad[0], rad[5], rad[7], rad[16] = 0.32, 0.75, 0.63, 0.99
    c6 = rng.uniform(5.0, 40.0, (n, n))
    c6 = 0.5 * (c6 + c6.T)
    eps1 = np.array([0, 0, 0.85, 1, 1, 1.0])      # setnonb.f90:164-169
    eps2 = np.array([0, 0, 0.5, 1, 1, 1.0])       # setnonb.f90:173-178
    L = box_A / BOHR
    coul_cut = 10.0 / BOHR if periodic else 50.0 / BOHR
    if periodic:
        coul_cut = min(coul_cut, L / 2 - 0.1)
    zahn_a = 0.2 * BOHR
    from math import erfc, exp, sqrt, pi
    zac = zahn_a * coul_cut
    zahn_par = erfc(zac) / coul_cut ** 2 + 2 * zahn_a / sqrt(pi) * exp(-zac ** 2) / coul_cut
    return dict(
        n=n, at=np.array(Z, dtype=np.int32), q=np.array(q), xyz=xyz, molnum=np.array(molnum, dtype=np.int32), nmols=nmol,
        bond=np.array(bond, dtype=np.int32).reshape(-1, 2), vbond=np.array(vbond).reshape(-1, 3),
        angl=np.array(angl, dtype=np.int32).reshape(-1, 3), vangl=np.array(vangl).reshape(-1, 2),
        tors=np.array(tors, dtype=np.int32).reshape(-1, 6), vtors=np.array(vtors).reshape(-1, ldvt), ldvt=ldvt,
        nci=np.array(nci, dtype=np.int32).reshape(-1, 3),
        c6xy=np.asfortranarray(c6), r0ab=np.asfortranarray(r0ab), zab=np.asfortranarray(zab),
        r094=np.asfortranarray(r094), sr42=np.asfortranarray(sr42), rad=rad, eps1=eps1, eps2=eps2,
        periodic=int(periodic), zahn=int(zahn and periodic), box=np.array([L, L, L]), coul_cut=coul_cut,
        vdw_cut=min(10.0 / BOHR, L / 2 - 0.1) if periodic else 10.0 / BOHR, cut_low=0.8 * coul_cut, zahn_a=zahn_a,
        zahn_par=zahn_par, e_zero=-1.2345, **_hb_tables(hb, hbl, vhb, q))


def _hb_tables(hb, hbl, vhb, q):
    if not hb:
        return dict(nhb=0)
    scalehb, scalexb = np.zeros(94), np.zeros(94)
    scalehb[6], scalehb[7], scalehb[8], scalehb[16] = 0.8, 0.3, 0.1, 2.0       # hbpara-scaled N, O, F, Cl
    scalexb[16], scalexb[34], scalexb[52] = 0.3, 0.6, 0.8
    return dict(nhb=len(hbl), hb=np.array(hbl, dtype=np.int32).reshape(-1, 3), vhb=np.array(vhb).reshape(-1, 2),
                scalehb=scalehb, scalexb=scalexb, q_glob=np.array(q))


def internals_np(coord_def, xyz):
    """Internal coordinates of one structure: type 1 distance, 2 angle, 3 dihedral, 4 out-of-plane with 1-based atoms
    (dist.f90, ang.f90, dihed.f90, oop.f90); used to place the synthetic Gaussian centres near the structure"""
    x = np.asarray(xyz, dtype=np.float64)
    out = []
    for cd in np.asarray(coord_def):
        t, a = int(cd[0]), [int(v) - 1 for v in cd[1:]]
        if t == 1:
            out.append(np.linalg.norm(x[a[1]] - x[a[0]]))
        elif t == 2:
            u, v = x[a[0]] - x[a[1]], x[a[2]] - x[a[1]]
            out.append(np.arccos(u @ v / (np.linalg.norm(u) * np.linalg.norm(v))))
        elif t == 3:
            u, v, w = x[a[0]] - x[a[1]], x[a[3]] - x[a[2]], x[a[2]] - x[a[1]]
            u, v, w = u / np.linalg.norm(u), v / np.linalg.norm(v), w / np.linalg.norm(w)
            su, sv = np.sqrt(1.0 - (u @ w) ** 2), np.sqrt(1.0 - (v @ w) ** 2)
            out.append(np.arccos(np.clip(np.cross(u, w) @ np.cross(v, w) / (su * sv), -1.0, 1.0)))
        else:
            v1, v2, v3 = x[a[3]] - x[a[0]], x[a[3]] - x[a[1]], x[a[3]] - x[a[2]]
            v1, v2, v3 = v1 / np.linalg.norm(v1), v2 / np.linalg.norm(v2), v3 / np.linalg.norm(v3)
            nv = np.cross(v1, v2) + np.cross(v2, v3) + np.cross(v3, v1)
            out.append(v1 @ np.cross(v2, v3) / np.linalg.norm(nv))
    return np.array(out)


def make_dgevb(seed=0, mode=3, npoints=5, template=None, internals=None):
    """A gas-phase two-state system: one ethanol-like molecule described by two QMDFFs with different
    parameters, a redundant-free internal coordinate set (bonds, angles, a dihedral, an out-of-plane)
    and random distributed-Gaussian parameters of the given mode (evb_pars.dat layout,
    read_pes.f90:2203-2225).  internals(coord_def, xyz) places the Gaussian centres (default: internals_np; the tests
    pass the CPU restatement's xyz_2int so that their fixtures keep their bits)."""
    T1 = make_system(nmol=1, seed=seed, periodic=False, frac_formaldehyde=0.0, hb=True, template=template)
    T2 = make_system(nmol=1, seed=seed, periodic=False, frac_formaldehyde=0.0, hb=True, template=template)
    rng = np.random.default_rng(seed + 100)
    T2 = dict(T2)
    T2["vbond"] = T1["vbond"] * rng.uniform(0.9, 1.1, T1["vbond"].shape)
    T2["vangl"] = T1["vangl"] * rng.uniform(0.9, 1.1, T1["vangl"].shape)
    T2["q"] = T1["q"] * 1.1
    T2["q_glob"] = T2["q"]
    T2["c6xy"] = np.asfortranarray(T1["c6xy"] * 0.9)
    T2["e_zero"] = T1["e_zero"] + 0.01
    T2["xyz"] = T1["xyz"]
    coord_def = np.array([[1, 1, 2, 0, 0], [1, 2, 3, 0, 0], [1, 1, 4, 0, 0], [1, 3, 9, 0, 0], [1, 2, 7, 0, 0],
                          [2, 1, 2, 3, 0], [2, 2, 3, 9, 0], [2, 4, 1, 2, 0], [3, 4, 1, 2, 3], [3, 1, 2, 3, 9],
                          [4, 3, 7, 8, 2]], dtype=np.int32)
    if template is HEXANE:   # 5 C-C bonds, a C-H bond, 3 C-C-C angles, 3 C-C-C-C dihedrals: nat6 = 12
        coord_def = np.array([[1, 1, 2, 0, 0], [1, 2, 3, 0, 0], [1, 3, 4, 0, 0], [1, 4, 5, 0, 0], [1, 5, 6, 0, 0],
                              [1, 3, 11, 0, 0], [2, 1, 2, 3, 0], [2, 2, 3, 4, 0], [2, 3, 4, 5, 0], [3, 1, 2, 3, 4],
                              [3, 2, 3, 4, 5], [3, 3, 4, 5, 6]], dtype=np.int32)
    nat6 = len(coord_def)
    E = dict(mode=mode, coord_def=coord_def, g_thres=1e-10)
    if internals is None:
        internals = internals_np
    pts = np.array([internals(coord_def, T1["xyz"] + rng.normal(0, 0.08, T1["xyz"].shape)) for _ in range(npoints)])
    mat = {1: npoints, 2: npoints * (1 + nat6), 3: npoints * (1 + nat6 + nat6 * (nat6 + 1) // 2)}[mode]
    E.update(point_int=pts, alph=rng.uniform(0.5, 2.5, npoints), b_vec=rng.normal(0, 2e-4, mat))
    return T1, T2, E

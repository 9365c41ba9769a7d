"""Tables of the reference's SHIPPED QMDFF examples (tests/golden/qmdff_examples.npz, made by
tests/golden/make_qmdff_fixtures.py) in the form the C-ABI and the oracle receive them (crcl_qmdff_tables).

The library receives what the Fortran side built; this module restates that host-side set-up for the
tests, routine by routine, so that the kernels run on the reference's own parameter sets instead of synthetic ones:
  prepare.f90:103-130   order of the set-up: setnonb, rdsolvff, ncoord_qmdff, getc6 for every pair
  setnonb.f90:37-189    a1, s8, a2, scalehb / scalexb, sr42, r094, zab, r0ab = 16.5 / r0**1.5, eps1 / eps2
  valel.f90:34-80       valence electrons (H, C, N, O only here)
  setr0.f90             r0 / autoang
  ncoord_qmdff.f90:36-70  D3 coordination numbers (k1 = 16, cut-off r^2 <= 5000)
  getc6.f90:36-83       C6 interpolation over the reference systems (k3 = -4)
  rdsolvff.f90:96-99,132-142  torsion phases * pi, vhb from hbpara and scalehb / scalexb
  set_periodic.f90:69-104 cut-offs, Zahn parameters
Literals without a D exponent are REAL*4 in gfortran: F32() marks them.
"""
import math
import os

import numpy as np

BOHR = 0.52917721092          # set_periodic.f90:55, general.f90:256
AUTOANG_D3 = 0.52917726       # setr0.f90 parameter autoang
_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qmdff_examples.npz")


def F32(x):
    return float(np.float32(x))


def _fixture():
    return np.load(_FIX)


def valel(z):
    """valel.f90:42-57 for Z <= 10"""
    return float(z - 2) if 2 < z <= 10 else float(z)


def ncoord(at, xyz, rcov_of, cn_thr=5000.0):
    """ncoord_qmdff.f90:47-67; xyz in bohr"""
    n = len(at)
    rc = np.array([rcov_of[int(z)] for z in at])
    cn = np.zeros(n)
    for i in range(n):
        d = xyz - xyz[i]
        r2 = np.einsum("ij,ij->i", d, d)
        m = (r2 <= cn_thr)
        m[i] = False
        rr = (rc[i] + rc[m]) / np.sqrt(r2[m])
        cn[i] = np.sum(1.0 / (1.0 + np.exp(-16.0 * (rr - 1.0))))
    return cn


def getc6(c6ab_ij, mxi, mxj, cni, cnj):
    """getc6.f90:52-81 for one pair; c6ab_ij = c6ab(iat,jat,:,:,:)"""
    rsum = csum = 0.0
    c6mem, r_save = -1e99, 10000.0
    for i in range(mxi):
        for j in range(mxj):
            c6 = c6ab_ij[i, j, 0]
            if c6 > 0:
                r = (c6ab_ij[i, j, 1] - cni) ** 2 + (c6ab_ij[i, j, 2] - cnj) ** 2
                if r < r_save:
                    r_save, c6mem = r, c6
                t = math.exp(-4.0 * r)
                rsum += t
                csum += t * c6
    return csum / rsum if rsum > 1e-99 else c6mem


def hbpara(a, b, q):
    """hbpara.f90: exp(-a q) / (exp(-a q) + b)"""
    return math.exp(-a * q) / (math.exp(-a * q) + b)


def tables(tag, e_zero=0.0, periodic_angstrom=None, zahn=True, xyz_bohr=None):
    """tag: 'box' | 'min1' | 'min2'.  periodic_angstrom: box lengths as on the key file's `periodic` line, or None
    for a gas-phase system.  Returns the dict caracal_b200.RPMD.set_qmdff / oracle.Qmdff take."""
    F = _fixture()
    g = lambda k: F[tag + "_" + k]   # noqa: E731
    els = [int(z) for z in F["const_elements"]]
    ix = {z: i for i, z in enumerate(els)}
    at = g("at").astype(np.int32)
    n = len(at)
    q = g("q").astype(np.float64)
    ref_xyz = g("xyz").astype(np.float64)
    molnum = g("molnum").astype(np.int32)
    nmols = int(molnum.max())
    # ---- setnonb.f90 ----
    a1, s8, a2 = F32(0.45), F32(2.7), F32(4.0)
    r2r4 = {z: float(F["const_r2r4"][ix[z]]) for z in els}
    rcov = {z: float(F["const_rcov"][ix[z]]) for z in els}
    rad = np.zeros(94)
    scalehb, scalexb = np.zeros(94), np.zeros(94)
    for z, v in ((7, 0.8), (8, 0.3), (9, 0.1), (15, 2.0), (16, 2.0), (17, 2.0), (34, 2.0), (35, 2.0)):
        scalehb[z - 1] = F32(v)
    for z, v in ((17, 0.30), (35, 0.60), (53, 0.80), (85, 1.00)):
        scalexb[z - 1] = F32(v)
    T = lambda: np.zeros((94, 94))   # noqa: E731
    sr42, r094, zab, r0ab = T(), T(), T(), T()
    for zi in els:
        rad[zi - 1] = float(F["const_rad"][ix[zi]])
        for zj in els:
            sr42[zj - 1, zi - 1] = 3.0 * s8 * r2r4[zi] * r2r4[zj]
            r094[zj - 1, zi - 1] = a1 * math.sqrt(3.0 * r2r4[zi] * r2r4[zj]) + a2
            zab[zj - 1, zi - 1] = valel(zi) * valel(zj)
            r0 = float(F["const_r0_angstrom"][ix[zi], ix[zj]]) / AUTOANG_D3
            r0ab[zj - 1, zi - 1] = 16.5 / r0 ** 1.5
    eps1 = np.array([0.0, 0.0, F32(0.85), 1.0, 1.0, 0.0])
    eps2 = np.array([0.0, 0.0, 0.5, 0.5, 1.0, 1.0])
    # ---- rdsolvff.f90: torsion phases, vhb ----
    tors = g("tors").astype(np.int32)
    vtors = g("vtors_raw").astype(np.float64).copy()
    for i in range(len(tors)):
        for j in range(int(tors[i, 4])):
            vtors[i, 3 * j + 3] *= math.pi
    hb = g("hb").astype(np.int32)
    vhb = np.zeros((len(hb), 2))
    for i, (ia, ib, ih) in enumerate(hb):
        if at[ih - 1] == 1:
            vhb[i, 0] = hbpara(10.0, 5.0, q[ia - 1]) * scalehb[at[ia - 1] - 1]
            vhb[i, 1] = hbpara(10.0, 5.0, q[ib - 1]) * scalehb[at[ib - 1] - 1]
        else:
            vhb[i, 0] = scalexb[at[ih - 1] - 1] * hbpara(-6.5, 1.0, q[ih - 1])
    # ---- prepare.f90:118-130: CN on the structure of the force-field file, C6 of every pair ----
    cn = ncoord(at, ref_xyz, rcov)
    c6xy = np.zeros((n, n))
    c6ab, maxci = F["const_c6ab"], F["const_maxci"]
    cache = {}
    for i1 in range(n):
        for i2 in range(i1 + 1):
            key = (int(at[i1]), int(at[i2]), round(cn[i1], 12), round(cn[i2], 12))
            if key not in cache:
                a, b = ix[int(at[i1])], ix[int(at[i2])]
                cache[key] = getc6(c6ab[a, b], int(maxci[a]), int(maxci[b]), cn[i1], cn[i2])
            c6xy[i2, i1] = c6xy[i1, i2] = cache[key]
    # ---- set_periodic.f90:69-104 (dynamic.f90:341-344 defaults: Zahn, 10 A cut-offs) ----
    periodic = periodic_angstrom is not None
    box = np.array(periodic_angstrom, dtype=np.float64) / BOHR if periodic else np.array([0.0, 0.0, 0.0])
    coul_cut, vdw_cut = 10.0 / BOHR, 10.0 / BOHR
    zahn_a = zahn_par = 0.0
    if periodic:
        half = 0.5 * box.min()
        if zahn:
            zahn_a = 0.2 * BOHR
            coul_cut = 10.0 / BOHR
            zac = zahn_a * coul_cut
            zahn_par = math.erfc(zac) / coul_cut ** 2 + 2 * zahn_a / math.sqrt(math.pi) * math.exp(-zac ** 2) / coul_cut
        if coul_cut > half:
            coul_cut = half - 0.1
    else:
        coul_cut = 50.0          # read_pes.f90:1754 (as written); vdw_cut 10 A for non-periodic systems, ff_nonb.f90:82-84
    out = dict(
        n=n, at=at, q=q, xyz=(ref_xyz if xyz_bohr is None else np.asarray(xyz_bohr, dtype=np.float64)), molnum=molnum,
        nmols=nmols, bond=g("bond").astype(np.int32), vbond=g("vbond").astype(np.float64),
        angl=g("angl").astype(np.int32), vangl=g("vangl").astype(np.float64), tors=tors, vtors=vtors,
        ldvt=vtors.shape[1], nci=g("nci").astype(np.int32), c6xy=np.asfortranarray(c6xy),
        r0ab=np.asfortranarray(r0ab), zab=np.asfortranarray(zab), r094=np.asfortranarray(r094),
        sr42=np.asfortranarray(sr42), rad=rad, eps1=eps1, eps2=eps2, periodic=int(periodic),
        zahn=int(zahn and periodic), box=box, coul_cut=coul_cut, vdw_cut=vdw_cut, cut_low=coul_cut, zahn_a=zahn_a,
        zahn_par=zahn_par, e_zero=float(e_zero), cn=cn)
    if len(hb) or nmols > 1:
        out.update(nhb=len(hb), hb=hb.reshape(-1, 3), vhb=vhb.reshape(-1, 2), scalehb=scalehb, scalexb=scalexb,
                   q_glob=q.copy())
    else:
        out.update(nhb=0)
    return out


def box_start_bohr():
    """examples/dynamic/ethanol_box/box.xyz, the xyzstart structure of equilibration.key"""
    return _fixture()["box_start_angstrom"] / BOHR


def dgevb_example():
    """examples/evbopt/DG-EVB: the two QMDFFs with the key file's eshift, coord_def.inp and every fifth frame of
    struc.xyz (the reaction path between the two minima)"""
    F = _fixture()
    e1, e2 = F["dgevb_eshift"]
    return tables("min1", e_zero=e1), tables("min2", e_zero=e2), F["dgevb_coord_def"].astype(np.int32), \
        F["dgevb_struc_angstrom"][::5] / BOHR


def dgevb_path():
    """all 42 path structures of struc.xyz (bohr) and their reference energies (comment lines, hartree)"""
    F = _fixture()
    return F["dgevb_struc_angstrom"] / BOHR, F["dgevb_struc_energy"]

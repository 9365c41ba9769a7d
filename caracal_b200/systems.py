"""Example systems of the path: the analytic gas-phase surfaces with a transition-state neighbourhood geometry and the
MECHA{} section a calc_rate key file would give for them.  Used by bench.py, __graft_entry__.smoke(), profiles/ and (re-exported
through tests/common.py) the parity tests, so that the product's bench does not depend on the test tree."""
import numpy as np

from .api import Mechanism, atomic_mass_au

BOHR = 0.52917721092  # general.f90:256


def h3_ts():
    """examples/calc_rate/h+h2/ts.xyz (collinear, 0.929764 A)."""
    return np.array([[0, 0, -0.929764359586], [0, 0, 0], [0, 0, 0.929764359586]]) / BOHR


def oh3_ts():
    """O at origin, r(OH)=0.97 A, H2 (0.76 A) approaching at r(O-H')=1.35 A."""
    return np.array([[0, 0, 0], [0.97 * np.cos(1.8), 0.97 * np.sin(1.8), 0], [1.35, 0, 0], [1.35 + 0.76, 0, 0]]) / BOHR


def ch5_ts():
    """CBE saddle point neighbourhood: atom order H,C,H,H,H,H_b; r(C-H')=1.39 A, r(H'-Hb)=0.873 A."""
    t = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]]) / np.sqrt(3)
    q = np.zeros((6, 3))
    q[0] = t[0] * 1.39
    q[2], q[3], q[4] = t[1] * 1.09, t[2] * 1.09, t[3] * 1.09
    q[5] = t[0] * (1.39 + 0.873)
    return q / BOHR


def brh2_ts():
    """Collinear saddle of the DIM-3C surface located with the oracle (profiles: r(H-Br) = 1.4402 A,
    r(H-H) = 1.3973 A, 21.0 kcal/mol above Br + H2); atom order H, Br, H (egrad_brh2.f:47-62)."""
    return np.array([[0, 0, 0], [0, 0, -2.72158888], [0, 0, 2.64056088]])


def o3_ts():
    """the shallow C2v minimum of the 1 1A" surface (oracle: r = 1.35709 A, 105.25 deg, +10.17 kcal/mol above
    O + O2), stretched along one bond towards O + O2; central atom first"""
    r1, r2, th = 1.60 / BOHR, 1.30 / BOHR, np.deg2rad(108.0)
    return np.array([[0, 0, 0], [r1, 0, 0], [r2 * np.cos(th), r2 * np.sin(th), 0]])


def ch4oh_ts():
    """examples/explore/ts_irc_ch4oh/ts_start.xyz of the reference (Angstrom): the start structure of its own saddle
    search on this surface; atom 4 is the hydrogen in flight.  Atom order H, C, H, H, H, O, H(O)."""
    return np.array([[-4.62878267, 1.25606861, 0.95459788], [-4.85261637, 2.15380812, 0.37457524],
                     [-4.27740626, 2.99438311, 0.76831501], [-4.53003946, 1.95346377, -0.88649386],
                     [-5.91912714, 2.37708958, 0.44643735], [-4.21407574, 1.75722671, -2.12170961],
                     [-3.93964920, 2.62885961, -2.44704966]]) / BOHR


def geh4oh_ts():
    """GeH4 + OH near the abstraction saddle region: tetrahedral GeH4 (r0ch = 1.525 A, egrad_geh4oh.f:2006) with the
    hydrogen in flight (atom 1) at 1.62 A, O 1.35 A beyond it, H(O) at 0.97 A and 100 deg.  Atom order
    H, Ge, H, H, H, O, H(O)."""
    t = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]]) / np.sqrt(3)
    q = np.zeros((7, 3))
    q[0] = t[0] * 1.62
    q[2], q[3], q[4] = t[1] * 1.525, t[2] * 1.525, t[3] * 1.525
    q[5] = t[0] * (1.62 + 1.35)
    e2 = t[1] - (t[1] @ t[0]) * t[0]
    e2 /= np.linalg.norm(e2)
    th = np.deg2rad(100.0)
    q[6] = q[5] + 0.97 * (np.cos(th) * (-t[0]) + np.sin(th) * e2)
    return q / BOHR


def ch4cn_ts():
    """CH4 + CN near its (early) abstraction saddle region: tetrahedral CH4 (r0ch = 1.094 A, egrad_ch4cn.f:2078) with the
    hydrogen in flight (atom 1) at 1.20 A, the carbon of CN 1.55 A beyond it, N at 1.172 A and 172 deg (the H-C-N
    reference angle is 180 deg; off the axis so that the bend's cross product is not the zero vector).  Atom order
    H, C, H, H, H, C(N), N."""
    t = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]]) / np.sqrt(3)
    q = np.zeros((7, 3))
    q[0] = t[0] * 1.20
    q[2], q[3], q[4] = t[1] * 1.094, t[2] * 1.094, t[3] * 1.094
    q[5] = t[0] * (1.20 + 1.55)
    e2 = t[1] - (t[1] @ t[0]) * t[0]
    e2 /= np.linalg.norm(e2)
    th = np.deg2rad(172.0)
    q[6] = q[5] + 1.172 * (np.cos(th) * (-t[0]) + np.sin(th) * e2)
    return q / BOHR


def _nh3_axes(theta_deg):
    """three unit vectors with mutual angle theta, on a cone around +z"""
    sa = np.sqrt((1.0 - np.cos(np.deg2rad(theta_deg))) / 1.5)
    ca = np.sqrt(1.0 - sa * sa)
    return np.array([[sa * np.cos(f), sa * np.sin(f), ca] for f in np.deg2rad([0.0, 120.0, 240.0])])


def clnh3_ts():
    """NH3 + Cl -> NH2 + HCl in its (late) saddle region: pyramidal NH3 (r = 1.014 A, H-N-H 107 deg) with the hydrogen in
    flight (atom 1) at 1.33 A from N and the chlorine 1.40 A beyond it on the N-H axis.  Atom order H, N, H, H, Cl
    (egrad_clnh3.f:85-91)."""
    u = _nh3_axes(107.0)
    q = np.zeros((5, 3))
    q[0] = u[0] * 1.33
    q[2], q[3] = u[1] * 1.014, u[2] * 1.014
    q[4] = u[0] * (1.33 + 1.40)
    return q / BOHR


def nh3oh_ts():
    """NH3 + OH -> NH2 + H2O in its (early) saddle region: pyramidal NH3 with the hydrogen in flight (atom 1) at 1.15 A,
    the oxygen 1.30 A beyond it, the hydroxyl hydrogen at 0.97 A and 100 deg to the O-H' axis.  Atom order
    H, N, H, H, O, H(O) (egrad_nh3oh.f BLOCK DATA: nnc = 2, nnb = 5, nnh = 1, 3, 4, nno = 6)."""
    u = _nh3_axes(107.0)
    q = np.zeros((6, 3))
    q[0] = u[0] * 1.15
    q[2], q[3] = u[1] * 1.014, u[2] * 1.014
    q[4] = u[0] * (1.15 + 1.30)
    e2 = np.cross(u[0], [0.0, 1.0, 0.0])
    e2 /= np.linalg.norm(e2)
    th = np.deg2rad(100.0)
    q[5] = q[4] + 0.97 * (np.cos(th) * (-u[0]) + np.sin(th) * e2)
    return q / BOHR


def h2co_ts():
    """formaldehyde on its way to H2 + CO (molecular channel): planar, both hydrogens swung to one side of the C-O axis,
    C-H 1.10 / 1.60 A, H-H 1.25 A.  Atom order C, O, H, H (main_h2co.f90:3182)."""
    q = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 1.17], [1.05, 0.0, -0.33], [1.30, 0.0, -0.93]])
    q[3] = q[2] + (q[3] - q[2]) / np.linalg.norm(q[3] - q[2]) * 1.25
    return q / BOHR


SYSTEMS = {
    "h3": dict(pes="h3", symbols=["H", "H", "H"], ts=h3_ts,
               # examples/calc_rate/h+h2/rate.key: reactant1 1 2, reactant2 3, bond_form 2-3, bond_break 1-2
               mecha=dict(bond_form=[[2, 3]], bond_break=[[1, 2]], reactants=[[1, 2], [3]], dist_inf=16.0)),
    "oh3": dict(pes="oh3", symbols=["O", "H", "H", "H"], ts=oh3_ts,
                mecha=dict(bond_form=[[1, 3]], bond_break=[[3, 4]], reactants=[[1, 2], [3, 4]], dist_inf=16.0)),
    "brh2": dict(pes="brh2", symbols=["H", "BR", "H"], ts=brh2_ts,
                 # Br + H2 -> HBr + H: reactant1 2, reactant2 1 3, bond_form 2-1, bond_break 1-3
                 mecha=dict(bond_form=[[2, 1]], bond_break=[[1, 3]], reactants=[[2], [1, 3]], dist_inf=16.0)),
    "o3": dict(pes="o3", symbols=["O", "O", "O"], ts=o3_ts,
               # O + O2 exchange: the bond 1-2 breaks (atom 2 leaves), fragments O2 (1,3) and O (2); bond_form 2-3
               mecha=dict(bond_form=[[2, 3]], bond_break=[[1, 2]], reactants=[[1, 3], [2]], dist_inf=16.0)),
    "ch4oh": dict(pes="ch4oh", symbols=["H", "C", "H", "H", "H", "O", "H"], ts=ch4oh_ts,
                  # CH4 + OH -> CH3 + H2O, atom 4 transferred: reactant1 1-5, reactant2 6 7, bond_form 4-6, bond_break 2-4
                  mecha=dict(bond_form=[[4, 6]], bond_break=[[2, 4]], reactants=[[1, 2, 3, 4, 5], [6, 7]], dist_inf=16.0)),
    "geh4oh": dict(pes="geh4oh", symbols=["H", "GE", "H", "H", "H", "O", "H"], ts=geh4oh_ts,
                   mecha=dict(bond_form=[[1, 6]], bond_break=[[2, 1]], reactants=[[1, 2, 3, 4, 5], [6, 7]], dist_inf=16.0)),
    "ch4cn": dict(pes="ch4cn", symbols=["H", "C", "H", "H", "H", "C", "N"], ts=ch4cn_ts,
                  mecha=dict(bond_form=[[1, 6]], bond_break=[[2, 1]], reactants=[[1, 2, 3, 4, 5], [6, 7]], dist_inf=16.0)),
    "clnh3": dict(pes="clnh3", symbols=["H", "N", "H", "H", "CL"], ts=clnh3_ts,
                  mecha=dict(bond_form=[[1, 5]], bond_break=[[2, 1]], reactants=[[1, 2, 3, 4], [5]], dist_inf=16.0)),
    "nh3oh": dict(pes="nh3oh", symbols=["H", "N", "H", "H", "O", "H"], ts=nh3oh_ts,
                  mecha=dict(bond_form=[[1, 5]], bond_break=[[2, 1]], reactants=[[1, 2, 3, 4], [5, 6]], dist_inf=16.0)),
    # H2 + CO -> H2CO read as an association: both C-H bonds form, the H-H bond breaks
    "h2co": dict(pes="h2co", symbols=["C", "O", "H", "H"], ts=h2co_ts,
                 mecha=dict(bond_form=[[1, 3], [1, 4]], bond_break=[[3, 4]], reactants=[[1, 2], [3, 4]], dist_inf=16.0)),
    "ch4h": dict(pes="ch4h", symbols=["H", "C", "H", "H", "H", "H"], ts=ch5_ts,
                 # reactant1 1 2 3 4 5, reactant2 6, bond_form 1-6, bond_break 2-1
                 mecha=dict(bond_form=[[1, 6]], bond_break=[[2, 1]], reactants=[[1, 2, 3, 4, 5], [6]], dist_inf=16.0)),
}


def masses(name):
    return np.array([atomic_mass_au(s) for s in SYSTEMS[name]["symbols"]])


def mechanism(name, dist_inf=None):
    """dist_inf: R_inf in bohr as module evb_mod holds it.  The key file gives DIST_INF in Angstrom and
    calc_rate_read.f90:693 divides by bohr; pass 16.0 / BOHR for the shipped examples.  The parity
    fixtures use the default of the table above (16 bohr)."""
    s = SYSTEMS[name]
    kw = dict(s["mecha"])
    if dist_inf is not None:
        kw["dist_inf"] = dist_inf
    return Mechanism(ts_struc=s["ts"](), **kw)


def ring_polymer(name, nbeads, rng, spread=0.05):
    ts = SYSTEMS[name]["ts"]()
    return ts[None] + rng.normal(0, spread, (nbeads,) + ts.shape)

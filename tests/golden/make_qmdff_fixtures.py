"""Generates tests/golden/qmdff_examples.npz from the reference's SHIPPED QMDFF example inputs (run in the build
container, where /root/reference exists; the GPU box only sees the committed fixture):

  examples/dynamic/ethanol_box/box.qmdff + box.xyz   (1125 atoms, 125 ethanol; SURVEY.md 8(d) C5 parity input)
  examples/evbopt/DG-EVB/min1.qmdff, min2.qmdff, coord_def.inp, struc.xyz   (6 atoms; C4 parity input)

and the element constants the reference's set-up reads from DATA statements, for the elements those files
contain (H, C, N, O): r2r4 / rcov / rad (setnonb.f90:43-110), the D3 cut-off radii (setr0.f90) and the D3
reference C6 table c6ab (copyc6.f90 `pars`, decoded as copyc6.f90:35688-35705 + limit.f90 do).  Only DATA is taken
from the reference -- the set-up arithmetic itself (prepare.f90:103-130, setnonb.f90, ncoord_qmdff.f90, getc6.f90,
rdsolvff.f90's hbpara step, set_periodic.f90:69-104) is restated in tests/qmdff_file.py and runs at test time.

gfortran reads a literal without a D exponent as REAL*4: r2r4, rcov and the setr0 table are such
literals and are stored here already rounded to single precision; rad and pars carry D exponents.

Usage: python tests/golden/make_qmdff_fixtures.py
"""
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[DdEe][-+]?\d+)?")
ELEMENTS = [1, 6, 7, 8]


def numbers(text):
    return [float(t.replace("D", "E").replace("d", "e")) for t in NUM.findall(text)]


def constructor(src, name):
    """values of `name = (/ ... /)` (first occurrence)"""
    m = re.search(r"^\s*" + re.escape(name) + r"\s*=\s*\(/(.*?)/\)", src, re.S | re.M)
    return np.array(numbers(m.group(1).replace("&", " ")))


def chunks(src, name):
    """values of all `name( a: b)=(/ ... /)` chunks, concatenated in order of a"""
    out = {}
    for m in re.finditer(re.escape(name) + r"\(\s*(\d+)\s*:\s*(\d+)\s*\)\s*=\s*\(/(.*?)/\)", src, re.S):
        vals = numbers(m.group(3).replace("&", " "))
        assert len(vals) == int(m.group(2)) - int(m.group(1)) + 1, (name, m.group(1))
        out[int(m.group(1))] = vals
    return np.array([v for k in sorted(out) for v in out[k]])


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def element_constants():
    strip = lambda s: "\n".join(l for l in s.splitlines() if not l.lstrip().startswith("!"))   # noqa: E731
    setnonb = strip(open(os.path.join(REF, "src/setnonb.f90")).read())
    r2r4, rcov, rad = f32(constructor(setnonb, "r2r4")), f32(constructor(setnonb, "rcov")), constructor(setnonb, "rad")
    assert len(r2r4) == len(rcov) == len(rad) == 94
    r0 = f32(chunks(strip(open(os.path.join(REF, "src/setr0.f90")).read()), "r0ab"))
    assert len(r0) == 4465
    r0mat = np.zeros((94, 94))
    k = 0
    for i in range(94):                      # setr0.f90: do i; do j = 1, i
        for j in range(i + 1):
            r0mat[i, j] = r0mat[j, i] = r0[k]
            k += 1
    pars = chunks(strip(open(os.path.join(REF, "src/copyc6.f90")).read()), "pars")
    assert len(pars) == 161925
    rec = pars.reshape(-1, 5)
    c6ab = -np.ones((94, 94, 5, 5, 3))
    maxci = np.zeros(94, dtype=np.int32)
    for c6, ci, cj, cn1, cn2 in rec:
        iat, jat = int(ci), int(cj)
        iadr, jadr = 1 + (iat - 1) // 100, 1 + (jat - 1) // 100      # limit.f90
        iat, jat = iat - 100 * (iadr - 1), jat - 100 * (jadr - 1)
        maxci[iat - 1] = max(maxci[iat - 1], iadr)
        maxci[jat - 1] = max(maxci[jat - 1], jadr)
        c6ab[iat - 1, jat - 1, iadr - 1, jadr - 1] = (c6, cn1, cn2)
        c6ab[jat - 1, iat - 1, jadr - 1, iadr - 1] = (c6, cn2, cn1)
    e = np.array(ELEMENTS) - 1
    return dict(elements=np.array(ELEMENTS, dtype=np.int32), r2r4=r2r4[e], rcov=rcov[e], rad=rad[e],
                r0_angstrom=r0mat[np.ix_(e, e)], c6ab=c6ab[np.ix_(e, e)], maxci=maxci[e])


def read_qmdff(path):
    """rdsolvff.f90:47-176 / rdo.f90: the file as written by qmdffgen (list-directed reads)"""
    tok = open(path).read().split("\n")
    n = int(float(tok[0].split()[0]))
    atoms = np.array([[float(x) for x in tok[2 + i].split()] for i in range(n)])
    cnt = [int(x) for x in tok[2 + n].split()]
    nbond, nangl, ntors, nhb, nnci = cnt[:5]
    body = " ".join(tok[3 + n:]).split()
    pos = [0]

    def take(k, conv=float):
        v = [conv(x) for x in body[pos[0]:pos[0] + k]]
        pos[0] += k
        return v
    bond, vbond, angl, vangl, tors, vt = [], [], [], [], [], []
    for _ in range(nbond):
        bond.append(take(2, int))
        vbond.append(take(3))
    for _ in range(nangl):
        angl.append(take(3, int))
        vangl.append(take(2))
    for _ in range(ntors):
        t = take(6, int)
        tors.append(t)
        vt.append(take(2 + 3 * t[4]))
    hb = np.array(take(3 * nhb, int), dtype=np.int32).reshape(-1, 3)
    nci = np.array(take(3 * nnci, int), dtype=np.int32).reshape(-1, 3)
    assert pos[0] == len(body), (path, pos[0], len(body))
    ldvt = max(5, 2 + 3 * max(t[4] for t in tors)) if tors else 5
    vtors = np.zeros((ntors, ldvt))
    for i, row in enumerate(vt):
        vtors[i, :len(row)] = row
    return dict(at=atoms[:, 0].astype(np.int32), xyz=atoms[:, 1:4], q=atoms[:, 4], molnum=atoms[:, 5].astype(np.int32),
                bond=np.array(bond, dtype=np.int32).reshape(-1, 2), vbond=np.array(vbond).reshape(-1, 3),
                angl=np.array(angl, dtype=np.int32).reshape(-1, 3), vangl=np.array(vangl).reshape(-1, 2),
                tors=np.array(tors, dtype=np.int32).reshape(-1, 6), vtors_raw=vtors, hb=hb, nci=nci)


def read_xyz(path, frame=0):
    lines = open(path).read().split("\n")
    n = int(lines[0].split()[0])
    blk = lines[frame * (n + 2) + 2:frame * (n + 2) + 2 + n]
    return np.array([[float(x) for x in l.split()[1:4]] for l in blk])


def main():
    out = {"const_" + k: v for k, v in element_constants().items()}
    box = read_qmdff(os.path.join(REF, "examples/dynamic/ethanol_box/box.qmdff"))
    assert set(box["at"]) <= set(ELEMENTS)
    out.update({"box_" + k: v for k, v in box.items()})
    out["box_start_angstrom"] = read_xyz(os.path.join(REF, "examples/dynamic/ethanol_box/box.xyz"))
    out["box_periodic_angstrom"] = np.array([27.0, 27.0, 27.0])          # equilibration.key: periodic 27.0 27.0 27.0
    for tag in ("min1", "min2"):
        ff = read_qmdff(os.path.join(REF, "examples/evbopt/DG-EVB/%s.qmdff" % tag))
        assert set(ff["at"]) <= set(ELEMENTS)
        out.update({tag + "_" + k: v for k, v in ff.items()})
    out["dgevb_eshift"] = np.array([-132.2974220, -132.3079820])          # evbopt.key: eshift
    cd = [[int(x) for x in l.split()] for l in open(os.path.join(REF, "examples/evbopt/DG-EVB/coord_def.inp")) if l.split()]
    out["dgevb_coord_def"] = np.array([[len(c) - 1] + c + [0] * (4 - len(c)) for c in cd], dtype=np.int32)   # 2 atoms: bond (type 1), 3: angle, 4: dihedral
    spath = os.path.join(REF, "examples/evbopt/DG-EVB/struc.xyz")
    out["dgevb_struc_angstrom"] = np.array([read_xyz(spath, f) for f in range(42)])
    # the comment line of every frame holds the reference (QM) energy of that path structure
    out["dgevb_struc_energy"] = np.array([float(open(spath).read().split("\n")[8 * f + 1]) for f in range(42)])
    path = os.path.join(HERE, "qmdff_examples.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

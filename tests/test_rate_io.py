"""On-disk formats and restart markers of a calc_rate run (calc_rate.f90:1106-1148,1420-1478,1690-1734,
1896-1904,1992-1997; recross.f90:134-226,420-440,652-666): what the host pipeline leaves behind is what the reference's
own readers take, a run through files equals a run in memory, and an interrupted run resumes where it stopped.
CPU only: the rank-local compute is the oracle behind the product's host interface (tests/oracle_handle.py)."""
import os

import numpy as np

from caracal_b200 import rate as R
from caracal_b200 import rate_io as IO
from tests import common as C
from tests.oracle_handle import OracleRPMD

KW = dict(umbr_lo=0.94, umbr_hi=1.02, umbr_dist=0.02, gen_steps=20, equi_steps=10, umbr_steps=30, umbr_traj=3, xi_min=0.94,
          xi_max=1.02, nbins=100, recr_equi=16, child_tot=24, child_interv=4, child_point=4, child_evol=12,
          andersen_step=10, npaths=2, pmf_minloc="PMF_MIN")


def handles():
    name, nb, kelvin = "h3", 2, 300.0
    m, mech = C.masses(name), C.mechanism(name)
    beta, dt = C.beta_calc_rate(kelvin), C.dt_au(0.1)
    g, g1 = OracleRPMD(name, nb, m, beta, dt), OracleRPMD(name, 1, m, beta, dt)
    for h in (g, g1):
        h.set_mechanism(mech)
        h.set_seed(C.SEED)
    return g, g1, m, mech, kelvin, beta


def run(workdir=None, **extra):
    g, g1, m, mech, kelvin, beta = handles()
    return R.calc_rate(g, g1, C.h3_ts(), m, mech, kelvin, beta, workdir=workdir, names=["H", "H", "H"], **dict(KW, **extra))


def fortran_list_read(line):
    """what `read(unit,*) a, b, c` takes from a line"""
    return [float(t) for t in line.replace(",", " ").split()]


def test_names_and_numbers():
    assert IO.folder_name(300.0, 8) == "300K_8bead"
    assert IO.bias_name(0.95) == "bias_0.9500" and IO.bias_name(-0.05) == "bias_-0.0500" and IO.bias_name(1.0) == "bias_1.0000"
    for v in (0.5, -1.5, 1e-3, 123.456, 0.0, 1.0 / 3.0, -2.5e-7, 6.02e23):
        s = IO.fortran_real(v)
        assert float(s) == v and len(s) >= 26


def test_run_through_files_equals_run_in_memory_and_leaves_the_reference_files(tmp_path):
    mem = run(None)
    out = run(str(tmp_path))
    d = os.path.join(str(tmp_path), "300K_2bead")
    for f in ("current_calc", "start_finished", "equilibrated_struc.xyz", "xi_pos.dat", "xi_equi_real.dat",
              "equilibrated_ens.dat", "sampling_finished", "umbr_int.dat", "pmf_integration.dat", "recross_status",
              "recross_num_tmp.dat", "recross_denom_tmp.dat", "recross_parent_pos.dat", "recrossing_time.dat",
              "recross_finished"):
        assert os.path.exists(os.path.join(d, f)), f
    # the structures make the round trip through Angstrom text (17 digits): rounding level only
    assert np.abs(out["struc_equi"] - mem["struc_equi"]).max() < 1e-14
    assert np.abs(out["average"] - mem["average"]).max() < 1e-12 and np.abs(out["variance"] - mem["variance"]).max() < 1e-12
    assert abs(out["kappa"] - mem["kappa"]) < 1e-12 and abs(out["delta_w_kj"] - mem["delta_w_kj"]) < 1e-9
    # statistics/bias_<xi>: 2 header lines + umbr_traj lines + blank + header + averaged line (calc_rate.f90:1441)
    nwin, ntraj = len(out["xi_wins"]), KW["umbr_traj"]
    for w, xi in enumerate(out["xi_wins"]):
        lines = open(os.path.join(d, "statistics", IO.bias_name(xi))).read().split("\n")[:-1]
        assert len(lines) == ntraj + 5 and lines[0].lstrip().startswith("#") and lines[ntraj + 3].strip() == "# Averaged values:"
        rows = [fortran_list_read(l) for l in lines[2:2 + ntraj]]
        assert [int(r[0]) for r in rows] == list(range(1, ntraj + 1))
        x, a, v = fortran_list_read(lines[ntraj + 4])          # read(50,*) xi_val,average(i),variance(i)  (:1890)
        assert abs(x - xi) < 1e-15 and a == out["average"][w] and v == out["variance"][w]
        assert abs(a - sum(r[1] for r in rows) / ntraj) < 1e-15
    # equilibrated_struc.xyz as calc_rate.f90:1302-1313 reads it
    toks = open(os.path.join(d, "equilibrated_struc.xyz")).read().split("\n")
    assert int(toks[0]) == 3 and toks[1].split()[0] == "ideal:" and toks[2].split()[0] == "H"
    assert abs(float(toks[2].split()[1]) / IO.BOHR - out["struc_equi"][0, 0, 0]) < 1e-14
    # umbr_int.dat / pmf_integration.dat / recrossing_time.dat
    rows = [fortran_list_read(l) for l in open(os.path.join(d, "umbr_int.dat")).read().split("\n")[1:-1]]
    assert len(rows) == nwin and abs(rows[1][1] - out["average"][1]) < 1e-15
    pm = [fortran_list_read(l) for l in open(os.path.join(d, "pmf_integration.dat")).read().split("\n")[1:-1]]
    assert len(pm) == KW["nbins"] - 1 and abs(max(r[1] for r in pm) - out["pmf"][:-1].max() * 2625.50) < 1e-9
    rt = [fortran_list_read(l) for l in open(os.path.join(d, "recrossing_time.dat")).read().split("\n")[4:-1]]
    assert len(rt) == KW["child_evol"] and abs(rt[-1][1] - out["kappa_t"][-1]) < 1e-15
    assert abs(rt[0][0] - C.dt_au(0.1) * float(np.float32(2.41888428E-2))) < 1e-15    # back to 0.1 fs
    assert int(open(os.path.join(d, "recross_status")).read()) == KW["child_tot"] // KW["child_point"]
    assert abs(float(open(os.path.join(d, "recross_finished")).read()) - out["kappa_t"][-1]) < 1e-15


def test_interrupted_run_resumes_from_the_files(tmp_path):
    full = run(str(tmp_path / "a"))
    # the same run, interrupted: one statistics file removed and the marker with it; recrossing stopped after 2 of 3 rounds
    d = str(tmp_path / "b")
    part = run(d, rounds_per_launch=1)
    assert abs(part["kappa"] - full["kappa"]) < 1e-12         # chunked launches add up to the single launch
    f = IO.RunFolder(d, 300.0, 2)
    xi_kill = full["xi_wins"][2]
    os.remove(f.stats_path(xi_kill))
    os.remove(f.f("sampling_finished"))
    os.remove(f.f("recross_finished"))
    st, num, den, qpar = f.recross_resume(KW["child_evol"], 2, 3)
    assert st == 6 and qpar is not None and qpar.shape == (2, 3, 3)
    # wind the recrossing files back to "4 of 6 bunches done" with the sums of the first two rounds
    g, g1, m, mech, kelvin, beta = handles()
    q_start = np.repeat(full["struc_equi"][int(np.argmin(np.abs(full["xi_wins"] - full["xi_barrier"])))][None], 2, axis=0)
    calls = []
    orig = g.recross_children
    g.recross_children = lambda *a, **k: (calls.append((a[1], k.get("pair0", 0))), orig(*a, **k))[1]
    n2, d2, _, _ = R.recrossing(g, q_start, full["xi_barrier"], 15.0, kelvin, KW["recr_equi"], KW["child_tot"],
                                KW["child_interv"], KW["child_point"], KW["child_evol"], folder=None)
    # (sharded / folder-less call: one launch over all pairs)
    assert calls == [(12, 0)]
    # partial sums of rounds 0 and 1 = pairs [0, 12): written as the checkpoint of an interrupted run
    g.recross_children = orig
    ts = int(np.argmin(np.abs(full["xi_wins"] - full["xi_barrier"])))
    parents = R.recrossing(g, q_start, full["xi_barrier"], 15.0, kelvin, KW["recr_equi"], KW["child_tot"],
                           KW["child_interv"], KW["child_point"], KW["child_evol"])[2]
    npart, dpart, _ = g.recross_children(parents, 2 * 6, KW["child_evol"], full["xi_barrier"], pair0=0)[:3]
    f.recross_checkpoint(4, npart, dpart, parents[-1])
    res = run(d)
    assert abs(res["kappa"] - full["kappa"]) < 1e-12
    assert np.abs(res["average"] - full["average"]).max() < 1e-12
    assert os.path.exists(f.stats_path(xi_kill)) and os.path.exists(f.f("sampling_finished"))
    assert int(open(f.f("recross_status")).read()) == 6
    # a finished window is not run again: its file is untouched (same bytes, older mtime is enough a proxy: compare text)
    w0 = open(f.stats_path(full["xi_wins"][0])).read()
    assert w0 == open(IO.RunFolder(str(tmp_path / "a"), 300.0, 2).stats_path(full["xi_wins"][0])).read()


def test_stats_resume_line_count_rule(tmp_path):
    """calc_rate.f90:1441-1478: umbr_traj + 5 lines = done; otherwise the trajectories on file are kept"""
    f = IO.RunFolder(str(tmp_path), 300.0, 4)
    assert f.stats_resume(0.9, 5) == (1, 0.0, 0.0)
    assert f.stats_write(0.9, 1, [0.91, 0.92], [1e-4, 2e-4], 5) is None
    first, sa, sv = f.stats_resume(0.9, 5)
    assert first == 3 and abs(sa - 1.83) < 1e-15 and abs(sv - 3e-4) < 1e-18
    a, v = f.stats_write(0.9, 3, [0.93, 0.94, 0.95], [3e-4, 4e-4, 5e-4], 5, sa, sv)
    assert abs(a - 0.93) < 1e-15 and abs(v - 3e-4) < 1e-18
    assert f.stats_resume(0.9, 5)[0] == 6
    assert f.stats_read([0.9], 5)[0][0] == a

"""On-disk formats and restart markers of a calc_rate.x run, so that a run through the GPU path
leaves the files the reference leaves and resumes from the files the reference resumes from.

    <T>K_<n>bead/                       working folder (calc_rate.f90:705-730)
      current_calc                      marker: the folder belongs to a running / unfinished calculation (:737)
      equilibrated_struc.xyz            start structures of all windows in Angstrom, centre of mass removed (:1106-1140)
      xi_pos.dat, xi_equi_real.dat,     the windows' xi, the xi actually reached, the last energy (:1107-1113)
      equilibrated_ens.dat
      start_finished                    marker (:1148)
      statistics/bias_<xi>              per window: 2 header lines, one line per trajectory (j, average, variance),
                                        blank, "# Averaged values:", (xi, average, variance) (:1507-1512,1690-1707);
                                        line-count based resume (:1420-1478)
      sampling_finished                 marker (:1734)
      umbr_int.dat                      xi, average, variance of every window (:1896-1904)
      pmf_integration.dat               bin centre, PMF in kJ/mol (:1992-1997)
      recross_status                    number of finished child bunches (recross.f90:134-141,434-436)
      recross_num_tmp.dat               child_evol lines: accumulated numerators (:420-424)
      recross_denom_tmp.dat             accumulated denominator (:425-427)
      recross_parent_pos.dat            the reference opens it with status="replace" and then READS from it
                                        (recross.f90:428-433), so its file is always empty and a restart continues from the
                                        start structure; here the parent positions are written (bohr, one atom per line,
                                        beads outer) so that a restart really continues the parent chain
      recrossing_time.dat               t (fs), kappa(t) (:652-660)
      recross_finished                  final kappa (:664-666)

Numbers are written the way gfortran's list-directed output writes REAL(8) (17 significant digits); every reader here
accepts anything Fortran's list-directed READ accepts (blank separated fields), so files of either origin are read.
"""
import math
import os

import numpy as np

BOHR = 0.52917721092          # general.f90:256


def fortran_real(x):
    """REAL(8) the way gfortran's list-directed output shows it: 17 significant digits, F form for 0.1 <= |x| < 1e16
    (`   1.5000000000000000     `, `  0.50000000000000000     `), else E form with a three-digit exponent
    (`   1.0000000000000000E-003`).  Only the digits matter to a reader; the blanks follow gfortran for diff-ability."""
    x = float(x)
    if x != x:
        return "                       NaN"
    if math.isinf(x):
        return "                  Infinity" if x > 0 else "                 -Infinity"
    a, sign = abs(x), ("-" if (x < 0 or (x == 0 and math.copysign(1.0, x) < 0)) else "")
    if a == 0.0:
        body = "0.0000000000000000"
    elif 0.1 <= a < 1e16:
        e = int(math.floor(math.log10(a))) + 1                 # digits in front of the point (<= 0: none)
        body = "%.*f" % (17 - max(e, 0), a)
        if len(body.replace(".", "").lstrip("0")) > 17:        # rounding carried into a new leading digit
            body = "%.*f" % (17 - max(e + 1, 0), a)
    else:
        m, ex = ("%.16E" % a).split("E")
        return (sign + m).rjust(22) + "E%+04d" % int(ex)
    return (sign + body).rjust(22) + "    "


def folder_name(kelvin, nbeads):
    """calc_rate.f90:705-728: "<int(T)>K_<nbeads>bead" """
    return "%dK_%dbead" % (int(kelvin), int(nbeads))


def bias_name(xi):
    """calc_rate.f90:1391-1399: statistics/bias_<f6.4 of |xi|>, a '-' in front for negative xi"""
    return "bias_%s%6.4f" % ("-" if xi < 0.0 else "", abs(xi))


def touch(path):
    with open(path, "a"):
        pass


class RunFolder:
    """The working folder of one (temperature, bead count) and its restart markers"""

    def __init__(self, root, kelvin, nbeads):
        self.path = os.path.join(root, folder_name(kelvin, nbeads))
        os.makedirs(self.path, exist_ok=True)
        touch(self.f("current_calc"))

    def f(self, *names):
        return os.path.join(self.path, *names)

    def has(self, marker):
        return os.path.exists(self.f(marker))

    def finish(self):
        """calc_rate.f90 removes current_calc when the whole rate calculation is through"""
        if self.has("current_calc"):
            os.remove(self.f("current_calc"))

    # ---- phase 1: start structures (calc_rate.f90:1106-1148, read back at :1302-1313) ----------------------------
    def write_start_structures(self, names, xi_wins, start_xis, struc_equi, equi_energy=None):
        nwin, natoms = struc_equi.shape[0], struc_equi.shape[1]
        with open(self.f("equilibrated_struc.xyz"), "w") as fs, open(self.f("xi_pos.dat"), "w") as fx, \
                open(self.f("xi_equi_real.dat"), "w") as fr, open(self.f("equilibrated_ens.dat"), "w") as fe:
            fr.write(" # These are the Xi values for the structures in equilibrated_struc.xyz\n")
            for i in range(nwin):
                fs.write(" %11d\n" % natoms)
                fe.write(fortran_real(0.0 if equi_energy is None else equi_energy[i]) + "\n")
                fx.write(fortran_real(xi_wins[i]) + "\n")
                fr.write(fortran_real(start_xis[i]) + "\n")
                fs.write(" ideal:" + fortran_real(xi_wins[i]) + " real:" + fortran_real(start_xis[i]) + "\n")
                for j in range(natoms):
                    fs.write(" %-2s" % names[j] + "".join(fortran_real(v * BOHR) for v in struc_equi[i, j]) + "\n")
        touch(self.f("start_finished"))

    def read_start_structures(self):
        """-> xi_wins[nwin], struc_equi[nwin, natoms, 3] in bohr, exactly as calc_rate.f90:1302-1313 reads them
        (the structures of a run always make this round trip through the file)"""
        xi = np.array([float(line.split()[0]) for line in open(self.f("xi_pos.dat")) if line.strip()])
        toks = open(self.f("equilibrated_struc.xyz")).read().split("\n")
        struc, pos = [], 0
        for _ in range(len(xi)):
            natoms = int(toks[pos].split()[0])
            rows = [toks[pos + 2 + j].split() for j in range(natoms)]
            struc.append([[float(r[1]) / BOHR, float(r[2]) / BOHR, float(r[3]) / BOHR] for r in rows])
            pos += 2 + natoms
        return xi, np.array(struc)

    # ---- phase 2: statistics/bias_<xi> -------------------------------------------------------------------------------
    def stats_path(self, xi):
        os.makedirs(self.f("statistics"), exist_ok=True)
        return self.f("statistics", bias_name(xi))

    def stats_resume(self, xi, umbr_traj):
        """calc_rate.f90:1420-1478: -> (first trajectory still to run (1-based), sum of averages, sum of variances) of
        the trajectories already on file; a file with umbr_traj + 5 lines is complete (first = umbr_traj + 1)"""
        path = self.stats_path(xi)
        if not os.path.exists(path):
            return 1, 0.0, 0.0
        lines = open(path).read().split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        num_lines = len(lines)
        if num_lines == umbr_traj + 5:
            return umbr_traj + 1, 0.0, 0.0
        num_remain = umbr_traj - num_lines + 2
        sa = sv = 0.0
        for ln in lines[2:num_lines]:
            t = ln.split()
            sa += float(t[1])
            sv += float(t[2])
        return umbr_traj - num_remain + 1, sa, sv

    def stats_write(self, xi, first_traj, avg, var, umbr_traj, sum_avg0=0.0, sum_var0=0.0):
        """appends trajectories first_traj.. (1-based) with their (average, variance) and, when the window is complete,
        the averaged block (calc_rate.f90:1507-1512,1690-1707); -> (window average, window variance) or None"""
        path = self.stats_path(xi)
        with open(path, "a" if first_traj > 1 else "w") as f:
            if first_traj == 1:
                f.write(" # The distribution characteristics for the actual umbrella window:\n")
                f.write(" #             Traj-No.        average           variance\n")
            for k, (a, v) in enumerate(zip(avg, var)):
                f.write(" %11d" % (first_traj + k) + fortran_real(a) + fortran_real(v) + "\n")
            if first_traj - 1 + len(avg) < umbr_traj:
                return None
            a = (sum_avg0 + float(np.sum(avg))) / umbr_traj
            v = (sum_var0 + float(np.sum(var))) / umbr_traj
            f.write("\n # Averaged values:\n" + fortran_real(xi) + fortran_real(a) + fortran_real(v) + "\n")
        return a, v

    def stats_read(self, xi_wins, umbr_traj):
        """calc_rate.f90:1869-1893: the averaged line of every window"""
        avg, var = np.zeros(len(xi_wins)), np.zeros(len(xi_wins))
        for i, xi in enumerate(xi_wins):
            lines = open(self.stats_path(xi)).read().split("\n")
            t = lines[2 + umbr_traj + 2].split()
            avg[i], var[i] = float(t[1]), float(t[2])
        return avg, var

    def write_umbr_int(self, xi_wins, average, variance):
        """calc_rate.f90:1896-1904"""
        with open(self.f("umbr_int.dat"), "w") as f:
            f.write(" # xi_position      average(xi)       variance(xi)\n")
            for x, a, v in zip(xi_wins, average, variance):
                f.write(fortran_real(x) + fortran_real(a) + fortran_real(v) + "\n")

    def write_pmf(self, bin_coord, pmf_hartree):
        """calc_rate.f90:1992-1997: nbins-1 lines, kJ/mol"""
        with open(self.f("pmf_integration.dat"), "w") as f:
            f.write(" # xi-value     PMF(kJ/mol)\n")
            for i in range(len(pmf_hartree) - 1):
                f.write(fortran_real(bin_coord[i]) + fortran_real(pmf_hartree[i] * 2625.50) + "\n")

    # ---- phase 4: recrossing restart files (recross.f90:134-226,420-440,634-666) ---------------------------------
    def recross_resume(self, child_evol, nbeads, natoms):
        """-> (recross_status, num_total[child_evol], denom_total, parent q or None); (0, zeros, 0.0, None) when
        nothing usable is on file (every incomplete file resets the status, as the reference does)"""
        zero = (0, np.zeros(child_evol), 0.0, None)
        try:
            status = int(open(self.f("recross_status")).read().split()[0])
        except (OSError, ValueError, IndexError):
            return zero
        if status == 0:
            return zero
        try:
            num = np.array([float(t) for t in open(self.f("recross_num_tmp.dat")).read().split()])
            den = float(open(self.f("recross_denom_tmp.dat")).read().split()[0])
        except (OSError, ValueError, IndexError):
            return zero
        if len(num) < child_evol:
            return zero
        q = None
        try:
            v = np.array([float(t) for t in open(self.f("recross_parent_pos.dat")).read().split()])
            if len(v) == nbeads * natoms * 3:
                q = v.reshape(nbeads, natoms, 3)
        except (OSError, ValueError):
            pass
        return status, num[:child_evol].copy(), den, q

    def recross_checkpoint(self, bunches_done, num_total, denom_total, q_parent):
        with open(self.f("recross_num_tmp.dat"), "w") as f:
            f.write("".join(fortran_real(v) + "\n" for v in num_total))
        with open(self.f("recross_denom_tmp.dat"), "w") as f:
            f.write(fortran_real(denom_total) + "\n")
        with open(self.f("recross_parent_pos.dat"), "w") as f:
            for b in range(q_parent.shape[0]):
                for a in range(q_parent.shape[1]):
                    f.write("".join(fortran_real(v) for v in q_parent[b, a]) + "\n")
        with open(self.f("recross_status"), "w") as f:
            f.write(" %11d\n" % bunches_done)

    def write_recrossing_time(self, num_total, denom_total, dt_au):
        """recross.f90:652-660; 2.41888428E-2 is a REAL*4 literal there"""
        fs = float(np.float32(2.41888428E-2))
        with open(self.f("recrossing_time.dat"), "w") as f:
            f.write(" # This is the time dependent recrossing factor calculated with EVB-QMDFF!\n #\n")
            f.write(" #  t(fs)           kappa(t)    \n #\n")
            for i, v in enumerate(num_total):
                f.write(fortran_real((i + 1) * dt_au * fs) + fortran_real(v / denom_total) + "\n")

    def recross_finished(self, kappa=None):
        """write (kappa given) or read (-> kappa or None) the final recrossing factor"""
        if kappa is not None:
            with open(self.f("recross_finished"), "w") as f:
                f.write(fortran_real(kappa) + "\n")
            return kappa
        try:
            return float(open(self.f("recross_finished")).read().split()[0])
        except (OSError, ValueError, IndexError):
            return None
